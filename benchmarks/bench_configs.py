#!/usr/bin/env python
"""Secondary measurements for the BASELINE.json configs that are not the headline bench line
(bench.py = configs[1]).  One JSON line per measurement on stdout.

    python benchmarks/bench_configs.py [--tokens 100000000] [--what decode,locate,sweep,negative]

  decode   configs[3]: fused constrained-decode step, 256 beams x 128,256 logits (fp32 and bf16):
           K1 search of the beam histories + K2 enumerate/mask fused with the logits rows
  locate   configs[4] (single-shard part): 65,536 prefixes, locate the first K=8 rows of every range,
           extract 32 tokens at every hit; sampled (sa_dens 32 / isa_dens 64) and resident
           (sa_dens 1, keep_text) variants
  sweep    K1/K2 throughput vs batch size (4,096 .. 65,536 prefixes)
  negative configs[1] with 25% corrupted prefixes (early exit)
All timings: CUDA events on the launching stream, >= 3 warm-ups, inputs larger than L2.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VOCAB, EOS, SHIFT = 128256, 128009, 10


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=100_000_000)
    ap.add_argument("--what", default="config0,decode,locate,sweep,negative")
    ap.add_argument("--steps", type=int, default=50)
    args = ap.parse_args()
    what = set(args.what.split(","))
    from fm_index_constrained_decoding_b200 import FMIndex, synth
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    peak = 6529.1
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    if "config0" in what:
        # configs[0]: sample corpus (stand-in tokenizer, tests/golden/sample_corpus.json), batch 1, through the
        # reference-shaped scalar API: latency, not throughput (one launch + sync per call on the GPU)
        g = json.load(open(os.path.join(ROOT, "tests", "golden", "sample_corpus.json")))
        from oracle.oracle import OracleFMIndex
        small = FMIndex(device=0)
        small.initialize(g["docs"], in_memory=True)
        o = OracleFMIndex()
        o.initialize(g["docs"])
        qs0 = [q["q"] for q in g["queries"]]

        def run(ixx, reps):
            t = time.perf_counter()
            for _ in range(reps):
                for q in qs0:
                    r = ixx.get_range(q)
                    ixx.get_distinct_count(*r)
            return (time.perf_counter() - t) / (reps * len(qs0)) * 1e6
        run(small, 2)
        gpu_us = run(small, 10)
        run(o, 2)
        cpu_us = run(o, 10)
        print(json.dumps({"config": "configs[0] sample_corpus.tsv (stand-in tokenizer), batch 1: get_range + get_distinct_count",
                          "tokens": g["size"] - 1, "gpu_us_per_query": gpu_us, "cpu_oracle_us_per_query": cpu_us,
                          "note": "latency-dominated: each scalar GPU call is launch + synchronise; the CPU oracle keeps the "
                                  "reference's 3 x sigma allocation per distinct_count call"}), flush=True)
        del small, o
    if not (what - {"config0"}):
        return
    toks, beg = synth.zipf_corpus(args.tokens, seed=0, device=dev)
    t0 = time.time()
    ix = FMIndex(device=0)
    ix.initialize_from_device_tokens(toks, beg)
    torch.cuda.synchronize()
    base = {"tokens": args.tokens, "build_s": time.time() - t0, "index_mb": ix.info()["device_bytes"] / 1e6}

    if "decode" in what:
        B = 256
        qs = [synth.sample_queries(toks, B, seed=300 + i, beginnings=beg) for i in range(8)]
        for dtype, path in ((torch.float32, "pipe"), (torch.float32, "tma"), (torch.float32, "regs"), (torch.bfloat16, "pipe"), (torch.bfloat16, "regs")):
            os.environ["FMIDX_LOGITS_PATH"] = path
            g = torch.Generator(device=dev).manual_seed(3)
            # 8 distinct logits buffers (8 x 131 MB fp32 > L2) rotated so no step re-reads a cached row
            logits = [torch.randn(B, VOCAB, generator=g, device=dev).to(dtype) for _ in range(8)]
            out = torch.empty_like(logits[0])
            k = [0]

            def step():
                q, offs = qs[k[0] % 8]
                sp, ep = ix.get_range_batch(q, offs)
                ix.mask_logits(sp, ep, logits[k[0] % 8], out=out, always_allow=EOS)
                k[0] += 1

            def mask_only():
                ix.mask_logits(sp0, ep0, logits[k[0] % 8], out=out, always_allow=EOS)
                k[0] += 1
            ms = timed(step, args.steps)
            sp0, ep0 = ix.get_range_batch(*qs[0])
            ms_mask = timed(mask_only, args.steps)
            # the same kernel when every beam has >= 8 tokens of history (narrow ranges only)
            sp0, ep0 = ix.get_range_batch(*synth.sample_queries(toks, B, len_min=8, seed=999, beginnings=beg))
            ms_mask_narrow = timed(mask_only, args.steps)
            sp0, ep0 = ix.get_range_batch(*qs[0])
            nbytes = 2 * B * VOCAB * logits[0].element_size() + B * 16032
            print(json.dumps({"config": "configs[3] fused decode step: 256 beams x 128256 logits", "dtype": str(dtype), "logits_staging": path, **base,
                              "ms_per_step": ms, "steps_per_sec": 1e3 / ms, "beams_per_sec": B * 1e3 / ms,
                              "mask_logits_kernel_ms": ms_mask, "mask_logits_kernel_ms_narrow_beams": ms_mask_narrow, "algorithmic_bytes": nbytes,
                              "mask_logits_GBps": nbytes / (ms_mask * 1e-3) / 1e9, "frac_of_measured_hbm": nbytes / (ms_mask * 1e-3) / 1e9 / peak,
                              "floor_ms_at_peak": nbytes / (peak * 1e9) * 1e3}), flush=True)
            del logits, out

    if "sweep" in what or "negative" in what:
        for B in ([4096, 16384, 65536] if "sweep" in what else [4096]):
            for neg in ([0.0, 0.25] if "negative" in what else [0.0]):
                qs = [synth.sample_queries(toks, B, seed=500 + i, beginnings=beg, negative_frac=neg) for i in range(4)]
                k = [0]
                sp, ep = ix.get_range_batch(*qs[0])
                mask = torch.empty((B, (VOCAB + 31) // 32), dtype=torch.int32, device=dev)

                def search():
                    ix.get_range_batch(*qs[k[0] % 4])
                    k[0] += 1

                def masks():
                    ix.distinct_mask_batch(sp, ep, VOCAB, out=mask)
                ms_s = timed(search, 20)
                ms_m = timed(masks, 20)
                steps = int((qs[0][1][1:] - qs[0][1][:-1]).sum())
                print(json.dumps({"config": "configs[1] sweep", "batch": B, "negative_frac": neg, **base, "search_ms": ms_s,
                                  "search_qps": B / ms_s * 1e3, "search_GBps_at_1088B_per_step": 1088 * steps / (ms_s * 1e-3) / 1e9,
                                  "mask_ms": ms_m, "masks_per_sec": B / ms_m * 1e3,
                                  "search_plus_mask_qps": B / (ms_s + ms_m) * 1e3}), flush=True)
                del mask

    if "locate" in what:
        Q, K, M = 65536, 8, 32
        q, offs = synth.sample_queries(toks, Q, seed=700, beginnings=beg)
        for label, kw in [("sampled sa_dens=32 isa_dens=64 (reference defaults)", {"layout": "sampled"}),
                          ("resident sa_dens=1 isa_dens=1 keep_text (default)", {})]:
            if kw:
                ix2 = FMIndex(device=0, **kw)
                ix2.initialize_from_device_tokens(toks, beg)
            else:
                ix2 = ix
            sp, ep = ix2.get_range_batch(q, offs)
            rows = sp.unsqueeze(1) + torch.arange(K, device=dev).unsqueeze(0)
            valid = rows < ep.unsqueeze(1)
            rows = torch.where(valid, rows, torch.zeros_like(rows)).contiguous().view(-1)
            n_hits = int(valid.sum())
            ms_search = timed(lambda: ix2.get_range_batch(q, offs), 10)
            ms_loc_hits = timed(lambda: ix2.locate_batch(rows), 10)
            # throughput on independent rows: most prefixes of this corpus have < 8 matches, so the hit rows above
            # repeat row 0; uniform random rows are the honest locate/extract load
            gr = torch.Generator(device=dev).manual_seed(11)
            rrows = torch.randint(0, ix2.size(), (Q * K,), generator=gr, device=dev)
            ms_loc = timed(lambda: ix2.locate_batch(rrows), 10)
            pos = ix2.locate_batch(rrows)
            begins = pos.clamp(min=0)
            ends = (begins + M).clamp(max=ix2.size())
            ms_ext = timed(lambda: ix2.extract_batch(begins, ends, M), 10)
            print(json.dumps({"config": "configs[4] single shard: 65536 prefixes, locate K=8 + extract 32", "variant": label,
                              "tokens": args.tokens, "index_mb": ix2.info()["device_bytes"] / 1e6, "search_ms": ms_search,
                              "search_qps": Q / ms_search * 1e3, "valid_hits": n_hits, "locate_hit_rows_ms": ms_loc_hits,
                              "locate_random_rows_ms": ms_loc, "rows_located": Q * K,
                              "locates_per_sec": Q * K / ms_loc * 1e3, "extract_ms": ms_ext,
                              "extracts_per_sec": Q * K / ms_ext * 1e3, "tokens_extracted_per_sec": Q * K * M / ms_ext * 1e3}), flush=True)
            if kw:
                del ix2


if __name__ == "__main__":
    main()
