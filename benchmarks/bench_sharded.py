#!/usr/bin/env python
"""Sharded configs of BASELINE.json on N GPUs of one box (one process per GPU, torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29577 \
        benchmarks/bench_sharded.py [--what c3,c4,c5] [--c3-tokens 1000000000] [--c5-tokens 4000000000]

  c3  configs[2]: 1 B-token corpus sharded by document over the N GPUs (STRONG scaling: each GPU holds 1B/N
      tokens); 4,096 replicated prefixes; per-shard masks OR-reduce-scattered, counts all-reduced
  c4  configs[3] sharded: fused decode step, 256 beams x 128,256 fp32 logits; every rank masks the B/N rows it owns
      with the pieces of the other shards OR-ed inside the fused kernel
  c5  configs[4]: 4 B-token corpus, N x (4B/N)-token shards; 65,536 prefixes; locate the first K=8 rows of every
      range on every shard, all-gather the [N, 65536, 8] global positions, extract 32 tokens at each local hit
Rank 0 prints one JSON line per config.  Device time = max over ranks (CUDA events + all-reduce MAX).
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VOCAB, EOS = 128256, 128009


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="c3,c4,c5")
    ap.add_argument("--c3-tokens", type=int, default=1_000_000_000)
    ap.add_argument("--c5-tokens", type=int, default=4_000_000_000)
    ap.add_argument("--steps", type=int, default=30)
    args = ap.parse_args()
    what = set(args.what.split(","))
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    os.environ["FMIDX_DEVICE"] = str(local)
    sys.stdout.flush()
    out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)  # library banners -> stderr
    dist.init_process_group("nccl", device_id=dev)
    from fm_index_constrained_decoding_b200 import FMIndex, sharded, synth

    def emit(d):
        if rank == 0:
            out.write(json.dumps(d) + "\n")
            out.flush()

    def build_shard(total_tokens):
        per = total_tokens // world
        toks, beg = synth.zipf_corpus(per, seed=1000 * rank + 7, device=dev)
        t0 = time.time()
        ix = FMIndex(device=local)
        ix.initialize_from_device_tokens(toks, beg)
        torch.cuda.synchronize()
        return toks, beg, ix, time.time() - t0, per

    if "c3" in what or "c4" in what:
        toks, beg, ix, t_build, per = build_shard(args.c3_tokens)
        sh = sharded.ShardedFMIndex(ix, doc_base=0, pos_base=rank * per)
        if "c3" in what:
            B = 4096
            # queries: substrings of rank 0's shard, broadcast so every rank searches the same prefixes
            q, offs = synth.sample_queries(toks, B, seed=100, beginnings=beg)
            n_tok = torch.tensor([q.numel()], device=dev)
            dist.broadcast(n_tok, 0)
            if rank != 0:
                q = torch.empty(int(n_tok[0]), dtype=torch.int32, device=dev)
            dist.broadcast(q, 0)
            dist.broadcast(offs, 0)
            res = {}

            def step():
                sp, ep = sh.get_range_batch(q, offs)
                res["mask"] = sh.distinct_mask_batch(sp, ep, VOCAB)
                res["count"] = sh.count_batch(sp, ep)
            ms = timed(step, args.steps)
            words = (VOCAB + 31) // 32
            emit({"config": "configs[2] 1B-token corpus doc-sharded over N GPUs, 4096 replicated prefixes: search + masks OR-reduce-scattered + counts all-reduced",
                  "n_gpus": world, "tokens_total": per * world, "tokens_per_gpu": per, "build_s_per_shard": t_build,
                  "ms_per_step": ms, "prefix_queries_per_sec": B / ms * 1e3, "masks_per_sec": B / ms * 1e3,
                  "nccl_bytes_sent_per_gpu_per_step": (world - 1) * (B // world) * words * 4 + B * 8,
                  "matched_somewhere": int((res["count"] > 0).sum()), "scaling": "strong (corpus fixed, shards shrink)"})
        if "c4" in what:
            B = 256
            own = B // world
            q, offs = synth.sample_queries(toks, B, seed=300, beginnings=beg)
            n_tok = torch.tensor([q.numel()], device=dev)
            dist.broadcast(n_tok, 0)
            if rank != 0:
                q = torch.empty(int(n_tok[0]), dtype=torch.int32, device=dev)
            dist.broadcast(q, 0)
            dist.broadcast(offs, 0)
            g = torch.Generator(device=dev).manual_seed(3 + rank)
            logits = [torch.randn(own, VOCAB, generator=g, device=dev) for _ in range(8)]
            k = [0]

            def step4():
                sp, ep = sh.get_range_batch(q, offs)
                sh.mask_logits_rows(sp, ep, logits[k[0] % 8], always_allow=EOS)
                k[0] += 1
            ms = timed(step4, args.steps)
            emit({"config": "configs[3] sharded: fused decode step, 256 beams x 128256 fp32 logits, index doc-sharded over N GPUs, "
                            "each rank masks the B/N rows it owns (other shards' mask pieces OR-ed inside the fused kernel)",
                  "n_gpus": world, "tokens_total": per * world, "ms_per_step": ms, "steps_per_sec": 1e3 / ms, "beams_per_sec": B / ms * 1e3})
        del ix, sh, toks, beg
        torch.cuda.empty_cache()

    if "c5" in what:
        toks, beg, ix, t_build, per = build_shard(args.c5_tokens)
        sh = sharded.ShardedFMIndex(ix, doc_base=0, pos_base=rank * per)
        Q, K, M = 65536, 8, 32
        q, offs = synth.sample_queries(toks, Q, seed=700, beginnings=beg)
        n_tok = torch.tensor([q.numel()], device=dev)
        dist.broadcast(n_tok, 0)
        if rank != 0:
            q = torch.empty(int(n_tok[0]), dtype=torch.int32, device=dev)
        dist.broadcast(q, 0)
        dist.broadcast(offs, 0)
        res = {}

        def step5():
            sp, ep = sh.get_range_batch(q, offs)
            hits, valid = sh.locate_topk(sp, ep, K)          # [N, Q, K] global positions on every rank
            mine = hits[rank].clamp(min=0) - sh.pos_base      # local hits -> extract 32 tokens at each
            begins = mine.view(-1)
            ends = (begins + M).clamp(max=ix.size())
            res["text"], res["len"] = ix.extract_batch(begins, ends, M)
            res["valid"] = valid
        ms = timed(step5, 10)
        emit({"config": "configs[4] 4B-token corpus, N shards, 65536 prefixes: search + locate first K=8 rows per shard + all-gather of "
                        "[N,65536,8] global positions + extract 32 tokens at every local hit slot",
              "n_gpus": world, "tokens_total": per * world, "tokens_per_gpu": per, "build_s_per_shard": t_build,
              "index_gb_per_gpu": ix.info()["device_bytes"] / 1e9, "ms_per_step": ms, "prefix_queries_per_sec": Q / ms * 1e3,
              "locates_per_sec_all_gpus": world * Q * K / ms * 1e3, "tokens_extracted_per_sec_all_gpus": world * Q * K * M / ms * 1e3,
              "valid_hits_all_shards": int(res["valid"].sum()), "all_gather_bytes_per_gpu": world * Q * K * 8})
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
