"""Fused decode step (configs[3]) probe: 256 beams x 128,256 logits, fp32 / bf16, unigram table on / off, history mixes.
    python benchmarks/decode_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fm_index_constrained_decoding_b200 import FMIndex, synth  # noqa: E402

dev = torch.device("cuda", 0)
V, EOS, R = 128256, 128009, 256
toks, beg = synth.zipf_corpus(100_000_000, seed=0, device=dev)
ix = FMIndex(device=0)
ix.initialize_from_device_tokens(toks, beg)
PEAK = 6529.1
g = torch.Generator(device=dev).manual_seed(3)
for hist in ("len1-32", "len8-32"):
    q, offs = synth.sample_queries(toks, R, len_min=1 if hist == "len1-32" else 8, seed=7, beginnings=beg)
    sp, ep = ix.get_range_batch(q, offs)
    for table in (1, 0):
        ix.tune(unigram_table=table)
        for name, dt, esz in (("fp32", torch.float32, 4), ("bf16", torch.bfloat16, 2)):
            bufs = [torch.randn(R, V, device=dev, generator=g, dtype=torch.float32).to(dt) for _ in range(8)]
            out = torch.empty_like(bufs[0])
            for k in range(5):
                ix.mask_logits(sp, ep, bufs[k % 8], out=out, always_allow=EOS)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 300
            e0.record()
            for k in range(n):
                ix.mask_logits(sp, ep, bufs[k % 8], out=out, always_allow=EOS)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            gb = R * V * esz * 2 / (ms * 1e-3) / 1e9
            print("%s table=%d %s: %.1f us  %.0f GB/s  %.3f of HBM" % (hist, table, name, ms * 1e3, gb, gb / PEAK), flush=True)
            del bufs, out
