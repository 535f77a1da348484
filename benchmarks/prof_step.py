"""Minimal driver for ncu: the headline step (configs[1]) a few times on one stream, nothing else.
    python benchmarks/prof_step.py [--steps 8] [--tokens 100000000] [--decode]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fm_index_constrained_decoding_b200 import FMIndex, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--tokens", type=int, default=100_000_000)
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--decode", action="store_true", help="also run the fused decode step (256 beams, fp32 + bf16)")
args = ap.parse_args()
dev = torch.device("cuda", 0)
toks, beg = synth.zipf_corpus(args.tokens, seed=0, device=dev)
qsets = [synth.sample_queries(toks, args.batch, seed=100 + i, beginnings=beg) for i in range(8)]
ix = FMIndex(device=0)
ix.initialize_from_device_tokens(toks, beg)
V = 128256
out = None
for k in range(3 + args.steps):
    q, offs = qsets[k % 8]
    out = ix.query_batch(q, offs, V, out=out)
torch.cuda.synchronize()
if args.decode:
    q, offs = synth.sample_queries(toks, 256, seed=7, beginnings=beg)
    for dt in (torch.float32, torch.bfloat16):
        lg = torch.randn(256, V, device=dev).to(dt)
        ob = torch.empty_like(lg)
        for _ in range(4):
            sp, ep = ix.get_range_batch(q, offs)
            ix.mask_logits(sp, ep, lg, out=ob, always_allow=128009)
    torch.cuda.synchronize()
print("done")
