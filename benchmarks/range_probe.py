"""The OVERLAPPED headline step under `ncu --replay-mode range`: counters of the whole multi-stream region (what one
kernel-at-a-time profiling cannot see).   ncu --replay-mode range --metrics ... python benchmarks/range_probe.py [--phase P]"""
import argparse
import ctypes as C
import os
import sys
import faulthandler
faulthandler.dump_traceback_later(400, exit=True)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fm_index_constrained_decoding_b200 import FMIndex, synth, _native  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--phase", type=int, default=0)
ap.add_argument("--steps", type=int, default=64)
ap.add_argument("--pipes", type=int, default=8)
ap.add_argument("--lazy", type=int, default=0, help="fmidx_tune lazy_clear level for the profiled steps")
args = ap.parse_args()
dev = torch.device("cuda", 0)
V, SHIFT, B, NP = 128256, 10, 4096, args.pipes
WORDS = (V + 31) // 32
toks, beg = synth.zipf_corpus(100_000_000, seed=0, device=dev)
qsets = [synth.sample_queries(toks, B, seed=100 + i, beginnings=beg) for i in range(8)]
ix = FMIndex(device=0)
ix.initialize_from_device_tokens(toks, beg)
L = _native.lib()
streams = [torch.cuda.Stream(device=dev) for _ in range(NP)]
bufs = [(torch.empty(B, dtype=torch.int64, device=dev), torch.empty(B, dtype=torch.int64, device=dev),
         torch.empty((B, WORDS), dtype=torch.int32, device=dev)) for _ in range(NP)]
cargs = {}
for qi, (q, offs) in enumerate(qsets):
    for pi in range(NP):
        sp, ep, m = bufs[pi]
        cargs[(qi, pi)] = (ix._h, pi, C.c_void_p(q.data_ptr()), C.c_void_p(offs.data_ptr()), B, SHIFT, V, C.c_void_p(sp.data_ptr()),
                           C.c_void_p(ep.data_ptr()), C.c_void_p(m.data_ptr()), C.c_void_p(streams[pi].cuda_stream))
ix.tune(lazy_clear=args.lazy)
for k in range(4 * NP):
    L.fmidx_query_batch(*cargs[((k + k // NP) % 8, k % NP)])
torch.cuda.synchronize()
ix.tune(probe_phase=args.phase)
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
for k in range(4 * NP, 4 * NP + args.steps):
    L.fmidx_query_batch(*cargs[((k + k // NP) % 8, k % NP)])
torch.cuda.synchronize()
rt.cudaProfilerStop()
print("range done: %d steps, %d in flight, phase %d" % (args.steps, NP, args.phase))
