"""Phase-isolated throughput of the headline step with graph replay (no host launch bound): 8 steps in flight.
    python benchmarks/overlap_probe2.py"""
import ctypes as C
import os
import sys
import time
import faulthandler
faulthandler.dump_traceback_later(150, exit=True)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fm_index_constrained_decoding_b200 import FMIndex, synth, _native  # noqa: E402
from fm_index_constrained_decoding_b200.index import DeviceGraphSteps  # noqa: E402

dev = torch.device("cuda", 0)
V, SHIFT, B, NP = 128256, 10, 4096, 8
WORDS = (V + 31) // 32
toks, beg = synth.zipf_corpus(100_000_000, seed=0, device=dev)
qsets = [synth.sample_queries(toks, B, seed=100 + i, beginnings=beg) for i in range(8)]
ix = FMIndex(device=0)
ix.initialize_from_device_tokens(toks, beg)
L = _native.lib()
streams = [torch.cuda.Stream(device=dev) for _ in range(NP)]
sp = [C.c_void_p(s.cuda_stream) for s in streams]
maxtok = max(q.numel() for q, _ in qsets)
g = {}
for phase, first in ((0, 0), (1, 8), (2, 16)):
    ix.tune(probe_phase=phase)
    g[phase] = DeviceGraphSteps(ix, NP, B, maxtok, V, first_slot=first, buffers=[(qsets[i][0], qsets[i][1]) for i in range(NP)])
ix.tune(probe_phase=0)
# phase-2 graphs need valid ranges in their sp/ep buffers
for i in range(NP):
    g[2].sp[i].copy_(g[0].sp[i]); g[2].ep[i].copy_(g[0].ep[i])
torch.cuda.synchronize()
zbuf = [torch.empty((B, WORDS), dtype=torch.int32, device=dev) for _ in range(NP)]


def timed(fn, steps=400):
    for k in range(40):
        fn(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    for s in streams:
        s.wait_event(e0)
    for k in range(steps):
        fn(k)
    for s in streams:
        e = torch.cuda.Event(); e.record(s); main.wait_event(e)
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3


def rep(phase, first):
    def f(k):
        L.fmidx_query_graph_launch_on(ix._h, first + k % NP, sp[k % NP])
    return f


rt = C.CDLL("libcudart.so.12")
rt.cudaMemsetAsync.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]
zp = [C.c_void_p(z.data_ptr()) for z in zbuf]


def memset_only(k):
    rt.cudaMemsetAsync(zp[k % NP], 0, B * WORDS * 4, sp[k % NP])


def memset_plus_search(k):
    rt.cudaMemsetAsync(zp[k % NP], 0, B * WORDS * 4, sp[k % NP])
    L.fmidx_query_graph_launch_on(ix._h, 8 + k % NP, sp[k % NP])


def both_plus(k):
    L.fmidx_query_graph_launch_on(ix._h, 8 + k % NP, sp[k % NP])
    L.fmidx_query_graph_launch_on(ix._h, 16 + k % NP, sp[k % NP])


print("K1_ZERO env:", os.environ.get("FMIDX_K1_ZERO"))
print("full step            %7.2f us" % timed(rep(0, 0)))
print("search only          %7.2f us" % timed(rep(1, 8)))
print("enumeration only     %7.2f us" % timed(rep(2, 16)))
print("memset 65.7MB only   %7.2f us" % timed(memset_only))
print("search; enumeration as 2 graphs %7.2f us" % timed(both_plus))
print("memset; search       %7.2f us" % timed(memset_plus_search))
