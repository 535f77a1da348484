"""Knob sweep of the overlapped headline step with graph replay (8 steps in flight): scan_max, bytes_occ.
    python benchmarks/knob_sweep.py"""
import ctypes as C
import os
import sys
import faulthandler
faulthandler.dump_traceback_later(200, exit=True)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fm_index_constrained_decoding_b200 import FMIndex, synth, _native  # noqa: E402
from fm_index_constrained_decoding_b200.index import DeviceGraphSteps  # noqa: E402

dev = torch.device("cuda", 0)
V, B, NP = 128256, 4096, 8
toks, beg = synth.zipf_corpus(100_000_000, seed=0, device=dev)
qsets = [synth.sample_queries(toks, B, seed=100 + i, beginnings=beg) for i in range(8)]
ix = FMIndex(device=0)
ix.initialize_from_device_tokens(toks, beg)
L = _native.lib()
streams = [torch.cuda.Stream(device=dev) for _ in range(NP)]
sp = [C.c_void_p(s.cuda_stream) for s in streams]
maxtok = max(q.numel() for q, _ in qsets)


def timed(first, steps=600):
    def fn(k):
        L.fmidx_query_graph_launch_on(ix._h, first + k % NP, sp[k % NP])
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


configs = [("default", {}), ("scan_max=1024", {"scan_max": 1024}), ("scan_max=4096", {"scan_max": 4096}), ("scan_max=65536", {"scan_max": 65536})]
keep = []
for i, (name, kw) in enumerate(configs):
    ix.tune(scan_max=16384, bytes_occ=5)
    ix.tune(**kw)
    keep.append(DeviceGraphSteps(ix, NP, B, maxtok, V, first_slot=8 * i, buffers=[(qsets[j][0], qsets[j][1]) for j in range(NP)]))
for rep in range(2):
    for i, (name, kw) in enumerate(configs):
        print("%-18s %7.2f us/step" % (name, timed(8 * i)), flush=True)
del keep
