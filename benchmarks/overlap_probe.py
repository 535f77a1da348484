"""Where does the overlapped headline step (configs[1]) spend its time?  Times the step with 1..16 batches in flight and
each phase ALONE with the same number in flight (fmidx_tune probe_phase), plus the host's issue rate.
    python benchmarks/overlap_probe.py [--tokens 100000000] [--batch 4096]"""
import argparse
import ctypes as C
import os
import sys
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fm_index_constrained_decoding_b200 import FMIndex, synth, _native  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tokens", type=int, default=100_000_000)
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--steps", type=int, default=400)
ap.add_argument("--pipes", default="1,2,4,8,12,16")
args = ap.parse_args()
dev = torch.device("cuda", 0)
V, SHIFT = 128256, 10
WORDS = (V + 31) // 32
B = args.batch
toks, beg = synth.zipf_corpus(args.tokens, seed=0, device=dev)
qsets = [synth.sample_queries(toks, B, seed=100 + i, beginnings=beg) for i in range(8)]
ix = FMIndex(device=0)
ix.initialize_from_device_tokens(toks, beg)
L = _native.lib()
NP = 16
streams = [torch.cuda.Stream(device=dev) for _ in range(NP)]
bufs = [(torch.empty(B, dtype=torch.int64, device=dev), torch.empty(B, dtype=torch.int64, device=dev),
         torch.empty((B, WORDS), dtype=torch.int32, device=dev)) for _ in range(NP)]
cargs = {}
for qi, (q, offs) in enumerate(qsets):
    for pi in range(NP):
        sp, ep, m = bufs[pi]
        cargs[(qi, pi)] = (ix._h, pi, C.c_void_p(q.data_ptr()), C.c_void_p(offs.data_ptr()), B, SHIFT, V, C.c_void_p(sp.data_ptr()),
                           C.c_void_p(ep.data_ptr()), C.c_void_p(m.data_ptr()), C.c_void_p(streams[pi].cuda_stream))


def run(npipe, steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    for s in streams[:npipe]:
        s.wait_event(e0)
    t0 = time.perf_counter()
    for k in range(steps):
        L.fmidx_query_batch(*cargs[(k % 8, k % npipe)])
    t_host = time.perf_counter() - t0
    for s in streams[:npipe]:
        e = torch.cuda.Event()
        e.record(s)
        main.wait_event(e)
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3, t_host / steps * 1e6


for phase, name in ((0, "both"), (1, "search+zero only"), (2, "enumeration only")):
    ix.tune(probe_phase=0)
    for pi in range(NP):  # valid ranges in every buffer for phase 2
        L.fmidx_query_batch(*cargs[(pi % 8, pi)])
    torch.cuda.synchronize()
    ix.tune(probe_phase=phase)
    for npipe in [int(x) for x in args.pipes.split(",")]:
        run(npipe, 50)
        us, host = run(npipe, args.steps)
        print("%-18s in flight %2d: %7.2f us/step (%6.1f M q/s)   host issue %5.2f us/step" % (name, npipe, us, B / us, host), flush=True)
ix.tune(probe_phase=0)

# the same step as ONE graph launch (fmidx_query_graph_prepare_device): is the eager step bound by the host's launch rate?
from fm_index_constrained_decoding_b200.index import DeviceGraphSteps  # noqa: E402
gs = DeviceGraphSteps(ix, NP, B, max(q.numel() for q, _ in qsets), V, first_slot=16,
                      buffers=[(qsets[i % 8][0], qsets[i % 8][1]) for i in range(NP)])
glaunch = L.fmidx_query_graph_launch_on


def run_graph(npipe, steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    for s in streams[:npipe]:
        s.wait_event(e0)
    sp = [C.c_void_p(s.cuda_stream) for s in streams]
    t0 = time.perf_counter()
    for k in range(steps):
        glaunch(ix._h, 16 + k % npipe, sp[k % npipe])
    t_host = time.perf_counter() - t0
    for s in streams[:npipe]:
        e = torch.cuda.Event()
        e.record(s)
        main.wait_event(e)
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3, t_host / steps * 1e6


for npipe in [int(x) for x in args.pipes.split(",")]:
    run_graph(npipe, 50)
    us, host = run_graph(npipe, args.steps)
    print("%-18s in flight %2d: %7.2f us/step (%6.1f M q/s)   host issue %5.2f us/step" % ("graph replay", npipe, us, B / us, host), flush=True)
# parity of the replayed step with the eager one
for i in range(NP):
    spb, epb, mb = bufs[i]
    L.fmidx_query_batch(*cargs[(i % 8, i)])
torch.cuda.synchronize()
ok = all(torch.equal(gs.sp[i], bufs[i][0]) and torch.equal(gs.ep[i], bufs[i][1]) and torch.equal(gs.mask[i], bufs[i][2]) for i in range(NP))
print("graph replay == eager:", ok)
