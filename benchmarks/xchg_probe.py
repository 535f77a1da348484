"""Probe (1 GPU, torchrun nproc=1): where does the sharded step lose time?"""
import os, sys, ctypes as C, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fm_index_constrained_decoding_b200 import FMIndex, synth, sharded, _native as N
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr); os.environ["FMIDX_DEVICE"] = str(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
toks, beg = synth.zipf_corpus(100_000_000, seed=0 if os.environ.get("SAME") else 1000 * lr, device=dev)
ix = FMIndex(device=lr); ix.initialize_from_device_tokens(toks, beg)
if lr != 0 and not os.environ.get("SAME"):
    toks, beg = synth.zipf_corpus(100_000_000, seed=0, device=dev)
_print = print
def print(*a):
    _print("rank", lr, *a, flush=True)
B, V, words = 4096, 128256, 4008
qs = [synth.sample_queries(toks, B, seed=100 + i, beginnings=beg) for i in range(8)]
rng = [ix.get_range_batch(*q) for q in qs]
L = N.lib(); s = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
def t(fn, n=200):
    for i in range(6): fn(i)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize(); return round(e0.elapsed_time(e1) / n * 1000, 1)
m1 = [torch.empty((B, words), dtype=torch.int32, device=dev) for _ in range(3)]
def plain1(i):
    sp, ep = rng[i % 8]; L.fmidx_distinct_mask_batch(ix._h, C.c_void_p(sp.data_ptr()), C.c_void_p(ep.data_ptr()), B, 10, V, C.c_void_p(m1[0].data_ptr()), None, s())
def plain3(i):
    sp, ep = rng[i % 8]; L.fmidx_distinct_mask_batch(ix._h, C.c_void_p(sp.data_ptr()), C.c_void_p(ep.data_ptr()), B, 10, V, C.c_void_p(m1[i % 3].data_ptr()), None, s())
print("K2 one torch buffer", t(plain1), "| three rotating torch buffers", t(plain3))
px = sharded.PeerMaskExchange(ix, B, V, n_buffers=3)
def local_ipc(i):
    sp, ep = rng[i % 8]; px.last_seq[i % 3] = 0; px.compute_local(i % 3, sp, ep)
print("K2 via xchg_local into 3 IPC buffers (no wait)", t(local_ipc))
def srch(i):
    ix.get_range_batch(*qs[i % 8])
print("K1", t(srch))
def serial(i):
    sp, ep = ix.get_range_batch(*qs[i % 8]); px.compute_local(i % 3, sp, ep); px.exchange(i % 3)
print("serial step K1+local+exchange (one stream)", t(serial))
comm = torch.cuda.Stream(device=dev); evk = [torch.cuda.Event() for _ in range(3)]; evc = [torch.cuda.Event() for _ in range(3)]
main = torch.cuda.current_stream()
def overl(i):
    bi = i % 3
    main.wait_event(evc[(i + 1) % 3])
    sp, ep = ix.get_range_batch(*qs[i % 8]); px.compute_local(bi, sp, ep); evk[bi].record(main)
    with torch.cuda.stream(comm):
        comm.wait_event(evk[bi]); px.exchange(bi); evc[bi].record(comm)
print("overlapped step", t(overl))
def host_only(i):
    pass
import time
t0 = time.perf_counter()
for i in range(2000): overl(i)
torch.cuda.synchronize(); print("overlapped, wall us/step", round((time.perf_counter() - t0) / 2000 * 1e6, 1))
px.close(); dist.destroy_process_group()
