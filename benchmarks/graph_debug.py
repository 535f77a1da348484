import faulthandler, sys, os, ctypes as C
faulthandler.dump_traceback_later(45, exit=True)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fm_index_constrained_decoding_b200 import FMIndex, synth, _native
from fm_index_constrained_decoding_b200.index import DeviceGraphSteps
dev = torch.device("cuda", 0)
V, B = 128256, 4096
toks, beg = synth.zipf_corpus(2_000_000, seed=0, device=dev)
qsets = [synth.sample_queries(toks, B, seed=100 + i, beginnings=beg) for i in range(4)]
ix = FMIndex(device=0)
ix.initialize_from_device_tokens(toks, beg)
print("index built", flush=True)
out = ix.query_batch(qsets[0][0], qsets[0][1], V)
torch.cuda.synchronize()
print("eager ok", flush=True)
gs = DeviceGraphSteps(ix, 4, B, max(q.numel() for q, _ in qsets), V, first_slot=16, buffers=[(q, o) for q, o in qsets])
torch.cuda.synchronize()
print("captured", flush=True)
for rep in range(3):
    for i in range(4):
        gs.launch(i)
    for i in range(4):
        gs.wait(i)
    print("replayed", rep, flush=True)
for i in range(4):
    sp, ep, m = ix.query_batch(qsets[i][0], qsets[i][1], V)
    torch.cuda.synchronize()
    print(i, torch.equal(sp, gs.sp[i]), torch.equal(ep, gs.ep[i]), torch.equal(m, gs.mask[i]), flush=True)
st = torch.cuda.Stream()
for rep in range(2):
    for i in range(4):
        gs.launch(i, st)
st.synchronize()
print("external stream ok", flush=True)
