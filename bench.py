#!/usr/bin/env python
"""bench.py -- headline benchmark of the FM-index query path (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--tokens 100000000] [--batch 4096]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One STEP = one batch of `--batch` ragged prefixes (length 1..32, substrings of corpus documents)
pushed through the hot path on a device-resident index of a synthetic `--tokens`-token corpus
(vocab 128,256): K1 batched backward search -> [sp,ep), then K2 distinct-next-token enumeration
-> one 128,256-bit vocab mask per prefix.  `value` = prefixes/s with inputs resident in HBM,
consecutive (independent) steps kept in flight on 4 streams through fmidx_query_batch; the roofline's
per-kernel durations come from a serial pass of the same steps.  `e2e` = the same through the
host-buffer C-ABI call (fmidx_query_submit / fmidx_query_wait: H2D of the prefixes, D2H of ranges and the
per-prefix distinct counts inside the timed region).  Multi-GPU (N>1): the corpus is sharded by
document, one 100M-token shard per GPU (weak scaling), prefixes are replicated, per-shard masks
are OR-reduced by the peer-memory pull exchange (own kernels over CUDA-IPC buffers; NCCL all-to-all
only as fallback) so that rank r ends up owning the final masks of its 1/N of the prefixes.

--impl reference times the CPU oracle (oracle/fm_oracle.c: sdsl-semantics restatement; the real
reference needs sdsl-lite and cannot be built here) on all host threads on the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VOCAB, EOS, SHIFT = 128256, 128009, 10
BYTES_PER_SEARCH_STEP = 1088      # 2 boundaries x 17 levels x 32-B rank cell
MASK_BYTES = 16032                # one 128,256-bit vocab mask
N_QUERY_SETS = 8                  # distinct prefix batches rotated through the timed steps
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` capture of this very
# command (profiles/r1_v3_ncu_full_summary.csv; 100M tokens, batch 4096).  k_mask = k_mask_top + k_mask_expand +
# k_mask_bytes: 2.32 + 7.36 + 27.42 MB read; the 65.7 MB of mask rows are written once (zeroed in k_mask_top, bits
# ORed in by L2 atomics) -- ncu sees 10.1 MB of them leave L2 inside the kernels' own windows, the rest drains later.
NCU_TRAFFIC = {"k_mask": 2.32e6 + 7.36e6 + 27.42e6 + 65.7e6, "k_search": 11.44e6}


_JSON_OUT = None


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (pynvml)."""

    def __init__(self, dev):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(dev)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv = None
            log("clock sampling unavailable:", e)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "display_clock_setting": 0x100}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def build_workload(args, rank, device):
    """Synthetic corpus shard + query batches for this rank."""
    from fm_index_constrained_decoding_b200 import synth
    # every rank owns a different 100M-token shard (seed differs); queries come from shard 0's
    # generator on every rank (replicated prefixes) so rank 0's positives are real matches there
    toks, beg = synth.zipf_corpus(args.tokens, seed=1000 * rank, device=device)
    q_src, q_beg = (toks, beg) if rank == 0 else synth.zipf_corpus(args.tokens, seed=0, device=device)
    qsets = [synth.sample_queries(q_src, args.batch, seed=100 + i, beginnings=q_beg) for i in range(N_QUERY_SETS)]
    del q_src
    return toks, beg, qsets


def reversed_symbols(toks, beg):
    """The text handed to FMIndex::initialize: per-document reversal + SHIFT (index.py:52,62)."""
    n = toks.numel()
    pos = torch.arange(n, device=toks.device, dtype=torch.int64)
    doc = torch.searchsorted(beg, pos, right=True) - 1
    src = beg[doc] + beg[doc + 1] - 1 - pos
    return (toks.long()[src] + SHIFT).to(torch.int32)


def oracle_from_gpu_sa(toks, beg, device_index, ix=None):
    """CPU oracle over the SAME corpus.  Suffix-sorting 100M tokens on one host core would take
    longer than the whole bench, so the oracle's structures are built from the suffix array the
    GPU builder produced (the oracle's own SA-IS is cross-checked against it in tests/)."""
    from fm_index_constrained_decoding_b200 import _native, fm_index
    from oracle.oracle import NativeOracle
    sym = reversed_symbols(toks, beg)
    tmp = ix if (ix is not None and ix.info()["sa_dens"] == 1) else None
    if tmp is None:
        tmp = fm_index.FMIndex(device=device_index, sa_dens=1, isa_dens=1 << 20, keep_text=False)
        _native.check(_native.lib().fmidx_initialize_device(tmp._h, C.c_void_p(sym.data_ptr()), sym.numel(),
                                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    sa = np.empty(sym.numel() + 1, dtype=np.uint32)
    _native.check(_native.lib().fmidx_debug_copy_sa(tmp._h, sa.ctypes.data_as(C.c_void_p), sa.size))
    del tmp
    text = np.concatenate([sym.cpu().numpy().astype(np.uint32), np.zeros(1, np.uint32)])
    o = NativeOracle()
    o.initialize_from_sa(text, sa)
    return o


def cpu_step(o, toks_np, offs_np, nthreads):
    t0 = time.perf_counter()
    o.query_batch(toks_np, offs_np, shift=SHIFT, want_mask=True, vocab=VOCAB, nthreads=nthreads)
    return time.perf_counter() - t0


def run_reference(args):
    """--impl reference: the CPU implementation of the path on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nthreads = os.cpu_count() or 1
    from fm_index_constrained_decoding_b200 import synth
    note = ""
    if torch.cuda.is_available():
        device = torch.device("cuda", 0)
        torch.cuda.set_device(device)
        os.environ.setdefault("FMIDX_DEVICE", "0")
        toks, beg, qsets = build_workload(args, 0, device)
        t0 = time.time()
        o = oracle_from_gpu_sa(toks, beg, 0)
        log("oracle built in %.1fs" % (time.time() - t0))
    else:
        # no GPU: the oracle suffix-sorts itself (SA-IS, one core), on a corpus cut to 10M tokens
        from oracle.oracle import NativeOracle
        args.tokens = min(args.tokens, 10_000_000)
        note = " [no GPU: corpus cut to %d tokens, oracle's own SA-IS]" % args.tokens
        toks, beg = synth.zipf_corpus(args.tokens, seed=0, device="cpu")
        qsets = [synth.sample_queries(toks, args.batch, seed=100 + i, beginnings=beg) for i in range(N_QUERY_SETS)]
        o = NativeOracle()
        o.initialize(reversed_symbols(toks, beg).numpy().astype(np.uint64))
    # bounded sample: each step = the first `cpu_batch` prefixes of the step's batch, sized from a probe so
    # that warm-up + steps stay within ~2 minutes of host time
    qn, on = qsets[0][0].cpu().numpy(), qsets[0][1].cpu().numpy()
    t_probe = cpu_step(o, qn[:on[256]], on[:257], nthreads) / 256
    budget = 120.0 / max(1, args.steps + args.warmup)
    cpu_batch = int(max(64, min(args.batch, args.cpu_batch, budget / max(t_probe, 1e-9))))
    host = []
    for q, offs in qsets:
        qn, on = q.cpu().numpy(), offs.cpu().numpy()
        host.append((qn[:on[cpu_batch]].copy(), on[:cpu_batch + 1].copy()))
    for w in range(args.warmup):
        cpu_step(o, *host[w % len(host)], nthreads)
    t = 0.0
    for k in range(args.steps):
        t += cpu_step(o, *host[k % len(host)], nthreads)
    value = cpu_batch * args.steps / t
    line = {
        "impl": "reference", "metric": "prefix_queries_per_sec", "value": value, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(args, cpu_batch),
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": nthreads, "kind": "port",
                         "sample": "%d of the %d prefixes of each step (search + distinct-token mask), oracle/fm_oracle.c "
                                   "(sdsl-semantics restatement, not sdsl-lite), pthread pool%s" % (cpu_batch, args.batch, note)},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(args, batch):
    return {"workload": "configs[1]: synthetic %dM-token corpus per GPU, vocab 128256, batch %d prefixes of length 1-32: "
                        "backward search + distinct next-token mask" % (args.tokens // 1_000_000, batch),
            "tokens_per_gpu": args.tokens, "batch": batch, "prefix_len": "U{1..32} clipped at document end",
            "corpus": "Zipf(1.1) token ids, docs U[64,1024] ending in EOS 128009, 1% planted repeats",
            "l2": "index (>=5x L2) larger than L2; %d distinct prefix batches rotated" % N_QUERY_SETS,
            "sharding": "by document, one shard per GPU, prefixes replicated; per step every rank publishes its occupancy words "
                        "+ counts into the peers' inboxes and pulls + ORs the occupied 4096-token slices of the rows it owns "
                        "straight from peer memory over NVLink (own kernels, no NCCL in the loop)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--tokens", type=int, default=100_000_000)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--cpu-batch", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layout", default="resident", choices=["resident", "sampled"])
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    # the one JSON line must be the only thing on stdout: NCCL / library banners that print to fd 1
    # are sent to stderr, the line itself goes to the saved descriptor
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the FM-index path has no CPU fallback")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    os.environ["FMIDX_DEVICE"] = str(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    from fm_index_constrained_decoding_b200 import FMIndex, _native
    L = _native.lib()
    t0 = time.time()
    toks, beg, qsets = build_workload(args, rank, device)
    torch.cuda.synchronize()
    t_gen = time.time() - t0
    t0 = time.time()
    ix = FMIndex(device=local, layout=args.layout)
    ix.initialize_from_device_tokens(toks, beg)
    torch.cuda.synchronize()
    t_build = time.time() - t0
    info = ix.info()
    log("rank %d: corpus %.1fs, index build %.1fs, %d levels, %.0f MB on device" %
        (rank, t_gen, t_build, info["levels"], info["device_bytes"] / 1e6))

    B = args.batch
    words = (VOCAB + 31) // 32
    stream = torch.cuda.current_stream()
    sptr = C.c_void_p(stream.cuda_stream)
    # N > 1: NBUF (4) buffer sets.  The cross-shard exchange of step k runs on its own stream while the kernels of
    # step k+1 fill the next set; the bitmaps live in peer-visible memory and the exchange PULLS only the occupied
    # 4096-token slices of the rows this rank owns (sharded.PeerMaskExchange), instead of an all-to-all of 16 KB rows.
    NBUF = int(os.environ.get("BENCH_NBUF", "4")) if world > 1 else 1
    bufs = [(torch.empty(B, dtype=torch.int64, device=device), torch.empty(B, dtype=torch.int64, device=device),
             torch.empty((B, words), dtype=torch.int32, device=device) if world == 1 else None) for _ in range(NBUF)]
    sp, ep, mask = bufs[0]
    own = B // world  # prefixes whose final mask this rank owns
    from fm_index_constrained_decoding_b200 import sharded
    result = {}
    comm_stream = torch.cuda.Stream(device=device) if world > 1 else None
    ev_kernels = [torch.cuda.Event() for _ in bufs]
    ev_comm = [torch.cuda.Event() for _ in bufs]
    px = None
    if world > 1:
        # peer-memory exchange (CUDA IPC); if any rank cannot set it up, every rank falls back to the NCCL
        # all-to-all + fmidx_mask_or path
        try:
            px = sharded.PeerMaskExchange(ix, B, VOCAB, n_buffers=NBUF)
            ok = torch.ones(1, device=device)
        except Exception as e:  # noqa: BLE001
            log("rank %d: peer-memory exchange unavailable (%r), using the NCCL all-to-all path" % (rank, e))
            ok = torch.zeros(1, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok[0]) == 0.0:
            px = None
            bufs = [(b0, b1, torch.empty((B, words), dtype=torch.int32, device=device)) for b0, b1, _ in bufs]
        mask = px.local_views[0] if px is not None else bufs[0][2]
    cnt_bufs = [torch.empty(B, dtype=torch.int64, device=device) for _ in bufs]

    # N > 1: every buffer set has its own kernel stream and scratch slot, so the kernels of up to NBUF consecutive steps
    # overlap each other as well as the exchanges (which stay ordered on the exchange stream)
    kstreams = [torch.cuda.Stream(device=device) for _ in bufs] if world > 1 else [stream]

    def step_full(k, ev=None, q_off=None, serial=False):
        q, offs = q_off if q_off is not None else qsets[k % N_QUERY_SETS]
        bi = k % NBUF
        bsp, bep, bmask = bufs[bi]
        ks = stream if serial or world == 1 else kstreams[bi]
        kptr = C.c_void_p(ks.cuda_stream)
        if world > 1:
            # set bi was last used by step k-NBUF: this rank's exchange of that step has to be complete (the peers'
            # pulls of it are awaited on the GPU through the ack flags inside compute_local)
            ks.wait_event(ev_comm[bi])
            if q_off is not None and ks is not stream:  # inputs were copied on the caller's stream
                e_in = torch.cuda.Event()
                e_in.record(stream)
                ks.wait_event(e_in)
        if ev:
            ev[0].record(ks)
        _native.check(L.fmidx_search_batch(ix._h, C.c_void_p(q.data_ptr()), C.c_void_p(offs.data_ptr()), B, SHIFT,
                                           C.c_void_p(bsp.data_ptr()), C.c_void_p(bep.data_ptr()), kptr))
        if ev:
            ev[1].record(ks)
        if px is not None:
            px.compute_local(bi, bsp, bep, SHIFT, ws_slot=bi, stream=ks)
        else:
            _native.check(L.fmidx_distinct_mask_batch(ix._h, C.c_void_p(bsp.data_ptr()), C.c_void_p(bep.data_ptr()), B, SHIFT,
                                                      VOCAB, C.c_void_p(bmask.data_ptr()), None, kptr))
        if ev:
            ev[2].record(ks)
        if world > 1:
            # rank r ends up owning the final masks of prefixes [r*B/W, (r+1)*B/W); total occurrence counts are summed
            if px is not None:
                # publish / wait / pull+OR / counts / ack: one host call, on the set's own stream right behind its
                # kernels (every set has its own flags, so the exchanges of the sets in flight overlap each other)
                result["mask"] = px.exchange(bi, stream=ks)
                result["count"] = px.total_counts(bi)
                ev_comm[bi].record(ks)
            else:
                ev_kernels[bi].record(ks)
                comm_stream.wait_event(ev_kernels[bi])
                with torch.cuda.stream(comm_stream):
                    result["mask"] = sharded.or_reduce_scatter(bmask, index=ix)[1]
                    result["count"] = sharded.sum_counts(bsp, bep)
                ev_comm[bi].record(comm_stream)
        else:
            result["mask"] = bmask

    def drain():
        if world > 1:
            for e in ev_comm:
                stream.wait_event(e)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N = 1: consecutive steps are independent batches, so NPIPE (4) of them are kept in flight on as many streams through
    # fmidx_query_batch (one call = search + masks of one batch; each stream has its own output buffers and scratch
    # slot).  The latency-bound search of one batch then overlaps the issue-bound enumeration of the other.  The
    # per-kernel durations for the roofline come from a SERIAL pass of the same steps below (events on one stream).
    NPIPE = int(os.environ.get("BENCH_NPIPE", "4"))
    pstreams = [torch.cuda.Stream(device=device) for _ in range(NPIPE)] if world == 1 else []
    pbufs = [bufs[0]] + [(torch.empty_like(sp), torch.empty_like(ep), torch.empty_like(mask)) for _ in range(NPIPE - 1)] \
        if world == 1 else []

    # ctypes argument tuples are built once: the host has ~45 us per step at four batches in flight
    pipe_args = {}
    for qi, (q, offs) in enumerate(qsets):
        for pi in range(len(pstreams)):
            psp, pep, pmask = pbufs[pi]
            pipe_args[(qi, pi)] = (ix._h, pi, C.c_void_p(q.data_ptr()), C.c_void_p(offs.data_ptr()), B, SHIFT, VOCAB,
                                   C.c_void_p(psp.data_ptr()), C.c_void_p(pep.data_ptr()), C.c_void_p(pmask.data_ptr()),
                                   C.c_void_p(pstreams[pi].cuda_stream))
    query_batch = L.fmidx_query_batch

    def step_pipe(k):
        if query_batch(*pipe_args[(k % N_QUERY_SETS, k % NPIPE)]) != 0:
            _native.check(-1)

    def fork():
        e = torch.cuda.Event()
        e.record(stream)
        for ps in pstreams + (kstreams if world > 1 else []):
            ps.wait_event(e)

    def join():
        for ps in pstreams + (kstreams if world > 1 else []):
            e = torch.cuda.Event()
            e.record(ps)
            stream.wait_event(e)

    step = step_pipe if world == 1 else step_full
    fork()
    for w in range(args.warmup):
        step(w)
    join()
    barrier()
    launches0 = L.fmidx_launch_count()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    e_start.record(stream)
    fork()
    for k in range(args.steps):
        step(k)
    join()
    drain()  # the timed region ends when the last exchange has completed
    e_end.record(stream)
    barrier()
    total_ms = e_start.elapsed_time(e_end)
    launches = L.fmidx_launch_count() - launches0
    # serial pass with events around each kernel group (N = 1: this is where the roofline's launch durations come
    # from; N > 1: the timed region itself is serial on the kernel stream, so it is simply repeated)
    n_ev = min(args.steps, 500)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n_ev)]
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    es0.record(stream)
    for k in range(n_ev):
        step_full(k, evs[k], serial=True)
    drain()
    es1.record(stream)
    barrier()
    serial_ms_per_step = es0.elapsed_time(es1) / n_ev
    if sampler and total_ms < 150:  # region too short for NVML's sampling period: keep the same load running
        t_end = time.time() + 0.6
        k = 0
        scratch_mask = mask if world == 1 else torch.empty((B, words), dtype=torch.int32, device=device)
        while time.time() < t_end:  # local kernels only: the other ranks are not in this loop
            q, offs = qsets[k % N_QUERY_SETS]
            L.fmidx_search_batch(ix._h, C.c_void_p(q.data_ptr()), C.c_void_p(offs.data_ptr()), B, SHIFT,
                                 C.c_void_p(sp.data_ptr()), C.c_void_p(ep.data_ptr()), sptr)
            L.fmidx_distinct_mask_batch(ix._h, C.c_void_p(sp.data_ptr()), C.c_void_p(ep.data_ptr()), B, SHIFT, VOCAB,
                                        C.c_void_p(scratch_mask.data_ptr()), None, sptr)
            k += 1
            if k % 50 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    if clocks is not None and total_ms < 150:
        clocks["note"] = "timed region %.0f ms is shorter than NVML's sampling period; sampled while the same step kept running" % total_ms
    t_search = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    t_mask = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    if world > 1:
        tt = torch.tensor([total_ms, t_search, t_mask], device=device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms, t_search, t_mask = tt.tolist()

    # ---- algorithmic bytes of one step, from the actual results ------------------------
    # "visited nodes per level, 2 cells each" for the enumeration and "2 boundaries x levels x
    # 1 rank cell" for a search step, with the cell size of the layout in use ("state which").  Ours
    # (DESIGN.md section 3): 3 nibble levels of 128-B cells + 1 byte level of 1,536-B cells, so
    #   K2: 2 x 128 B per visited nibble node + 2 x 1,536 B per visited byte-level node + the 16,032-B mask row,
    #       visited nodes counted exactly from the result masks (distinct symbol prefixes per level);
    #   K1: 2 x (3 x 128 + 1,536) = 3,840 B per executed step.
    # The same sums on the REFERENCE's layout (17 binary levels, 32-B cells: 64 B per visited node, 1,088 B
    # per search step) are reported beside them as *_reference_layout.
    search_steps, enum_nodes_bin, nib_nodes, byte_nodes, distinct_total = 0, 0, 0, 0, 0
    for q, offs in qsets:
        L.fmidx_search_batch(ix._h, C.c_void_p(q.data_ptr()), C.c_void_p(offs.data_ptr()), B, SHIFT,
                             C.c_void_p(sp.data_ptr()), C.c_void_p(ep.data_ptr()), sptr)
        L.fmidx_distinct_mask_batch(ix._h, C.c_void_p(sp.data_ptr()), C.c_void_p(ep.data_ptr()), B, SHIFT, VOCAB,
                                    C.c_void_p(mask.data_ptr()), None, sptr)
        torch.cuda.synchronize()
        search_steps += int((offs[1:] - offs[:-1]).sum())  # positive prefixes: every step executes
        sh = torch.arange(32, device=device, dtype=torch.int32)
        for r0 in range(0, B, 256):
            bits = ((mask[r0:r0 + 256].unsqueeze(-1) >> sh) & 1).bool().view(-1, words * 32)[:, :VOCAB]
            symb = torch.zeros((bits.shape[0], 1 << 17), dtype=torch.bool, device=device)
            symb[:, SHIFT:SHIFT + VOCAB] = bits  # symbol axis = token + SHIFT
            distinct_total += int(bits.sum())
            lvl = symb
            for l in range(17, 0, -1):          # nodes at binary depth l-1 = distinct (l-1)-bit prefixes
                lvl = lvl.view(lvl.shape[0], -1, 2).any(-1)
                cnt_l = int(lvl.sum())
                enum_nodes_bin += cnt_l
                if l - 1 == 9:                  # byte-level nodes: symbol >> 8
                    byte_nodes += cnt_l
                elif l - 1 in (5, 1, 0):        # nibble-level nodes: symbol >> 12, >> 16, root
                    nib_nodes += cnt_l
    search_steps /= N_QUERY_SETS
    enum_nodes_bin /= N_QUERY_SETS
    nib_nodes /= N_QUERY_SETS
    byte_nodes /= N_QUERY_SETS
    distinct_total /= N_QUERY_SETS
    bytes_search = 3840 * search_steps
    bytes_search_ref = BYTES_PER_SEARCH_STEP * search_steps
    bytes_mask = 2 * 128 * nib_nodes + 2 * 1536 * byte_nodes + MASK_BYTES * B
    bytes_mask_ref = 2 * 32 * enum_nodes_bin + MASK_BYTES * B
    peak, peak_src = measured_peaks()
    if t_mask >= t_search:
        dom = {"kernel": "k_mask (k_mask_top + k_mask_expand + k_mask_bytes: distinct-token enumeration -> vocab bitmap)", "bound": "hbm",
               "achieved": bytes_mask / (t_mask * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
               "traffic": NCU_TRAFFIC["k_mask"] if args.tokens == 100_000_000 and B == 4096 else None,
               "algorithmic_bytes_per_launch": bytes_mask, "avg_launch_ms": t_mask,
               "per_unit": "2 x 128 B per visited nibble node + 2 x 1,536 B per visited byte-level node + 16,032 B mask per prefix",
               "algorithmic_bytes_reference_layout": bytes_mask_ref,
               "achieved_reference_layout": bytes_mask_ref / (t_mask * 1e-3) / 1e9,
               "frac_reference_layout": bytes_mask_ref / (t_mask * 1e-3) / 1e9 / peak,
               "per_unit_reference_layout": "64 B per visited node of the 17-level binary tree + 16,032 B mask per prefix",
               "distinct_symbols_per_launch": distinct_total,
               "note": "issue/latency bound (ncu: k_mask_bytes 66% issue-active, 13% DRAM; k_mask_top / k_mask_expand latency-bound): the fraction measures distance from the HBM "
                       "bound, the reference-layout figure shows how much less this layout has to move"}
    else:
        dom = {"kernel": "k_search (batched backward search)", "bound": "hbm",
               "achieved": bytes_search / (t_search * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
               "traffic": NCU_TRAFFIC["k_search"] if args.tokens == 100_000_000 and B == 4096 else None,
               "algorithmic_bytes_per_launch": bytes_search, "avg_launch_ms": t_search,
               "per_unit": "3,840 B per backward-search step of the prefix (2 x (3 x 128 + 1,536) B)",
               "algorithmic_bytes_reference_layout": bytes_search_ref,
               "achieved_reference_layout": bytes_search_ref / (t_search * 1e-3) / 1e9,
               "frac_reference_layout": bytes_search_ref / (t_search * 1e-3) / 1e9 / peak}
    dom["frac"] = dom["achieved"] / peak
    # the rank probes are random accesses, so also measure what this GPU sustains for INDEPENDENT random
    # 32-B units / aligned 128-B lines over the index's own arrays (fmidx_gather_peak) and relate both
    # kernels' algorithmic bytes (own layout) to the 128-B figure
    try:
        gp = {}
        for unit in (32, 128):
            v = C.c_double()
            _native.check(L.fmidx_gather_peak(ix._h, unit, 1 << 25, C.byref(v)))
            gp["%dB_GBps" % unit] = v.value
        gp["note"] = ("independent random loads over the index's own arrays (18 B/token); fractions above 1 mean the kernel "
                      "touches less than its full-cell algorithmic figure (<=16-row search shortcut, partially read byte "
                      "cells, L2 hits on the top levels)")
        dom["random_gather_peak"] = gp
        dom["frac_of_random_gather_128B"] = dom["achieved"] / gp["128B_GBps"]
        dom["k_search_achieved_GBps"] = bytes_search / (t_search * 1e-3) / 1e9
        dom["k_search_frac_of_random_gather_128B"] = dom["k_search_achieved_GBps"] / gp["128B_GBps"]
    except Exception as e:  # noqa: BLE001 -- a measurement aid must not take the line down
        dom["random_gather_peak"] = {"error": repr(e)}
    dom["peak_source"] = peak_src
    dom["other_kernel_ms"] = {"k_search": t_search, "k_mask": t_mask}

    # ---- e2e: host buffers through the C ABI, copies inside the timed region -----------------------------
    e2e = None
    if rank == 0 or world > 1:
        hq = [(q.cpu().pin_memory(), offs.cpu().pin_memory()) for q, offs in qsets]
        hsp = torch.empty(B, dtype=torch.int64).pin_memory()
        hep = torch.empty(B, dtype=torch.int64).pin_memory()
        hnd = torch.empty(B, dtype=torch.int32).pin_memory()

        dq = [(torch.empty_like(q, device=device), torch.empty_like(offs, device=device)) for q, offs in hq[:1]]
        n32 = torch.empty(own, dtype=torch.int32, device=device)
        hcnt = torch.empty(B, dtype=torch.int64).pin_memory()

        NSLOT = int(os.environ.get("BENCH_NSLOT", "6"))  # batches in flight through the pipelined host API
        hres = [(torch.empty(B, dtype=torch.int64).pin_memory(), torch.empty(B, dtype=torch.int64).pin_memory(),
                 torch.empty(B, dtype=torch.int32).pin_memory()) for _ in range(NSLOT)]

        submit_args = {}
        for qi, (q, offs) in enumerate(hq):
            for slot in range(NSLOT):
                rsp, rep, rnd = hres[slot]
                submit_args[(qi, slot)] = (ix._h, slot, C.c_void_p(q.data_ptr()), C.c_void_p(offs.data_ptr()), B, SHIFT, VOCAB,
                                           C.c_void_p(rsp.data_ptr()), C.c_void_p(rep.data_ptr()), C.c_void_p(rnd.data_ptr()))
        query_wait, query_submit = L.fmidx_query_wait, L.fmidx_query_submit

        def e2e_step(k):
            q, offs = hq[k % N_QUERY_SETS]
            if world == 1:
                # fmidx_query_submit/wait: the H2D copy of this step's prefixes, its kernels and the D2H read of its
                # results are all inside the timed region; the wait for a slot happens when the slot is reused, so
                # the copies of one step overlap the kernels of the others
                slot = k % NSLOT
                if query_wait(ix._h, slot) != 0 or query_submit(*submit_args[(k % N_QUERY_SETS, slot)]) != 0:
                    _native.check(-1)
                return
            # N > 1: pinned host prefixes -> device, search + masks + cross-shard exchange, results -> host
            if px is not None:
                # pipelined like the N = 1 path: the step's copies, kernels, exchange and result reads are all queued on
                # the stream of its buffer set; the host only waits for a set's previous results when it reuses the set
                bi = k % NBUF
                ks = kstreams[bi]
                set_done[bi].synchronize()
                dtok, doff = set_in[bi]
                rsp, rep, rcnt, rnd = set_out[bi]
                n = q.numel()
                with torch.cuda.stream(ks):
                    dtok[:n].copy_(q, non_blocking=True)
                    doff.copy_(offs, non_blocking=True)
                step_full(k, q_off=(dtok[:n], doff))
                _native.check(L.fmidx_mask_popcount(ix._h, C.c_void_p(result["mask"].data_ptr()), own, VOCAB,
                                                    C.c_void_p(set_n32[bi].data_ptr()), C.c_void_p(ks.cuda_stream)))
                with torch.cuda.stream(ks):
                    rsp.copy_(bufs[bi][0], non_blocking=True)
                    rep.copy_(bufs[bi][1], non_blocking=True)
                    rcnt.copy_(result["count"], non_blocking=True)
                    rnd[:own].copy_(set_n32[bi], non_blocking=True)
                    set_done[bi].record(ks)
                return
            dtok = torch.empty(q.numel(), dtype=torch.int32, device=device)
            dtok.copy_(q, non_blocking=True)
            doff = dq[0][1]
            doff.copy_(offs, non_blocking=True)
            step_full(k, q_off=(dtok, doff))
            drain()
            _native.check(L.fmidx_mask_popcount(ix._h, C.c_void_p(result["mask"].data_ptr()), own, VOCAB, C.c_void_p(n32.data_ptr()), sptr))
            hsp.copy_(bufs[k % len(bufs)][0], non_blocking=True)
            hep.copy_(bufs[k % len(bufs)][1], non_blocking=True)
            hcnt.copy_(result["count"], non_blocking=True)
            hnd[:own].copy_(n32, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        if world > 1 and px is not None:
            max_tok = max(q.numel() for q, _ in hq)
            set_in = [(torch.empty(max_tok, dtype=torch.int32, device=device), torch.empty(B + 1, dtype=torch.int64, device=device))
                      for _ in range(NBUF)]
            set_out = [(torch.empty(B, dtype=torch.int64).pin_memory(), torch.empty(B, dtype=torch.int64).pin_memory(),
                        torch.empty(B, dtype=torch.int64).pin_memory(), torch.empty(B, dtype=torch.int32).pin_memory())
                       for _ in range(NBUF)]
            set_n32 = [torch.empty(own, dtype=torch.int32, device=device) for _ in range(NBUF)]
            set_done = [torch.cuda.Event() for _ in range(NBUF)]
            torch.cuda.synchronize()

        for w in range(args.warmup):
            e2e_step(w)
        barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            e2e_step(k)
        if world == 1:
            for slot in range(NSLOT):
                _native.check(L.fmidx_query_wait(ix._h, slot))  # every step's results are on the host
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], device=device, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt[0])
        h2d = int(np.mean([q.numel() * 4 + offs.numel() * 8 for q, offs in hq]))
        e2e = {"value": B * args.steps / dt, "unit": "queries/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": B * (8 + 8 + 4) if world == 1 else B * 24 + own * 4, "ms_per_step": 1e3 * dt / args.steps,
               "api": "fmidx_query_submit/fmidx_query_wait, %d batches in flight (pinned host prefixes in; ranges + distinct-token "
                      "counts out, every step; masks stay in HBM)" % NSLOT
                      if world == 1 else "pinned host prefixes -> H2D -> fmidx_search_batch + fmidx_xchg_local + fmidx_xchg_round (peer-memory "
                      "exchange) -> D2H of ranges, total counts and distinct counts of the owned prefixes, every step; %d steps in "
                      "flight, one per buffer set" % NBUF}

    # ---- CPU baseline beside the GPU numbers (rank 0, N=1) ---------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            nthreads = os.cpu_count() or 1
            t0 = time.time()
            o = oracle_from_gpu_sa(toks, beg, local, ix)
            log("oracle built in %.1fs" % (time.time() - t0))
            q, offs = qsets[0]
            qn, on = q.cpu().numpy(), offs.cpu().numpy()
            cb = min(B, args.cpu_batch)
            qn, on = qn[:on[cb]], on[:cb + 1]
            # parity spot check of the product against the oracle on this very batch
            osp, oep, omask, _ = o.query_batch(qn, on, shift=SHIFT, want_mask=True, vocab=VOCAB, nthreads=nthreads)
            step_full(0)
            torch.cuda.synchronize()
            ok = bool((sp[:cb].cpu().numpy().astype(np.uint64) == osp).all() and (ep[:cb].cpu().numpy().astype(np.uint64) == oep).all()
                      and (mask[:cb].cpu().numpy().view(np.uint32) == omask).all())
            tt = min(cpu_step(o, qn, on, nthreads) for _ in range(2))
            t1 = cpu_step(o, qn[:on[256]], on[:257], 1)
            cpu = {"value": cb / tt, "unit": "queries/s", "cores": nthreads, "kind": "port",
                   "sample": "one batch of %d prefixes (search + distinct-token mask), best of 2, oracle/fm_oracle.c "
                             "(sdsl-semantics restatement, not sdsl-lite) on a pthread pool" % cb,
                   "single_thread_queries_per_sec": 256 / t1, "parity_vs_gpu_on_sample": ok}
            del o
        except Exception as e:  # the baseline must never take the bench line down
            cpu = {"value": None, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}

    if rank == 0:
        value = B * args.steps / (total_ms * 1e-3)
        line = {
            "metric": "prefix_queries_per_sec", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": workload_config(args, B),
            "masks_per_sec": value, "roofline": dom, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "pipeline": {"streams": NPIPE if world == 1 else 1, "serial_ms_per_step": serial_ms_per_step,
                         "note": "value/ms_per_step: independent batches alternate between the streams (fmidx_query_batch); "
                                 "roofline launch durations come from the serial pass (one stream, events around each kernel group)"
                                 if world == 1 else "one stream + scratch slot + flag set per buffer set (4): kernels and exchange of a set back to back on its stream; roofline "
                                 "launch durations from a serial pass"},
            "clocks": clocks, "index": {"build_s": t_build, "device_mb": info["device_bytes"] / 1e6, "levels": info["levels"],
                                        "layout": "wavelet matrix: 3 nibble levels (128-B cells) + byte level (1,536-B cells); %s (sa_dens %d, isa_dens %d, text %s)" % (args.layout, info["sa_dens"], info["isa_dens"], bool(info["keep_text"]))},
        }
        emit(line)
    if world > 1:
        if px is not None:
            px.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
