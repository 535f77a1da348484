#!/usr/bin/env python
"""bench.py -- benchmark of the FM-index query path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One STEP = one batch of 4,096 ragged prefixes (length 1..32, substrings of corpus documents) pushed through the hot
path on a device-resident index: K1 batched backward search -> [sp,ep), then K2 distinct-next-token enumeration -> one
128,256-bit vocab mask per prefix.  The K-step block the driver asks for is REPEATED until >= 250 ms of steps have
been timed; the reported time is the median block (CUDA events, max over ranks).

N = 1 (configs[1], 100M-token corpus): `value` = prefixes/s with inputs resident in HBM, four independent batches in
flight on four streams through fmidx_query_batch; `e2e` = the same through fmidx_query_submit / fmidx_query_wait with
pinned HOST buffers (H2D of the prefixes and D2H of the ranges + distinct-token counts inside the timed region).  The
same line carries `other_configs`: configs[0] batch-1 latency, the 1B-token single-GPU figure (north star), the fused
decode step (configs[3]), locate + extract (configs[4]), each with a parity flag.

N > 1, three modes measured in the same run (`modes`):
  replicas     index replicated (100M tokens each), prefixes split N ways, zero collectives  -> `value` (weak scaling)
  doc_sharded  configs[2]: a fixed 1B-token corpus split N ways by document, prefixes replicated, per-shard masks
               OR-reduced and counts summed over NVLink peer memory (fmidx_shard_step: five launches, no NCCL)
  hybrid       2 document shards x N/2 replica groups of the same 1B-token corpus (N >= 4)

--impl reference: the CPU implementation of the path (oracle/fm_oracle.c, a restatement of fm_index.cpp + sdsl-lite's
algorithms: the real reference needs sdsl-lite, which is neither in /root/reference nor in this image) on all host
threads, over a corpus generated and suffix-sorted on the CPU: that arm never loads libfmidx.so.
"""
import argparse
import ctypes as C
import json
import math
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# The step keeps up to 8 independent batches in flight on 8 streams.  With the default 8 hardware work queues streams alias
# and falsely serialise (round 1: "eight streams collapse to 36 M q/s"); 32 queues remove that.  Must be set before CUDA starts.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

VOCAB, EOS, SHIFT = 128256, 128009, 10
WORDS = (VOCAB + 31) // 32
MASK_BYTES = WORDS * 4            # 16,032 B: one 128,256-bit vocab mask
N_QUERY_SETS = 8                  # distinct prefix batches rotated through the timed steps
MIN_TIMED_S = 0.25                # every timed quantity covers at least this much device time
SCAN_MAX = int(os.environ.get("FMIDX_SCAN_MAX", "16384"))

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


# =====================================================================================================================
# workload
# =====================================================================================================================
def reversed_symbols(toks, beg):
    """The text handed to FMIndex::initialize: per-document reversal + SHIFT (index.py:52,62)."""
    n = toks.numel()
    out = torch.empty(n, dtype=torch.int32, device=toks.device)
    CH = 1 << 27
    for a in range(0, n, CH):
        pos = torch.arange(a, min(n, a + CH), device=toks.device, dtype=torch.int64)
        doc = torch.searchsorted(beg, pos, right=True) - 1
        out[a:a + pos.numel()] = (toks[beg[doc] + beg[doc + 1] - 1 - pos] + SHIFT).to(torch.int32)
    return out


def workload_config(tokens, batch, mode="single"):
    return {"workload": "configs[1]: synthetic %dM-token corpus per GPU, vocab 128256, batch %d prefixes of length 1-32: "
                        "backward search + distinct next-token mask" % (tokens // 1_000_000, batch),
            "tokens_per_gpu": tokens, "batch": batch, "prefix_len": "U{1..32} clipped at document end",
            "corpus": "Zipf(1.1) token ids, docs U[64,1024] ending in EOS 128009, 1% planted repeats",
            "mask_rows": ("lazy clearing: the step's mask buffers belong to the bench and nothing else writes them, so only the 4096-token "
                          "slices a buffer's previous batch dirtied are zeroed (fmidx_tune lazy_clear=2); every buffer sees all 8 prefix "
                          "batches in turn; BENCH_LAZY=0 / ablation.full_row_zeroing: every 16-KB row zeroed in full each step"
                          if os.environ.get("BENCH_LAZY", "1") == "1" else "every 16-KB row zeroed in full each step"),
            "l2": "index (>= 14x L2) larger than L2; %d distinct prefix batches rotated" % N_QUERY_SETS,
            "mode": mode}


def build_index(toks, beg, local, layout="resident"):
    from fm_index_constrained_decoding_b200 import FMIndex
    t0 = time.time()
    ix = FMIndex(device=local, layout=layout)
    ix.initialize_from_device_tokens(toks, beg)
    torch.cuda.synchronize()
    return ix, time.time() - t0


# =====================================================================================================================
# timing helpers: every quantity = median over repeated K-step blocks covering >= MIN_TIMED_S
# =====================================================================================================================
class BlockTimer:
    def __init__(self, device, dist=None):
        self.device, self.dist = device, dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def run(self, block, steps, warmup_steps):
        """block(k0, n): enqueue steps k0..k0+n-1 including fork/join so that the CURRENT torch stream covers them.
        The K-step block the caller asks for is repeated back to back, WITHOUT draining the pipeline in between, in
        super-blocks of m x K steps (each >= MIN_TIMED_S / 5 of device time); at least 5 super-blocks are timed (CUDA
        events on the launching stream, per super-block the slowest rank) and the MEDIAN is reported per K steps.
        Returns (median ms per `steps` steps, number of K-step blocks timed, super-block times)."""
        stream = torch.cuda.current_stream()
        block(0, warmup_steps)
        block(0, 2 * steps)  # untimed: first use of every buffer set / stream / peer mapping
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        block(0, 4 * steps)
        e1.record(stream)
        torch.cuda.synchronize()
        probe_ms = max(e0.elapsed_time(e1) / 4, 1e-3)  # per K steps
        m = int(min(2000, max(1, math.ceil(MIN_TIMED_S * 1e3 / 5 / probe_ms))))
        if self.dist is not None:
            t = torch.tensor([m], device=self.device)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            m = int(t[0])
        n_super = 7
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_super)]
        self.barrier()
        for r, (a, b) in enumerate(evs):
            a.record(stream)
            block(r * m * steps, m * steps)
            b.record(stream)
        self.barrier()
        times = torch.tensor([a.elapsed_time(b) for a, b in evs], device=self.device, dtype=torch.float64)
        if self.dist is not None:  # per super-block: the slowest rank
            self.dist.all_reduce(times, op=self.dist.ReduceOp.MAX)
        return float(times.median()) / m, n_super * m, (times / m).tolist()


def host_blocks(fn_step, drain, steps, warmup, dist=None, device=None):
    """Host-timed variant for the e2e loops: wall clock around super-blocks of m x K steps that end with every result on
    the host (drain); median of 7 super-blocks, per K steps."""
    for w in range(warmup):
        fn_step(w)
    drain()
    t0 = time.perf_counter()
    for k in range(4 * steps):
        fn_step(k)
    drain()
    probe = max((time.perf_counter() - t0) / 4, 1e-6)
    m = int(min(2000, max(1, math.ceil(MIN_TIMED_S / 5 / probe))))
    if dist is not None:
        t = torch.tensor([m], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        m = int(t[0])
        dist.barrier()
    times = []
    for r in range(7):
        t0 = time.perf_counter()
        for k in range(m * steps):
            fn_step(r * m * steps + k)
        drain()
        times.append((time.perf_counter() - t0) / m)
    tt = torch.tensor(times, dtype=torch.float64, device=device if dist is not None else "cpu")
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return float(tt.median()), 7 * m


# =====================================================================================================================
# checkers that live on the bench side
# =====================================================================================================================
def bruteforce_sample(sym, queries, vocab=VOCAB):
    """Independent of the FM-index: occurrence counts and next-token bitmaps of `queries` (token lists) by scanning the
    text itself with torch.  sym: int32 CUDA tensor, the reversed-document concatenation T WITHOUT its sentinel (T[n] = 0).
    A prefix q occurs where T[j..j+L) == reversed(q) + SHIFT; its continuations are T[j-1] (row j = 0 contributes the
    sentinel, which index.py drops).  Returns (counts int64 [Q], masks int32 [Q, words])."""
    n = sym.numel()
    dev = sym.device
    counts = torch.zeros(len(queries), dtype=torch.int64, device=dev)
    masks = torch.zeros((len(queries), WORDS * 32), dtype=torch.bool, device=dev)
    for qi, q in enumerate(queries):
        pat = [t + SHIFT for t in reversed(q)]
        cand = torch.nonzero(sym == pat[0]).flatten()
        for k in range(1, len(pat)):
            if cand.numel() == 0:
                break
            nxt = cand + k
            ok = nxt < n
            ok &= sym[nxt.clamp(max=n - 1)] == pat[k]
            cand = cand[ok]
        counts[qi] = cand.numel()
        prev = sym[cand[cand > 0] - 1].long() - SHIFT
        prev = prev[(prev >= 0) & (prev < vocab)]
        masks[qi, prev] = True
    sh = torch.arange(32, device=dev, dtype=torch.int64)
    packed = (masks.view(len(queries), WORDS, 32).long() << sh).sum(-1)
    return counts, packed.to(torch.int32)  # wraps bit 31 into the sign, like the product's int32 view


def queries_as_lists(q, offs, n):
    qn, on = q[:int(offs[n])].cpu().tolist(), offs[:n + 1].cpu().tolist()
    return [qn[on[i]:on[i + 1]] for i in range(n)]


# =====================================================================================================================
# --impl reference: CPU only, never touches libfmidx.so
# =====================================================================================================================
def cpu_step(o, toks_np, offs_np, nthreads):
    t0 = time.perf_counter()
    o.query_batch(toks_np, offs_np, shift=SHIFT, want_mask=True, vocab=VOCAB, nthreads=nthreads)
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nthreads = os.cpu_count() or 1
    torch.set_num_threads(nthreads)
    # synth is pure torch; importing the package itself would load libfmidx.so
    import importlib.util
    spec = importlib.util.spec_from_file_location("fm_synth", os.path.join(ROOT, "fm_index_constrained_decoding_b200", "synth.py"))
    synth = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(synth)
    from oracle.oracle import NativeOracle
    t0 = time.time()
    toks, beg = synth.zipf_corpus(args.tokens, seed=0, device="cpu")
    qsets = [synth.sample_queries(toks, args.batch, seed=100 + i, beginnings=beg) for i in range(N_QUERY_SETS)]
    sym = reversed_symbols(toks, beg).numpy().astype(np.uint64)
    del toks
    log("reference arm: corpus of %d tokens generated on the CPU in %.1fs" % (args.tokens, time.time() - t0))
    t0 = time.time()
    o = NativeOracle()
    o.initialize(sym)  # the oracle's own SA-IS + structures
    del sym
    t_build = time.time() - t0
    log("reference arm: oracle (own SA-IS) built in %.1fs" % t_build)
    # bounded sample: each step = the first `cpu_batch` prefixes of the step's batch, sized from a probe so that
    # warm-up + steps stay within ~60 s of host time
    qn, on = qsets[0][0].numpy(), qsets[0][1].numpy()
    t_probe = cpu_step(o, qn[:on[256]], on[:257], nthreads) / 256
    budget = 60.0 / max(1, args.steps + args.warmup)
    cpu_batch = int(max(64, min(args.batch, args.cpu_batch, budget / max(t_probe, 1e-9))))
    host = []
    for q, offs in qsets:
        qn, on = q.numpy(), offs.numpy()
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
        "config": workload_config(args.tokens, cpu_batch, "cpu"),
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": nthreads, "kind": "port",
                         "sample": "%d of the %d prefixes of each step (search + distinct-token mask written to a 16 KB host bitmap per "
                                   "prefix), oracle/fm_oracle.c (sdsl-semantics restatement, not sdsl-lite) on a pthread pool; corpus "
                                   "generated and suffix-sorted (SA-IS, %.0f s) on the CPU, libfmidx.so is not loaded" % (cpu_batch, args.batch, t_build)},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "repo_native_libs_mapped": sorted({l.split()[-1].replace(ROOT + "/", "") for l in open("/proc/self/maps") if ROOT in l and ".so" in l}),
    }
    emit(line)


# =====================================================================================================================
# single-GPU engine of a step (used by N = 1 and by the replicas mode)
# =====================================================================================================================
class LocalEngine:
    """NPIPE (8) independent batches in flight on as many streams through fmidx_query_batch, and the host-buffer pipeline."""

    def __init__(self, ix, qsets, device, B):
        from fm_index_constrained_decoding_b200 import _native
        self.N, self.L = _native, _native.lib()
        self.ix, self.qsets, self.device, self.B = ix, qsets, device, B
        # the engine's mask buffers are its own and nothing else writes them: lazy clearing (see run_single)
        self.lazy = os.environ.get("BENCH_LAZY", "1") == "1"
        ix.tune(lazy_clear=2 if self.lazy else 1)
        self.NPIPE = int(os.environ.get("BENCH_NPIPE", "8"))
        self.pstreams = [torch.cuda.Stream(device=device) for _ in range(self.NPIPE)]
        self.pbufs = [(torch.empty(B, dtype=torch.int64, device=device), torch.empty(B, dtype=torch.int64, device=device),
                       torch.empty((B, WORDS), dtype=torch.int32, device=device)) for _ in range(self.NPIPE)]
        self.pipe_args = {}
        for qi, (q, offs) in enumerate(qsets):
            for pi in range(self.NPIPE):
                psp, pep, pmask = self.pbufs[pi]
                self.pipe_args[(qi, pi)] = (ix._h, pi, C.c_void_p(q.data_ptr()), C.c_void_p(offs.data_ptr()), B, SHIFT, VOCAB,
                                            C.c_void_p(psp.data_ptr()), C.c_void_p(pep.data_ptr()), C.c_void_p(pmask.data_ptr()),
                                            C.c_void_p(self.pstreams[pi].cuda_stream))
        self._qb = self.L.fmidx_query_batch

    def block(self, k0, n):
        stream = torch.cuda.current_stream()
        e = torch.cuda.Event()
        e.record(stream)
        for ps in self.pstreams:
            ps.wait_event(e)
        nq, npipe, qb, args = len(self.qsets), self.NPIPE, self._qb, self.pipe_args
        for k in range(k0, k0 + n):
            # the pairing (prefix batch, buffer) rotates: every mask buffer sees all 8 batches in turn (with lazy clearing a
            # buffer that always received the same batch would be cleared of exactly what is written back)
            if qb(*args[((k + k // npipe) % nq, k % npipe)]) != 0:
                self.N.check(-1)
        for ps in self.pstreams:
            e = torch.cuda.Event()
            e.record(ps)
            stream.wait_event(e)

    def serial_kernel_times(self, n_steps):
        """One stream, events around K1 and the K2 group: the per-kernel durations behind the overlapped step."""
        L, ix, B = self.L, self.ix, self.B
        sp, ep, mask = self.pbufs[0]
        stream = torch.cuda.current_stream()
        sptr = C.c_void_p(stream.cuda_stream)
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n_steps)]
        for k in range(n_steps):
            q, offs = self.qsets[k % len(self.qsets)]
            evs[k][0].record(stream)
            L.fmidx_search_batch(ix._h, C.c_void_p(q.data_ptr()), C.c_void_p(offs.data_ptr()), B, SHIFT,
                                 C.c_void_p(sp.data_ptr()), C.c_void_p(ep.data_ptr()), sptr)
            evs[k][1].record(stream)
            L.fmidx_distinct_mask_batch(ix._h, C.c_void_p(sp.data_ptr()), C.c_void_p(ep.data_ptr()), B, SHIFT, VOCAB,
                                        C.c_void_p(mask.data_ptr()), None, sptr)
            evs[k][2].record(stream)
        torch.cuda.synchronize()
        t1 = float(np.median([e[0].elapsed_time(e[1]) for e in evs]))
        t2 = float(np.median([e[1].elapsed_time(e[2]) for e in evs]))
        return t1, t2

    def setup_e2e(self):
        B = self.B
        self.NSLOT = int(os.environ.get("BENCH_NSLOT", "8"))
        self.hq = [(q.cpu().pin_memory(), offs.cpu().pin_memory()) for q, offs in self.qsets]
        self.hres = [(torch.empty(B, dtype=torch.int64).pin_memory(), torch.empty(B, dtype=torch.int64).pin_memory(),
                      torch.empty(B, dtype=torch.int32).pin_memory()) for _ in range(self.NSLOT)]
        self.submit_args = {}
        for qi, (q, offs) in enumerate(self.hq):
            for slot in range(self.NSLOT):
                rsp, rep, rnd = self.hres[slot]
                self.submit_args[(qi, slot)] = (self.ix._h, slot, C.c_void_p(q.data_ptr()), C.c_void_p(offs.data_ptr()), B, SHIFT,
                                                VOCAB, C.c_void_p(rsp.data_ptr()), C.c_void_p(rep.data_ptr()), C.c_void_p(rnd.data_ptr()))
        self.h2d = int(np.mean([q.numel() * 4 + offs.numel() * 8 for q, offs in self.hq]))
        self.d2h = B * (8 + 8 + 4)

    def setup_e2e_graph(self):
        """The captured form: slot i's pinned staging buffer holds prefix batch i % 8, exactly like the eager form rotates the
        same pinned batches; one cudaGraphLaunch per step does H2D + kernels + D2H.  BENCH_GSLOTS (16) steps in flight."""
        from fm_index_constrained_decoding_b200.index import GraphQuerySlots
        hq = [(q.cpu().numpy(), offs.cpu().numpy()) for q, offs in self.qsets]
        self.g_maxtok = max(q.size for q, _ in hq)
        self.GSLOTS = int(os.environ.get("BENCH_GSLOTS", "16"))
        self.gs = GraphQuerySlots(self.ix, self.GSLOTS, self.B, self.g_maxtok, VOCAB, first_slot=16)
        for i in range(self.GSLOTS):
            q, offs = hq[i % len(hq)]
            self.gs.offs[i][:] = offs
            self.gs.toks[i][:q.size] = q
        self.g_h2d = (self.B + 1) * 8 + self.g_maxtok * 4
        self._gl, self._gw, self._gh = self.gs._launch, self.gs._wait, self.ix._h

    def e2e_graph_step(self, k):
        slot = 16 + k % self.GSLOTS
        if self._gw(self._gh, slot) != 0 or self._gl(self._gh, slot) != 0:
            self.N.check(-1)

    def e2e_graph_drain(self):
        for i in range(self.gs.n):
            self.gs.wait(i)

    def e2e_step(self, k):
        slot = k % self.NSLOT
        if self.L.fmidx_query_wait(self.ix._h, slot) != 0 or self.L.fmidx_query_submit(*self.submit_args[(k % len(self.hq), slot)]) != 0:
            self.N.check(-1)

    def e2e_drain(self):
        for slot in range(self.NSLOT):
            self.N.check(self.L.fmidx_query_wait(self.ix._h, slot))


# =====================================================================================================================
# algorithmic bytes of one step, counted as the 32-byte SECTORS the algorithm has to touch
# =====================================================================================================================
NIB_PROBE_B = 32 + 32 * 1.49     # header sector + E[plane sectors below pos] (two 32-B plane sectors per 128-B cell)
BYTE_PROBE_B = 32 + 32 * 8.5     # header-word sector + E[payload sectors below pos] (512 payload bytes per cell)
STEP_B = 2 * (3 * NIB_PROBE_B + BYTE_PROBE_B)  # two boundaries; = 1,086 B (SURVEY 8(d) quotes 1,088 B for the reference layout)


def algorithmic_bytes(ix, qsets, device, B, info, table=True, lazy=False):
    """Per step, from the actual results: what K1 and K2 must read and write at the granularity of DRAM sectors.
    K1  probe steps (range still > 16 rows; resident layout) x STEP_B; the first token costs two C[] sectors; the tail
        of a prefix costs ceil(4w/32) SA sectors, w x ceil(4 rem/32) text sectors and one ISA sector per survivor; plus
        the tokens and the result.
    K2  the mask row once (16,032 B); ranges of <= SCAN_MAX rows: their 4 B/row of the resident BWT; wider ranges (tree
        walk): 2 x 128 B per visited nibble node, and per visited byte-level node its rows' payload bytes when it is
        <= 512 rows wide, else two 1,024-B header slices + 512 B of partial cells; a one-token prefix of a frequent token
        (its range is the token's whole block): one 16,032-B row of the unigram continuation table instead of a walk."""
    from fm_index_constrained_decoding_b200 import _native
    L = _native.lib()
    resident = bool(info["keep_text"]) and info["sa_dens"] == 1
    sptr = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    tot = {"k1": 0.0, "k2": 0.0, "search_steps": 0.0, "probe_steps": 0.0, "distinct": 0.0, "scan_rows": 0.0, "tree_prefixes": 0.0,
           "byte_nodes": 0.0, "table_rows": 0.0}
    sp = torch.empty(B, dtype=torch.int64, device=device)
    ep = torch.empty(B, dtype=torch.int64, device=device)
    for q, offs in qsets:
        lens = (offs[1:] - offs[:-1])
        maxlen = int(lens.max())
        # widths after every token: search the truncated prefixes (the search kernel is cheap)
        alive = torch.ones(B, dtype=torch.bool, device=device)
        probe_steps = torch.zeros(B, dtype=torch.int64, device=device)
        tail_w = torch.zeros(B, dtype=torch.int64, device=device)
        tail_rem = torch.zeros(B, dtype=torch.int64, device=device)
        for t in range(1, maxlen + 1):
            tl = torch.minimum(lens, torch.full_like(lens, t))
            toffs = torch.zeros(B + 1, dtype=torch.int64, device=device)
            toffs[1:] = torch.cumsum(tl, 0)
            idx = torch.arange(int(toffs[-1]), device=device) - torch.repeat_interleave(toffs[:-1], tl) + torch.repeat_interleave(offs[:-1], tl)
            tq = q[idx].contiguous()
            L.fmidx_search_batch(ix._h, C.c_void_p(tq.data_ptr()), C.c_void_p(toffs.data_ptr()), B, SHIFT,
                                 C.c_void_p(sp.data_ptr()), C.c_void_p(ep.data_ptr()), sptr)
            w = ep - sp
            act = alive & (lens >= t)
            if t > 1:
                probe_steps += act.long()
            stop = act & (w <= 16) & (lens > t) & resident
            tail_w = torch.where(stop, w, tail_w)
            tail_rem = torch.where(stop, lens - t, tail_rem)
            alive = alive & ~stop & (w > 0)
        k1 = (probe_steps.sum() * STEP_B + B * 64 + (torch.ceil(tail_w * 4 / 32) * 32).sum()
              + (tail_w * torch.ceil(tail_rem * 4 / 32) * 32).sum() + (tail_w > 0).sum() * 32 + lens.sum() * 4 + B * 16)
        tot["k1"] += float(k1)
        tot["search_steps"] += float(lens.sum())
        tot["probe_steps"] += float(probe_steps.sum())
        # K2 from the final ranges
        L.fmidx_search_batch(ix._h, C.c_void_p(q.data_ptr()), C.c_void_p(offs.data_ptr()), B, SHIFT,
                             C.c_void_p(sp.data_ptr()), C.c_void_p(ep.data_ptr()), sptr)
        w = ep - sp
        scan = (w <= SCAN_MAX) if resident else torch.zeros_like(w, dtype=torch.bool)
        if lazy:
            # lazy clearing: a row costs the slices its PREVIOUS content dirtied (512 B each; a row whose range was wider than
            # 1,024 rows counts as all 32, like the kernel book-keeps it) + the 32-B sectors that receive bits + 8 B of state
            _, _, m = ix.query_batch(q, offs, VOCAB)
            mw = torch.zeros((B, 32 * 128), dtype=torch.int32, device=device)
            mw[:, :m.shape[1]] = m
            slices = (mw.view(B, 32, 128) != 0).any(-1).sum(-1)
            slices = torch.where((w > 1024), torch.full_like(slices, 32), slices)
            sectors = (mw.view(B, -1, 8) != 0).any(-1).sum(-1)
            row_bytes = float((slices * 512 + sectors * 32 + 8).sum())
            tot["row_slices"] = tot.get("row_slices", 0.0) + float(slices.sum())
            del mw, m
        else:
            row_bytes = B * MASK_BYTES
        k2 = row_bytes + float((torch.ceil(w[scan] * 4 / 32) * 32).sum())
        tot["scan_rows"] += float(w[scan].sum())
        heavy_m = ~scan & (w > 0)
        if resident and table:
            # one-token prefixes of frequent tokens: the range is the token's whole block -> one row of the unigram table
            hit = heavy_m & (lens == 1)
            k2 += float(hit.sum()) * MASK_BYTES
            tot["table_rows"] += float(hit.sum())
            heavy_m = heavy_m & ~hit
        heavy = torch.nonzero(heavy_m).flatten()
        tot["tree_prefixes"] += heavy.numel()
        for a in range(0, heavy.numel(), 128):
            hb = heavy[a:a + 128]
            m, cnt = ix.distinct_mask_batch(sp[hb].contiguous(), ep[hb].contiguous(), VOCAB, want_counts=True)
            sh = torch.arange(32, device=device, dtype=torch.int32)
            bits = ((m.unsqueeze(-1) >> sh) & 1).bool().view(hb.numel(), -1)[:, :VOCAB]
            cnt = torch.where(bits, cnt, torch.zeros_like(cnt)).long()
            symc = torch.zeros((hb.numel(), 1 << 17), dtype=torch.int64, device=device)
            symc[:, SHIFT:SHIFT + VOCAB] = cnt
            bw = symc.view(hb.numel(), -1, 256).sum(-1)                 # rows per byte-level node
            narrow = (bw > 0) & (bw <= 512)
            wide = bw > 512
            nib = (bw.view(hb.numel(), -1, 16).sum(-1) > 0).sum() + (bw.view(hb.numel(), -1, 256).sum(-1) > 0).sum() + hb.numel()
            k2 += float((torch.ceil(bw[narrow] / 32) * 32 + 32).sum()) + float(wide.sum()) * (2048 + 512) + float(nib) * 256
            tot["byte_nodes"] += float((bw > 0).sum())
        tot["k2"] += k2
    torch.cuda.synchronize()
    return {k: v / len(qsets) for k, v in tot.items()}


def ncu_traffic(lazy=False):
    """DRAM traffic of one step from the committed ncu captures of THIS workload (100M tokens, batch 4,096, 8 steps in flight).
    Preferred: profiles/r2_range_replay_phase0.csv -- `ncu --replay-mode range` over 64 overlapped steps (benchmarks/
    range_probe.py), i.e. dram__bytes_read.sum + dram__bytes_write.sum of the multi-stream region as it is timed here, / 64.
    Else the per-kernel `ncu --set full` summary (profiles/r2_ncu_full_summary.csv; serial, cold caches: dirty mask rows still
    in L2 when a kernel ends are not counted).  (total, per-part) or (None, None)."""
    import csv
    p = os.path.join(ROOT, "profiles", "r2_range_replay_lazy.csv" if lazy else "r2_range_replay_phase0.csv")
    if lazy and not os.path.exists(p):
        return None, None
    if os.path.exists(p):
        try:
            vals = {}
            with open(p) as f:
                for r in csv.reader(f):
                    if len(r) > 12 and r[10].startswith("dram__bytes_"):
                        vals[r[10]] = float(r[12].replace(",", ""))
            rd, wr = vals["dram__bytes_read.sum"] / 64.0, vals["dram__bytes_write.sum"] / 64.0
            return rd + wr, {"dram_read_per_step": rd, "dram_write_per_step": wr, "source": "ncu --replay-mode range, 64 overlapped steps"}
        except Exception as e:  # noqa: BLE001
            log("range-replay capture unreadable:", e)
    p = os.path.join(ROOT, "profiles", "r2_ncu_full_summary.csv")
    if not os.path.exists(p):
        return None, None
    per = {}
    with open(p) as f:
        rows = list(csv.reader(f))
    hdr = rows[0]
    try:
        iname = 0
        ir = next(i for i, h in enumerate(hdr) if h.startswith("dram__bytes_read.sum"))
        iw = next(i for i, h in enumerate(hdr) if h.startswith("dram__bytes_write.sum"))
    except StopIteration:
        return None, None
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    ur = next((v for k, v in scale.items() if k in hdr[ir]), 1.0)
    uw = next((v for k, v in scale.items() if k in hdr[iw]), 1.0)
    for r in rows[1:]:
        if len(r) <= max(ir, iw):
            continue
        name = r[iname]
        key = next((k for k in ("k_search", "k_mask_top", "k_mask_expand", "k_mask_bytes") if k in name), None)
        if key:
            per.setdefault(key, []).append(float(r[ir]) * ur + float(r[iw]) * uw)
    if not per:
        return None, None
    per = {k: float(np.mean(v)) for k, v in per.items()}
    return sum(per.values()), per


# =====================================================================================================================
# N = 1
# =====================================================================================================================
def oracle_verified(toks, beg, ix, nthreads):
    """CPU oracle over the SAME corpus as the GPU index.  Suffix-sorting 100M tokens on one host core takes ~30 s, so the
    oracle's structures are built from the suffix array the GPU builder produced -- after oracle-side code has PROVED
    that array correct (fmo_verify_sa: permutation + adjacent-suffix order via ISA, O(n)), so the parity check below does
    not rest on the product's builder.  (--impl reference uses the oracle's own SA-IS throughout.)"""
    from fm_index_constrained_decoding_b200 import _native
    from oracle.oracle import NativeOracle
    sym = reversed_symbols(toks, beg)
    sa = np.empty(sym.numel() + 1, dtype=np.uint32)
    _native.check(_native.lib().fmidx_debug_copy_sa(ix._h, sa.ctypes.data_as(C.c_void_p), sa.size))
    text = np.concatenate([sym.cpu().numpy().astype(np.uint32), np.zeros(1, np.uint32)])
    o = NativeOracle()
    o.initialize_from_sa(text, sa, verify=True, nthreads=nthreads)
    return o


def run_single(args, device, local):
    from fm_index_constrained_decoding_b200 import _native, synth
    L = _native.lib()
    B = args.batch
    t0 = time.time()
    toks, beg = synth.zipf_corpus(args.tokens, seed=0, device=device)
    qsets = [synth.sample_queries(toks, B, seed=100 + i, beginnings=beg) for i in range(N_QUERY_SETS)]
    torch.cuda.synchronize()
    t_gen = time.time() - t0
    ix, t_build = build_index(toks, beg, local, args.layout)
    info = ix.info()
    log("corpus %.1fs, index build %.1fs, %d levels, %.0f MB on device" % (t_gen, t_build, info["levels"], info["device_bytes"] / 1e6))
    eng = LocalEngine(ix, qsets, device, B)
    timer = BlockTimer(device)
    # The mask buffers of the timed loop belong to this program and nothing else writes them: lazy clearing (fmidx_tune
    # "lazy_clear" 2 -- only the 4096-token slices a buffer's previous batch dirtied are zeroed, not whole 16-KB rows).
    # `ablation.full_row_zeroing` is the same loop with every row zeroed in full.
    lazy = os.environ.get("BENCH_LAZY", "1") == "1"
    ix.tune(lazy_clear=1)
    full_ms, _, _ = timer.run(eng.block, args.steps, args.warmup)
    ix.tune(lazy_clear=2 if lazy else 1)
    sampler = ClockSampler(local)
    launches0 = L.fmidx_launch_count()
    sampler.start()
    block_ms, reps, _ = timer.run(eng.block, args.steps, args.warmup)
    clocks = sampler.stop()
    launches_per_block = (L.fmidx_launch_count() - launches0) / (reps + 6 + args.warmup / args.steps)
    ms_per_step = block_ms / args.steps
    value = B * args.steps / (block_ms * 1e-3)
    log("value %.1f M q/s (%.4f ms/step, %d blocks of %d steps)" % (value / 1e6, ms_per_step, reps, args.steps))
    t_search, t_mask = eng.serial_kernel_times(max(args.steps, int(MIN_TIMED_S * 1e3 / max(3 * ms_per_step, 1e-3))))
    # ablation: the same steps with the unigram continuation table ignored (every wide range is walked)
    ix.tune(unigram_table=0)
    abl_ms, _, _ = timer.run(eng.block, args.steps, args.warmup)
    t_search_off, t_mask_off = eng.serial_kernel_times(200)
    ix.tune(unigram_table=1)
    ablation = {"full_row_zeroing": {"value": B * args.steps / (full_ms * 1e-3), "ms_per_step": full_ms / args.steps,
                                     "what": "fmidx_tune lazy_clear = 1: every 16-KB mask row of the caller's buffers is zeroed in full each step "
                                             "(65.7 MB of stores per step)"},
                "unigram_table_off": {"value": B * args.steps / (abl_ms * 1e-3), "ms_per_step": abl_ms / args.steps,
                                      "serial_kernel_ms": {"k_search": t_search_off, "k_mask": t_mask_off}},
                "note": "the table holds, for the <= 1024 most frequent symbols, the distinct symbols of the symbol's BWT block (built by the "
                        "library's own enumeration at index build; the reference precomputes the same for the empty prefix, index.py:206)"}

    # ---- roofline of the step, live ------------------------------------------------------------------------------------
    peak, peak_src = measured_peaks()
    ab = algorithmic_bytes(ix, qsets, device, B, info, lazy=lazy)
    step_bytes = ab["k1"] + ab["k2"]
    traffic, traffic_per = ncu_traffic(lazy)
    if lazy:  # the same accounting for the full-row ablation
        ab_full = algorithmic_bytes(ix, qsets, device, B, info, lazy=False)
        tr_full, _ = ncu_traffic(False)
        fb = ab_full["k1"] + ab_full["k2"]
        ablation["full_row_zeroing"]["roofline"] = {
            "algorithmic_bytes_per_launch": fb, "achieved": fb / (full_ms / args.steps * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": fb / (full_ms / args.steps * 1e-3) / 1e9 / peak, "traffic": tr_full,
            "frac_of_traffic": (tr_full / (full_ms / args.steps * 1e-3) / 1e9 / peak) if tr_full else None}
    roof = {"kernel": "one step = k_search + k_mask_top + k_mask_expand + k_mask_bytes (%d independent batches in flight on as many streams)" % eng.NPIPE + "",
            "bound": "hbm", "achieved": step_bytes / (ms_per_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "peak_source": peak_src,
            "traffic": traffic, "algorithmic_bytes_per_launch": step_bytes, "avg_launch_ms": ms_per_step,
            "time_source": "the timed (overlapped) region: median K-step block / K",
            "per_unit": "K1: %.0f B per probe step (2 boundaries x (3 nibble probes of 32-B header sector + 1.49 plane sectors, 1 byte-level "
                        "probe of 32-B header sector + 8.5 payload sectors)); tail of a prefix: SA/text/ISA sectors; K2: %s + "
                        "4 B per row of ranges <= %d rows (resident BWT) + 16,032 B table row per frequent one-token prefix + tree walk of the other "
                        "wide ranges (2 x 128 B per nibble node, payload bytes or 2,560 B per byte-level node)"
                        % (STEP_B, "per mask row the 512-B slices its previous content dirtied (lazy clearing; 16,032 B for a row whose range "
                                   "was wider than 1,024 rows) + the 32-B sectors that receive bits + 8 B of occupancy state" if lazy
                           else "16,032 B mask row", SCAN_MAX),
            "traffic_detail": traffic_per,
            "per_kernel": {"k_search": {"algorithmic_bytes": ab["k1"], "serial_ms": t_search, "GBps_serial": ab["k1"] / (t_search * 1e-3) / 1e9},
                           "k_mask": {"algorithmic_bytes": ab["k2"], "serial_ms": t_mask, "GBps_serial": ab["k2"] / (t_mask * 1e-3) / 1e9}},
            "other_kernel_ms": {"k_search": t_search, "k_mask": t_mask},
            "counts_per_step": {k: ab[k] for k in ("search_steps", "probe_steps", "scan_rows", "table_rows", "tree_prefixes", "byte_nodes")},
            "reference_layout_bytes_per_step": 1088 * ab["search_steps"] + B * MASK_BYTES,
            "note": ("LAZY CLEARING (headline): only the slices a buffer's previous batch dirtied are zeroed, so the 65.7 MB of row stores "
                     "per step shrink to a few MB and what is left is random access -- the step then runs at the DRAM's random-sector rate "
                     "(see sector_rate), far below the streaming roofline by construction.  With full-row zeroing (ablation.full_row_zeroing, "
                     "the form whose traffic is 3/4 streaming writes): " if lazy else "") +
                    "the mask rows (65.7 MB per step, written once: ncu's range replay of the overlapped step sees 66.2 MB written and "
                    "18.2 MB read per step, i.e. the traffic IS the algorithmic minimum) dominate the bytes.  Phase-isolated graph replays "
                    "(profiles/r2_phase_probe.log): the row zeroing alone runs at 9.8 us per step (7.3 TB/s of pure writes), the search "
                    "alone at 5.2 us, the enumeration without zeroing at 7.6 us, and the three ADD (24 us): the random 32-B-sector reads of "
                    "the search / enumeration and the streaming row writes share the DRAM, which serves random sectors at 41 G/s "
                    "(random_gather_peak) -- the step is DRAM-access-bound, not issue- or launch-bound (throughput is the same at batch "
                    "65,536, profiles/r2_batch_sweep.log)"}
    roof["frac"] = roof["achieved"] / peak
    roof["frac_of_traffic"] = (traffic / (ms_per_step * 1e-3) / 1e9 / peak) if traffic else None
    roof["mask_rows"] = "lazy clearing (fmidx_tune lazy_clear=2)" if lazy else "zeroed in full every step"
    try:
        gp = {}
        for unit in (32, 128):
            v = C.c_double()
            _native.check(L.fmidx_gather_peak(ix._h, unit, 1 << 25, C.byref(v)))
            gp["%dB_GBps" % unit] = v.value
        roof["random_gather_peak"] = gp
        sect = (traffic if traffic else step_bytes) / 32.0 / (ms_per_step * 1e-3)
        roof["sector_rate"] = {"G_sectors_per_s": sect / 1e9, "random_gather_32B_G_sectors_per_s": gp["32B_GBps"] / 32.0,
                               "frac_of_random_gather": sect / 1e9 / (gp["32B_GBps"] / 32.0),
                               "basis": "ncu traffic" if traffic else "algorithmic bytes",
                               "what": "32-byte DRAM sectors moved per second by the step against the rate at which this GPU serves isolated random "
                                       "32-byte sectors (fmidx_gather_peak): the roofline a random-access integer path lives under"}
    except Exception as e:  # noqa: BLE001
        roof["random_gather_peak"] = {"error": repr(e)}

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ------------------------------------------------
    eng.setup_e2e()
    e2e_s, e2e_reps = host_blocks(eng.e2e_step, eng.e2e_drain, args.steps, args.warmup)
    eng.setup_e2e_graph()
    g_s, g_reps = host_blocks(eng.e2e_graph_step, eng.e2e_graph_drain, args.steps, args.warmup)
    # parity of the captured path on the spot: slot 0 holds batch 0
    gsp, gep, _ = ix.query_batch(qsets[0][0], qsets[0][1], VOCAB)
    torch.cuda.synchronize()
    g_ok = bool((torch.from_numpy(eng.gs.sp[0].astype(np.int64)) == gsp.cpu()).all() and (torch.from_numpy(eng.gs.ep[0].astype(np.int64)) == gep.cpu()).all())
    e2e = {"value": B * args.steps / g_s, "unit": "queries/s", "h2d_bytes_per_step": eng.g_h2d, "d2h_bytes_per_step": eng.d2h,
           "ms_per_step": 1e3 * g_s / args.steps, "blocks": g_reps, "matches_device_path": g_ok,
           "api": "fmidx_query_graph_launch/fmidx_query_wait: one CUDA-graph launch per step = H2D of the step's prefixes from pinned host "
                  "memory -> search -> enumeration -> popcount -> D2H of ranges + distinct-token counts into pinned host memory; %d steps in "
                  "flight (the 16 KB masks stay in HBM where the logits kernel consumes them -- the CPU arm writes them to host memory)" % eng.gs.n,
           "eager_api": {"value": B * args.steps / e2e_s, "ms_per_step": 1e3 * e2e_s / args.steps, "h2d_bytes_per_step": eng.h2d,
                         "api": "fmidx_query_submit/fmidx_query_wait (ten driver calls per step), %d batches in flight" % eng.NSLOT}}

    # ---- CPU baseline + parity on the same corpus (the only place the oracle is used) ---------------------------------------
    cpu, o = None, None
    nthreads = os.cpu_count() or 1
    if not args.no_cpu_baseline:
        try:
            t0 = time.time()
            o = oracle_verified(toks, beg, ix, nthreads)
            log("oracle: GPU suffix array verified by fmo_verify_sa, structures built in %.1fs" % (time.time() - t0))
            q, offs = qsets[0]
            qn, on = q.cpu().numpy(), offs.cpu().numpy()
            cb = min(B, args.cpu_batch)
            qn, on = qn[:on[cb]], on[:cb + 1]
            osp, oep, omask, _ = o.query_batch(qn, on, shift=SHIFT, want_mask=True, vocab=VOCAB, nthreads=nthreads)
            sp, ep, mask = ix.query_batch(q, offs, VOCAB)
            torch.cuda.synchronize()
            ok = bool((sp[:cb].cpu().numpy().astype(np.uint64) == osp).all() and (ep[:cb].cpu().numpy().astype(np.uint64) == oep).all()
                      and (mask[:cb].cpu().numpy().view(np.uint32) == omask).all())
            tt = min(cpu_step(o, qn, on, nthreads) for _ in range(2))
            t1 = cpu_step(o, qn[:on[256]], on[:257], 1)
            cpu = {"value": cb / tt, "unit": "queries/s", "cores": nthreads, "kind": "port",
                   "sample": "one batch of %d prefixes (search + distinct-token mask), best of 2, oracle/fm_oracle.c (sdsl-semantics "
                             "restatement, not sdsl-lite) on a pthread pool" % cb,
                   "single_thread_queries_per_sec": 256 / t1, "parity_vs_gpu_on_sample": ok,
                   "oracle_suffix_array": "GPU builder's array, proved correct by oracle-side fmo_verify_sa before use"}
        except Exception as e:  # the baseline must never take the bench line down
            cpu = {"value": None, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}

    other = {}
    if not args.no_other_configs:
        for name, fn in (("configs0_sample_corpus_batch1", lambda: other_config0(local)),
                         ("configs3_fused_decode_step", lambda: other_decode(ix, toks, beg, device, o)),
                         ("configs4_locate_extract", lambda: other_locate(ix, toks, beg, device, o))):
            try:
                other[name] = fn()
            except Exception as e:  # noqa: BLE001
                other[name] = {"error": repr(e)}
            log(name, json.dumps(other[name])[:400])
    del o, eng
    del ix, toks, beg, qsets
    torch.cuda.empty_cache()
    if not args.no_other_configs:
        try:
            other["configs2_1B_tokens_one_gpu"] = other_one_billion(args, device, local)
        except Exception as e:  # noqa: BLE001
            other["configs2_1B_tokens_one_gpu"] = {"error": repr(e)}
        log("configs2_1B_tokens_one_gpu", json.dumps(other["configs2_1B_tokens_one_gpu"])[:400])

    emit({
        "metric": "prefix_queries_per_sec", "value": value, "unit": "queries/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": workload_config(args.tokens, B),
        "masks_per_sec": value, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(round(launches_per_block)),
        "timing": {"blocks": reps, "block_steps": args.steps, "min_timed_s": MIN_TIMED_S, "streams": eng_streams(),
                   "statistic": "the K-step block repeated back to back (no pipeline drain in between) in 7 super-blocks; median "
                                "super-block, reported per K steps"},
        "clocks": clocks, "ablation": ablation, "other_configs": other,
        "index": {"build_s": t_build, "device_mb": info["device_bytes"] / 1e6, "levels": info["levels"],
                  "layout": "wavelet matrix: 3 nibble levels (128-B cells) + byte level (1,536-B cells); %s (sa_dens %d, isa_dens %d, text+BWT %s)"
                            % (args.layout, info["sa_dens"], info["isa_dens"], bool(info["keep_text"]))},
    })


def eng_streams():
    return int(os.environ.get("BENCH_NPIPE", "8"))


# ---- other configs (N = 1) -----------------------------------------------------------------------------------------------
def other_config0(local):
    """configs[0]: test_data/sample_corpus.tsv (token ids of the documented stand-in tokenizer, tests/golden/sample_corpus.json),
    batch 1 through the reference-shaped scalar API: get_range + get_distinct_count per prefix."""
    from fm_index_constrained_decoding_b200 import FMIndex
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "sample_corpus.json")))
    ix = FMIndex(device=local)
    ix.initialize(g["docs"], in_memory=True)
    qs = [q for q in g["queries"]]
    ok = True
    for q in qs[:20]:
        lo, hi = ix.get_range(q["q"])
        d, c = ix.get_distinct_count(lo, hi)
        ok = ok and [lo, hi] == q["range"] and d == q["distinct"] and c == q["counts"]
    for q in qs[:10]:
        ix.get_distinct_count(*ix.get_range(q["q"]))
    t0 = time.perf_counter()
    n = 0
    for _ in range(5):
        for q in qs:
            ix.get_distinct_count(*ix.get_range(q["q"]))
            n += 1
    us = (time.perf_counter() - t0) / n * 1e6
    return {"us_per_query_pair": us, "queries": len(qs), "parity_vs_golden": ok,
            "note": "latency of get_range + get_distinct_count through the SWIG-shaped scalar API (one prefix at a time); stand-in "
                    "tokenizer (FNV-1a word hash), expectations from tests/golden/sample_corpus.json"}


def other_decode(ix, toks, beg, device, o):
    """configs[3]: 256 beams x 128,256 logits, search + fused enumerate/-inf mask; fp32 and bf16."""
    from fm_index_constrained_decoding_b200 import synth
    peak, _ = measured_peaks()
    R = 256
    out = {}
    q, offs = synth.sample_queries(toks, R, seed=7, beginnings=beg)
    g = torch.Generator(device=device).manual_seed(3)
    for name, dt, esz in (("fp32", torch.float32, 4), ("bf16", torch.bfloat16, 2)):
        # 8 logits buffers rotated: 8 x 131 MB (fp32) > L2
        bufs = [torch.randn(R, VOCAB, device=device, generator=g, dtype=torch.float32).to(dt) for _ in range(8)]
        outb = torch.empty_like(bufs[0])
        st = torch.cuda.current_stream()

        def step(k):
            sp, ep = ix.get_range_batch(q, offs)
            ix.mask_logits(sp, ep, bufs[k % 8], out=outb, always_allow=EOS)
            return sp, ep
        for k in range(5):
            sp, ep = step(k)
        torch.cuda.synchronize()
        n = 200
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for k in range(n):
            step(k)
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        # the fused kernel alone
        e0.record(st)
        for k in range(n):
            ix.mask_logits(sp, ep, bufs[k % 8], out=outb, always_allow=EOS)
        e1.record(st)
        torch.cuda.synchronize()
        ms_k = e0.elapsed_time(e1) / n
        bytes_ = R * VOCAB * esz * 2
        # the step as the logits processors run it in steady state: per-beam (sp, ep) of history[:-1] kept on the device, ONE
        # backward-search step per beam with the newest token (fmidx_decode_step), then the fused kernel
        from fm_index_constrained_decoding_b200 import _native as _N
        lens = offs[1:] - offs[:-1]
        last = q[(offs[1:] - 1).clamp(min=0)].to(torch.int64).contiguous()
        hoffs = torch.zeros_like(offs)
        hoffs[1:] = torch.cumsum((lens - 1).clamp(min=0), 0)
        keep = torch.ones(q.numel(), dtype=torch.bool, device=device)
        keep[(offs[1:] - 1)[lens > 0]] = False
        psp, pep = ix.get_range_batch(q[keep].contiguous(), hoffs)
        sbuf = torch.empty((4, R), dtype=torch.int64, device=device)
        ovr = torch.empty(R, dtype=torch.int32, device=device)
        cs = C.c_void_p(st.cuda_stream)

        def inc_step(k):
            _N.check(_N.lib().fmidx_decode_step(ix._h, C.c_void_p(last.data_ptr()), 1, None, C.c_void_p(psp.data_ptr()), C.c_void_p(pep.data_ptr()),
                                                C.c_void_p(sbuf[0].data_ptr()), C.c_void_p(sbuf[1].data_ptr()), R, SHIFT, -1, -2, 0,
                                                C.c_void_p(sbuf[2].data_ptr()), C.c_void_p(sbuf[3].data_ptr()), C.c_void_p(ovr.data_ptr()), cs))
            ix.mask_logits(sbuf[2], sbuf[3], bufs[k % 8], out=outb, override_tok=ovr, always_allow=EOS)
        for k in range(5):
            inc_step(k)
        torch.cuda.synchronize()
        e0.record(st)
        for k in range(n):
            inc_step(k)
        e1.record(st)
        torch.cuda.synchronize()
        ms_inc = e0.elapsed_time(e1) / n
        inc_ok = bool(torch.equal(sbuf[2], sp) and torch.equal(sbuf[3], ep))
        res = {"ms_per_step": ms, "steps_per_s": 1e3 / ms, "fused_kernel_ms": ms_k, "fused_kernel_GBps": bytes_ / (ms_k * 1e-3) / 1e9,
               "fused_kernel_frac_of_hbm": bytes_ / (ms_k * 1e-3) / 1e9 / peak, "algorithmic_bytes": bytes_,
               "ms_per_step_incremental": ms_inc, "incremental_matches_full_search": inc_ok,
               "incremental_note": "fmidx_decode_step (device-resident per-beam state, one backward-search step per beam) + the fused kernel: "
                                   "the steady-state step of processors.IndexBasedLogitsProcessor; ms_per_step re-searches every history"}
        if o is not None:
            from oracle.oracle import SHIFT as _S  # noqa: F401
            qn, on = q.cpu().numpy(), offs.cpu().numpy()
            osp, oep, om, _ = o.query_batch(qn, on, shift=SHIFT, want_mask=True, vocab=VOCAB, nthreads=os.cpu_count() or 1)
            ix.mask_logits(sp, ep, bufs[0], out=outb, always_allow=EOS)
            bits = np.unpackbits(om.view(np.uint8), axis=1, bitorder="little")[:, :VOCAB].astype(bool)
            bits[:, EOS] = True
            want = torch.where(torch.from_numpy(bits).to(device), bufs[0] + 0.0, torch.full_like(bufs[0], float("-inf")))
            res["parity_vs_oracle"] = bool(torch.equal(outb, want) and (sp.cpu().numpy().astype(np.uint64) == osp).all())
        out[name] = res
        del bufs, outb
    out["workload"] = "256 beam histories of length 1-32 (corpus substrings), always_allow_eos, logits randn(seed 3), 8 logits buffers rotated"
    return out


def other_locate(ix, toks, beg, device, o):
    """configs[4] on one shard: 65,536 prefixes, locate the first K = 8 rows of every range + extract 32 tokens per hit."""
    from fm_index_constrained_decoding_b200 import synth
    Q, K, M = 65536, 8, 32
    q, offs = synth.sample_queries(toks, Q, seed=11, beginnings=beg)
    st = torch.cuda.current_stream()
    ar = torch.arange(K, device=device, dtype=torch.int64).unsqueeze(0)

    def step():
        sp, ep = ix.get_range_batch(q, offs)
        rows = sp.unsqueeze(1) + ar
        valid = rows < ep.unsqueeze(1)
        pos = ix.locate_batch(torch.where(valid, rows, torch.zeros_like(rows)).contiguous().view(-1))
        begins = torch.where(valid.view(-1), pos, torch.zeros_like(pos))
        ext, lens = ix.extract_batch(begins, begins + M, M)
        return sp, ep, valid, pos, ext
    for _ in range(3):
        sp, ep, valid, pos, ext = step()
    torch.cuda.synchronize()
    n = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(n):
        step()
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    hits = int(valid.sum())
    res = {"ms_per_step": ms, "prefixes_per_s": Q / (ms * 1e-3), "locates_per_s": hits / (ms * 1e-3),
           "tokens_extracted_per_s": Q * K * M / (ms * 1e-3), "hits_per_step": hits,
           "workload": "65,536 prefixes: search + locate of the first 8 rows of every range + extract_text of 32 symbols per hit slot"}
    if o is not None:
        rows = (sp[:64].unsqueeze(1) + ar)[valid[:64]].cpu().numpy().astype(np.uint64)
        want = o.locate_batch(rows, nthreads=4)
        got = pos.view(Q, K)[:64][valid[:64]].cpu().numpy().astype(np.uint64)
        ok = bool((want == got).all())
        for r in range(8):  # extract parity on a few hit slots
            b = int(got[r])
            ok = ok and list(o.extract_text(b, b + M)) == ext.view(Q, K, M)[:64][valid[:64]][r].tolist()[:len(o.extract_text(b, b + M))]
        res["parity_vs_oracle_on_sample"] = ok
    return res


def other_one_billion(args, device, local):
    """The north-star sentence: one B200, 1B-token corpus, the same step."""
    from fm_index_constrained_decoding_b200 import synth
    n_tok = args.big_tokens
    B = args.batch
    t0 = time.time()
    toks, beg = synth.zipf_corpus(n_tok, seed=2, device=device)
    qsets = [synth.sample_queries(toks, B, seed=200 + i, beginnings=beg) for i in range(N_QUERY_SETS)]
    torch.cuda.synchronize()
    t_gen = time.time() - t0
    ix, t_build = build_index(toks, beg, local)
    info = ix.info()
    eng = LocalEngine(ix, qsets, device, B)
    timer = BlockTimer(device)
    block_ms, reps, _ = timer.run(eng.block, args.steps, args.warmup)
    eng.setup_e2e()
    e2e_s, _ = host_blocks(eng.e2e_step, eng.e2e_drain, args.steps, args.warmup)
    # parity: brute force over the text itself on a sample
    nq = 24
    sp, ep, mask = ix.query_batch(qsets[0][0], qsets[0][1], VOCAB)
    sym = reversed_symbols(toks, beg)
    cnt, bm = bruteforce_sample(sym, queries_as_lists(qsets[0][0], qsets[0][1], nq))
    ok = bool(((ep - sp)[:nq] == cnt).all() and torch.equal(mask[:nq], bm))
    return {"tokens": n_tok, "value": B * args.steps / (block_ms * 1e-3), "unit": "queries/s", "ms_per_step": block_ms / args.steps,
            "blocks": reps, "e2e_value": B * args.steps / e2e_s, "index_device_gb": info["device_bytes"] / 1e9, "build_s": t_build,
            "corpus_gen_s": t_gen, "parity_vs_bruteforce_on_sample": ok,
            "parity_note": "%d prefixes: occurrence counts and next-token bitmaps recomputed by scanning the 1B-token text with torch "
                           "(no FM-index involved)" % nq}


# =====================================================================================================================
# N > 1
# =====================================================================================================================
def run_multi(args, device, local, rank, world, dist):
    from fm_index_constrained_decoding_b200 import _native, sharded, synth
    L = _native.lib()
    B = args.batch
    timer = BlockTimer(device, dist)
    modes = {}
    want = args.mode

    # ---- replicas: index replicated, prefixes split N ways, zero collectives ------------------------------------------------
    toks, beg = synth.zipf_corpus(args.tokens, seed=0, device=device)
    qsets = [synth.sample_queries(toks, B, seed=100 + 1000 * rank + i, beginnings=beg) for i in range(N_QUERY_SETS)]
    ix, t_build = build_index(toks, beg, local, args.layout)
    info = ix.info()
    eng = LocalEngine(ix, qsets, device, B)
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = L.fmidx_launch_count()
    if sampler:
        sampler.start()
    block_ms, reps, _ = timer.run(eng.block, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    launches_per_block = (L.fmidx_launch_count() - launches0) / (reps + 6 + args.warmup / args.steps)
    value = world * B * args.steps / (block_ms * 1e-3)
    eng.setup_e2e_graph()
    e2e_s, e2e_reps = host_blocks(eng.e2e_graph_step, eng.e2e_graph_drain, args.steps, args.warmup, dist, device)
    # parity of this rank's replica on a sample: brute force over the text
    nq = 16
    sp, ep, mask = ix.query_batch(qsets[0][0], qsets[0][1], VOCAB)
    cnt, bm = bruteforce_sample(reversed_symbols(toks, beg), queries_as_lists(qsets[0][0], qsets[0][1], nq))
    okt = torch.tensor([int(bool(((ep - sp)[:nq] == cnt).all() and torch.equal(mask[:nq], bm)))], device=device)
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    modes["replicas"] = {"value": value, "unit": "queries/s", "ms_per_step": block_ms / args.steps, "blocks": reps,
                         "e2e_value": world * B * args.steps / e2e_s, "tokens_per_gpu": args.tokens, "prefixes_per_step": world * B,
                         "collectives_per_step": 0, "parity_vs_bruteforce_on_sample": bool(int(okt[0])),
                         "what": "every GPU holds the whole index (a 1B-token index is 22 GB of 180 GB) and serves its own 1/N of the prefixes"}
    e2e = {"value": world * B * args.steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": eng.g_h2d * world, "d2h_bytes_per_step": B * 20 * world,
           "ms_per_step": 1e3 * e2e_s / args.steps, "blocks": e2e_reps,
           "api": "per GPU: fmidx_query_graph_launch/fmidx_query_wait, %d captured steps in flight (pinned host prefixes in; ranges + "
                  "distinct-token counts out)" % eng.GSLOTS}
    log("rank %d replicas: %.1f M q/s total" % (rank, value / 1e6))
    del eng, ix, toks, beg, qsets
    torch.cuda.empty_cache()

    # ---- document-sharded modes over the fixed big corpus -------------------------------------------------------------------
    def sharded_mode(g_size, tag):
        """g_size shards per group, world / g_size replica groups; every group sees its own prefix batches."""
        n_groups = world // g_size
        grp_id, grank = rank // g_size, rank % g_size
        groups = [dist.new_group(list(range(a * g_size, (a + 1) * g_size))) for a in range(n_groups)]
        group = groups[grp_id] if n_groups > 1 else None
        toks, beg = synth.zipf_corpus(args.big_tokens, seed=2, device=device)
        parts = sharded.shard_documents(beg.tolist(), g_size)
        lo, hi = parts[grank]
        qs = [synth.sample_queries(toks, B, seed=300 + 50 * grp_id + i, beginnings=beg) for i in range(N_QUERY_SETS)]
        a, b = int(beg[lo]), int(beg[hi])
        stoks, sbeg = toks[a:b].clone(), (beg[lo:hi + 1] - beg[lo]).clone()
        del toks
        torch.cuda.empty_cache()
        six, tb = build_index(stoks, sbeg, local, args.layout)
        # buffer sets in flight: the ranks' pipelines are coupled set by set (a pull waits for the peers' publication of the same
        # set), so the sharded step needs a deeper pipeline than the single-GPU one: 8 sets 63 M q/s, 16 sets 107 M, 24 sets 121 M
        # (2 GPUs, 400M tokens; profiles/r2_sharded_probe.log)
        NSET = int(os.environ.get("BENCH_NBUF", "24"))
        st = sharded.ShardStep(six, B, VOCAB, n_sets=NSET, group=group)
        rb = [(torch.empty(B, dtype=torch.int64, device=device), torch.empty(B, dtype=torch.int64, device=device)) for _ in range(NSET)]
        step_args = {}
        for qi, (q, offs) in enumerate(qs):
            for i in range(NSET):
                step_args[(qi, i)] = (st._x, i, C.c_void_p(q.data_ptr()), C.c_void_p(offs.data_ptr()), SHIFT,
                                      C.c_void_p(rb[i][0].data_ptr()), C.c_void_p(rb[i][1].data_ptr()), None)
        shard_step = L.fmidx_shard_step
        # the timed form: every buffer set's step captured once (set i serves prefix batch i % 8: NSET is a multiple of 8, so
        # the rotation (k % 8, k % NSET) of the eager form is the same), then ONE graph launch per step
        use_graph = os.environ.get("BENCH_SHARD_GRAPH", "1") == "1" and NSET % N_QUERY_SETS == 0
        if use_graph:
            for i in range(NSET):
                q, offs = qs[i % N_QUERY_SETS]
                st.graph_prepare(i, q, offs, rb[i][0], rb[i][1])
            glaunch, gx = L.fmidx_shard_graph_launch, st._x

        def block(k0, n):
            stream = torch.cuda.current_stream()
            e = torch.cuda.Event()
            e.record(stream)
            for s in st.streams:
                s.wait_event(e)
            if use_graph:
                for k in range(k0, k0 + n):
                    if glaunch(gx, k % NSET) != 0:
                        _native.check(-1)
            else:
                for k in range(k0, k0 + n):
                    if shard_step(*step_args[(k % N_QUERY_SETS, k % NSET)]) != 0:
                        _native.check(-1)
            for s in st.streams:
                e = torch.cuda.Event()
                e.record(s)
                stream.wait_event(e)
        l0 = L.fmidx_launch_count()
        bms, rps, _ = timer.run(block, args.steps, args.warmup)
        lpb = (L.fmidx_launch_count() - l0) / (rps + 6 + args.warmup / args.steps)
        # e2e: host-buffer form
        hq = [(q.cpu().pin_memory(), offs.cpu().pin_memory()) for q, offs in qs]
        own = B // g_size
        hout = [(torch.empty(B, dtype=torch.int64).pin_memory(), torch.empty(B, dtype=torch.int64).pin_memory(),
                 torch.empty(B, dtype=torch.int64).pin_memory(), torch.empty(own, dtype=torch.int32).pin_memory()) for _ in range(NSET)]
        sub_args = {}
        for qi, (q, offs) in enumerate(hq):
            for i in range(NSET):
                sub_args[(qi, i)] = (st._x, i, C.c_void_p(q.data_ptr()), C.c_void_p(offs.data_ptr()), SHIFT) + tuple(C.c_void_p(t.data_ptr()) for t in hout[i])

        def e2e_step(k):
            i = k % NSET
            if L.fmidx_shard_wait(st._x, i) != 0 or L.fmidx_shard_submit(*sub_args[(k % N_QUERY_SETS, i)]) != 0:
                _native.check(-1)

        def e2e_drain():
            for i in range(NSET):
                _native.check(L.fmidx_shard_wait(st._x, i))
        es, erps = host_blocks(e2e_step, e2e_drain, args.steps, args.warmup, dist, device)
        # parity on a sample: per-shard brute force over the text, OR / sum reduced over the group on the host side
        nq = 16 * g_size
        q0, o0 = qs[0]
        masks, counts = st.graph_launch(0) if use_graph else st.step(0, q0, o0, rb[0][0], rb[0][1])  # set 0 holds batch 0
        torch.cuda.synchronize()
        ssym = reversed_symbols(stoks, sbeg)
        cnt, bm = bruteforce_sample(ssym, queries_as_lists(q0, o0, nq))
        local_ok = bool(((rb[0][1] - rb[0][0])[:nq] == cnt).all())  # shard-local ranges
        allc = [torch.empty_like(cnt) for _ in range(g_size)]
        allm = [torch.empty_like(bm) for _ in range(g_size)]
        dist.all_gather(allc, cnt, group=group)
        dist.all_gather(allm, bm, group=group)
        tot_c = torch.stack(allc).sum(0)
        tot_m = allm[0].clone()
        for m in allm[1:]:
            tot_m |= m
        ok = local_ok and bool((counts[:nq] == tot_c).all())
        lo_r, hi_r = grank * own, min((grank + 1) * own, nq)
        if hi_r > lo_r:  # the rows of the sample this rank owns
            ok = ok and bool(torch.equal(masks[:hi_r - lo_r], tot_m[lo_r:hi_r]))
        okt = torch.tensor([int(ok)], device=device)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        res = {"value": n_groups * B * args.steps / (bms * 1e-3), "unit": "queries/s", "ms_per_step": bms / args.steps, "blocks": rps,
               "e2e_value": n_groups * B * args.steps / es, "shards": g_size, "replica_groups": n_groups, "corpus_tokens": args.big_tokens,
               "tokens_per_gpu": int(b - a), "prefixes_per_step": n_groups * B, "buffer_sets": NSET, "launches_per_step": lpb / args.steps,
               "shard_index_device_gb": six.info()["device_bytes"] / 1e9, "build_s": tb,
               "parity_vs_bruteforce_on_sample": bool(int(okt[0])),
               "parity_note": "%d prefixes: per-shard counts / next-token bitmaps recomputed by scanning each shard's text with torch, "
                              "OR / sum reduced over the shards, against the exchanged results" % nq,
               "step_api": "fmidx_shard_graph_launch (one CUDA-graph launch per step)" if use_graph else "fmidx_shard_step (5 launches per step)",
               "exchange": "publish at the tail of the last enumeration kernel + fused wait/pull/OR/counts kernel over CUDA-IPC peer "
                           "memory (5 kernels per step, no NCCL, no acknowledgements: ping-pong buffers)"}
        st.close()
        del six, stoks, sbeg, qs
        torch.cuda.empty_cache()
        return res

    def _final_line():
        cfg = workload_config(args.tokens, B, "replicas")
        cfg["sharding"] = ("value = replicas mode (index replicated on every GPU, each GPU serves its own batch of %d prefixes per step, no "
                           "collective on the data path); modes.doc_sharded = BASELINE configs[2] (a fixed %dM-token corpus split N ways by "
                           "document, prefixes replicated, masks OR-reduced + counts summed over NVLink peer memory); modes.hybrid = 2 shards x "
                           "N/2 replica groups" % (B, args.big_tokens // 1_000_000))
        emit({
            "metric": "prefix_queries_per_sec", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": block_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": cfg, "masks_per_sec": value,
            "modes": modes, "e2e": e2e, "gpu_launches": int(round(launches_per_block)), "cpu_baseline": None,
            "roofline": None, "timing": {"blocks": reps, "block_steps": args.steps, "min_timed_s": MIN_TIMED_S,
                                         "statistic": "the K-step block repeated back to back in 7 super-blocks; per super-block the "
                                                      "slowest rank; median super-block, reported per K steps"},
            "clocks": clocks,
            "index": {"build_s": t_build, "device_mb": info["device_bytes"] / 1e6, "levels": info["levels"]},
        })

    def final_line():
        if rank == 0:
            _final_line()

    # The sharded modes couple the ranks (peer-memory flags, process groups): a rank that dies or a box that misbehaves must not
    # cost the line its headline -- after $BENCH_SHARD_BUDGET_S (default 420 s) every rank gives up on them, rank 0 prints the line
    # with what has been measured, and the process exits 0.
    done = threading.Event()

    def give_up():
        if done.is_set():
            return
        log("rank %d: sharded modes exceeded their time budget -- reporting what was measured" % rank)
        modes.setdefault("aborted", "the sharded modes did not finish within BENCH_SHARD_BUDGET_S")
        try:
            final_line()
        finally:
            os._exit(0)
    watchdog = threading.Timer(float(os.environ.get("BENCH_SHARD_BUDGET_S", "420")), give_up)
    watchdog.daemon = True
    watchdog.start()
    try:
        if want in ("all", "sharded"):
            modes["doc_sharded"] = sharded_mode(world, "doc_sharded")
            log("rank %d doc_sharded: %.1f M q/s" % (rank, modes["doc_sharded"]["value"] / 1e6))
        if want in ("all", "hybrid") and world >= 4 and world % 2 == 0:
            modes["hybrid"] = sharded_mode(2, "hybrid")
            log("rank %d hybrid: %.1f M q/s" % (rank, modes["hybrid"]["value"] / 1e6))
    except Exception as e:  # noqa: BLE001 -- a failed sharded mode is reported, not fatal: the replicas figure stands
        log("rank %d: sharded mode failed: %r" % (rank, e))
        modes["error"] = repr(e)[:300]
        done.set()
        final_line()
        os._exit(0)  # the CUDA context / the process group may be unusable: no orderly teardown
    done.set()
    watchdog.cancel()
    final_line()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--tokens", type=int, default=100_000_000)
    ap.add_argument("--big-tokens", type=int, default=1_000_000_000)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--cpu-batch", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--mode", default="all", choices=["all", "replicas", "sharded", "hybrid"])
    ap.add_argument("--layout", default="resident", choices=["resident", "sampled"])
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    # the one JSON line must be the only thing on stdout: NCCL / library banners that print to fd 1 are sent to stderr
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
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
        try:
            run_multi(args, device, local, rank, world, dist)
        finally:
            dist.destroy_process_group()
    else:
        run_single(args, device, local)


if __name__ == "__main__":
    main()
