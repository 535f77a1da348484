"""Document-sharded FM-index over the GPUs of one box: one process per GPU, each rank holds the
index of its own document shard.  Prefix batches are replicated; ranges are
shard-local; what crosses NVLink per step is

  * masks  : per-shard allowed-token bitmaps, OR-reduced.  NCCL has no bitwise-OR reduction, so the
             reduction is an all-to-all of bitmap row-blocks followed by a local OR: rank r ends up
             owning the final masks of rows [r*B/W, (r+1)*B/W) -- which is where the fused logits
             kernel consumes them (fmidx_mask_logits ORs the gathered pieces while it streams the
             logits rows, so there is no separate OR pass on that path);
  * counts : occurrence counts (ep - sp) per prefix, all-reduced (sum);
  * hits   : located positions / document ids, all-gathered with per-shard bases added.

The reference has no distributed path; parity for the sharded configuration is "one reference
index per shard, results OR/sum-reduced on the host" (sharding removes the cross-shard document
adjacencies of the monolithic concatenation).  The collectives below use torch.distributed only
(NCCL on GPUs; the same code runs on gloo/CPU tensors for the world_size-2 tests).
"""
import torch
import torch.distributed as dist


def shard_documents(beginnings, world):
    """Contiguous document groups balanced by token count.
    beginnings: cumulative doc offsets (len m+1).  Returns [(doc_lo, doc_hi)] * world."""
    beg = [int(x) for x in beginnings]
    m, n = len(beg) - 1, beg[-1]
    cuts = [0]
    for r in range(1, world):
        target = n * r / world
        lo, hi = cuts[-1], m
        while lo < hi:  # first doc boundary >= target
            mid = (lo + hi) // 2
            if beg[mid] < target:
                lo = mid + 1
            else:
                hi = mid
        cuts.append(min(max(lo, cuts[-1]), m))
    cuts.append(m)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def _or_pieces(pieces, index=None):
    """OR over dim 0.  CUDA tensors go through the library's own kernel (fmidx_mask_or); CPU tensors
    (the gloo tests) through torch."""
    if pieces.is_cuda and index is not None and pieces[0].numel() % 4 == 0:
        import ctypes as C
        from . import _native as N
        out = torch.empty_like(pieces[0])
        N.check(N.lib().fmidx_mask_or(index._h, C.c_void_p(pieces.data_ptr()), pieces.shape[0], pieces[0].numel(),
                                      C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out
    out = pieces[0].clone()
    for s in range(1, pieces.shape[0]):
        out |= pieces[s]
    return out


def or_reduce_scatter(mask_local, group=None, index=None):
    """mask_local [B, words] int32 (this shard's bitmaps for ALL B prefixes), B % world == 0.
    Returns (pieces [world, B/world, words], own) where pieces[s] is shard s's bitmap of the rows
    this rank owns and own = OR over s."""
    world = dist.get_world_size(group)
    B, words = mask_local.shape
    assert B % world == 0, "pad the batch to a multiple of the world size"
    pieces = torch.empty((world, B // world, words), dtype=mask_local.dtype, device=mask_local.device)
    dist.all_to_all_single(pieces.view(-1), mask_local.contiguous().view(-1), group=group)
    return pieces, _or_pieces(pieces, index)


def or_all(mask_local, group=None, index=None):
    """Every rank gets the OR over shards of all B rows (all-gather + OR)."""
    world = dist.get_world_size(group)
    gathered = torch.empty((world,) + tuple(mask_local.shape), dtype=mask_local.dtype, device=mask_local.device)
    dist.all_gather_into_tensor(gathered.view(-1), mask_local.contiguous().view(-1), group=group)
    return _or_pieces(gathered, index)


def sum_counts(sp, ep, group=None):
    """Total occurrence count of every prefix over all shards (all-reduce sum of ep - sp)."""
    cnt = (ep - sp).to(torch.int64)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    return cnt


def gather_hits(pos_local, valid_local, pos_base, group=None):
    """pos_local [Q, K] int64 local text positions of the first K rows of every range, valid_local
    [Q, K] bool; pos_base: this shard's offset in the global concatenation.  Returns
    (positions [world, Q, K] global, valid [world, Q, K]) on every rank."""
    world = dist.get_world_size(group)
    g = torch.where(valid_local, pos_local + pos_base, torch.full_like(pos_local, -1))
    out = torch.empty((world,) + tuple(g.shape), dtype=g.dtype, device=g.device)
    dist.all_gather_into_tensor(out.view(-1), g.contiguous().view(-1), group=group)
    return out, out >= 0


class ShardedFMIndex:
    """Rank-local view of a document-sharded index.

    index      : this rank's FMIndex (built over documents [doc_lo, doc_hi) of the corpus)
    doc_base   : global id of the shard's first document;  pos_base: its first token's global offset
    """

    def __init__(self, index, doc_base=0, pos_base=0, group=None):
        self.index = index
        self.doc_base = int(doc_base)
        self.pos_base = int(pos_base)
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._send = None
        self._rng = None

    # ranges are shard-local and never reduced
    def get_range_batch(self, toks, offs):
        return self.index.get_range_batch(toks, offs)

    def count_batch(self, sp, ep):
        return sum_counts(sp, ep, self.group)

    def distinct_mask_batch(self, sp, ep, vocab, everywhere=False):
        """OR over shards of the allowed-token bitmaps.  everywhere=False: this rank's B/world rows."""
        local = self.index.distinct_mask_batch(sp, ep, vocab)
        if everywhere:
            return or_all(local, self.group, self.index)
        return or_reduce_scatter(local, self.group, self.index)[1]

    def mask_logits_rows(self, sp, ep, logits_rows, override_tok=None, always_allow=-1):
        """Fused step for the rows this rank owns, from ranges (NCCL all-to-all form; CPU-testable plumbing):
        logits_rows [B/world, V] are the logits of beams [rank*B/world, (rank+1)*B/world), sp/ep this shard's ranges
        of ALL B beams.  The rows owned by OTHER ranks are enumerated into bitmaps and exchanged; the rows owned by
        this rank are enumerated inside the fused kernel, which ORs the pieces received from the other shards while
        it streams the logits.  On GPUs prefer decode_step(): same result through the peer-memory exchange."""
        B = sp.numel()
        own = B // self.world
        V = logits_rows.shape[1]
        words = (V + 31) // 32
        lo, hi = self.rank * own, (self.rank + 1) * own
        if self._send is None or self._send.shape != (B, words) or self._send.device != sp.device:
            self._send = torch.zeros((B, words), dtype=torch.int32, device=sp.device)  # own rows stay zero for good
        send = self._send
        if lo > 0:
            self.index.distinct_mask_batch(sp[:lo].contiguous(), ep[:lo].contiguous(), V, out=send[:lo])
        if hi < B:
            self.index.distinct_mask_batch(sp[hi:].contiguous(), ep[hi:].contiguous(), V, out=send[hi:])
        pieces = torch.empty((self.world, own, words), dtype=torch.int32, device=sp.device)
        dist.all_to_all_single(pieces.view(-1), send.view(-1), group=self.group)
        others = torch.cat([pieces[:self.rank], pieces[self.rank + 1:]]).contiguous() if self.world > 1 else None
        ovr = override_tok[lo:hi].contiguous() if override_tok is not None else None
        return self.index.mask_logits(sp[lo:hi].contiguous(), ep[lo:hi].contiguous(), logits_rows,
                                      override_tok=ovr, always_allow=always_allow, extra_masks=others)

    def decode_step(self, step, set_index, toks, offs, logits_rows, override_tok=None, always_allow=-1, stream=None):
        """The fused constrained-decoding step of a document-sharded index over the peer-memory exchange (no NCCL, no
        per-step allocation): `step` is a ShardStep for B beams; toks/offs the B replicated beam histories;
        logits_rows [B/world, V] the logits of the beams this rank owns.  One sharded step (search + enumeration +
        exchange -> final bitmaps of the owned rows), then the logits kernel applies them (x + 0.0 / x + -inf,
        override rows, always-allowed token) while it streams the rows.  Returns (masked rows, total counts [B])."""
        own, lo = step.own, step.rank * step.own
        if self._rng is None or self._rng[0].numel() != step.B or self._rng[0].device != logits_rows.device:
            self._rng = (torch.empty(step.B, dtype=torch.int64, device=logits_rows.device),
                         torch.empty(step.B, dtype=torch.int64, device=logits_rows.device),
                         torch.zeros(own, dtype=torch.int64, device=logits_rows.device))
        sp, ep, zero = self._rng
        s = stream if stream is not None else step.streams[set_index]
        masks, counts = step.step(set_index, toks, offs, sp, ep, stream=s)
        ovr = override_tok[lo:lo + own].contiguous() if override_tok is not None else None
        with torch.cuda.stream(s):
            out = self.index.mask_logits(zero, zero, logits_rows, override_tok=ovr, always_allow=always_allow,
                                         extra_masks=masks.unsqueeze(0))
        return out, counts

    def locate_topk(self, sp, ep, k):
        """First k rows of every range located on this shard, all-gathered as global positions."""
        rows = sp.unsqueeze(1) + torch.arange(k, device=sp.device, dtype=sp.dtype).unsqueeze(0)
        valid = rows < ep.unsqueeze(1)
        pos = self.index.locate_batch(torch.where(valid, rows, torch.zeros_like(rows)).contiguous().view(-1)).view(-1, k)
        return gather_hits(pos, valid, self.pos_base, self.group)


class _RawCuda:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can view it."""

    def __init__(self, ptr, n, typestr="<i4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class ShardStep:
    """One shard's side of the document-sharded step, over the C context `fmidx_shard_*` (include/fmidx.h).

    step(set, toks, offs) enqueues ONE whole step -- search of all B replicated prefixes on this shard, enumeration
    into peer-visible bitmaps (+ slice-occupancy words and occurrence counts), publication into every shard's inbox
    at the tail of the last enumeration kernel, and the fused wait + pull + OR + total-counts + ack kernel -- with one
    host call and six launches (or ONE graph launch: graph_prepare / graph_launch), no NCCL.  Afterwards rank r owns the final bitmaps of prefixes
    [r*B/world, (r+1)*B/world) and the total occurrence counts of all B prefixes.  `n_sets` independent buffer sets:
    steps on different sets may be in flight on different streams; every rank must issue the same steps on the same
    sets in the same order.  `group`: the process group whose ranks hold the shards (hybrid: one group per replica).
    """

    def __init__(self, index, B, vocab, n_sets=4, group=None):
        import ctypes as C
        from . import _native as N
        self.index, self.group = index, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        assert B % self.world == 0, "pad the batch to a multiple of the number of shards"
        self.B, self.vocab, self.own, self.n = B, vocab, B // self.world, n_sets
        self.words = (vocab + 31) // 32
        self.dev = dev = torch.device("cuda", index.device_index)
        L = self._L = N.lib()
        x = C.c_void_p()
        N.check(L.fmidx_shard_create(index._h, self.world, self.rank, B, vocab, n_sets, C.byref(x)))
        self._x = x
        # handle exchange (plumbing): kind 0 = bitmap buffer, 1 = inbox (two of each per set: steps ping-pong), 2 = the flag block
        mine = []
        for kind, sets in ((0, range(2 * n_sets)), (1, range(2 * n_sets)), (2, (0,))):
            for i in sets:
                hd = (C.c_ubyte * 64)()
                N.check(L.fmidx_shard_export(x, i, kind, hd))
                mine.append((kind, i, bytes(hd)))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        for r, items in enumerate(everyone):
            if r == self.rank:
                continue
            for kind, i, hb in items:
                N.check(L.fmidx_shard_import(x, i, kind, r, (C.c_ubyte * 64).from_buffer_copy(hb)))
        N.check(L.fmidx_shard_ready(x))
        self.own_masks, self.counts, self.streams, self._ptrs = [], [], [], []
        for i in range(n_sets):
            views = []
            self._ptrs.append([])
            for par in (0, 1):  # step k of a set writes bitmap buffer k & 1
                pm, pc, ps = C.c_void_p(), C.c_void_p(), C.c_void_p()
                N.check(L.fmidx_shard_result(x, i, par, C.byref(pm), C.byref(pc), C.byref(ps)))
                views.append(torch.as_tensor(_RawCuda(pm.value, self.own * self.words), device=dev).view(self.own, self.words))
                self._ptrs[i].append(pm.value)
            self.own_masks.append(views)
            self.counts.append(torch.as_tensor(_RawCuda(pc.value, B, "<i8"), device=dev))
            self.streams.append(torch.cuda.ExternalStream(ps.value, device=dev))
        torch.cuda.synchronize()
        dist.barrier(group=group)

    def step(self, i, toks, offs, sp, ep, shift=10, stream=None):
        """Enqueue the step of buffer set i (stream: a torch stream, default the set's own).  sp/ep (int64 [B], CUDA)
        receive this shard's ranges.  Returns (final bitmaps of the owned rows [B/world, words], total counts [B]); the
        bitmaps stay valid until the set's second-next step (the sets ping-pong between two buffers), the counts until
        its next step."""
        import ctypes as C
        from . import _native as N
        sptr = C.c_void_p(stream.cuda_stream) if stream is not None else None
        N.check(self._L.fmidx_shard_step(self._x, i, C.c_void_p(toks.data_ptr()), C.c_void_p(offs.data_ptr()), shift,
                                         C.c_void_p(sp.data_ptr()), C.c_void_p(ep.data_ptr()), sptr))
        return self.result(i)

    def graph_prepare(self, i, toks, offs, sp, ep, shift=10):
        """Capture the step of buffer set i for these FIXED device buffers (toks / offs are rewritten in place between launches;
        sp / ep receive this shard's ranges).  Runs one real step: every rank must prepare the same sets in the same order."""
        import ctypes as C
        from . import _native as N
        N.check(self._L.fmidx_shard_graph_prepare(self._x, i, C.c_void_p(toks.data_ptr()), C.c_void_p(offs.data_ptr()), shift,
                                                  C.c_void_p(sp.data_ptr()), C.c_void_p(ep.data_ptr())))
        self._keep = getattr(self, "_keep", []) + [(toks, offs, sp, ep)]

    def graph_launch(self, i):
        """Replay set i's captured step (ONE graph launch: search, enumeration, publication, pull) on the set's stream;
        result(i) gives the buffers it fills."""
        from . import _native as N
        if self._L.fmidx_shard_graph_launch(self._x, i) != 0:
            N.check(-1)
        return self.result(i)

    def result(self, i):
        """(owned bitmaps, total counts) of the step issued last on set i (also after a direct fmidx_shard_step call)."""
        import ctypes as C
        from . import _native as N
        pm = C.c_void_p()
        N.check(self._L.fmidx_shard_result(self._x, i, -1, C.byref(pm), None, None))
        return self.own_masks[i][self._ptrs[i].index(pm.value)], self.counts[i]

    def submit(self, i, toks, offs, sp, ep, counts=None, n_distinct_own=None, shift=10):
        """Host-buffer form on the set's own stream (pinned CPU tensors in and out); wait(i) blocks until they hold the
        results."""
        import ctypes as C
        from . import _native as N
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
        N.check(self._L.fmidx_shard_submit(self._x, i, p(toks), p(offs), shift, p(sp), p(ep), p(counts), p(n_distinct_own)))

    def wait(self, i):
        from . import _native as N
        N.check(self._L.fmidx_shard_wait(self._x, i))

    def close(self):
        if self._x is None:
            return
        torch.cuda.synchronize()
        dist.barrier(group=self.group)  # nobody is still pulling from anybody
        self.own_masks, self.counts = [], []
        self._L.fmidx_shard_destroy(self._x)
        self._x = None
        dist.barrier(group=self.group)
