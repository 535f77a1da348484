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
        """Fused step for the rows this rank owns: logits_rows [B/world, V] are the logits of beams
        [rank*B/world, (rank+1)*B/world).  sp/ep are this shard's ranges of ALL B beams.
        The rows owned by OTHER ranks are enumerated into bitmaps and exchanged (all-to-all); the
        rows owned by this rank are enumerated inside the fused kernel, which ORs the pieces
        received from the other shards while it streams the logits."""
        B = sp.numel()
        own = B // self.world
        V = logits_rows.shape[1]
        words = (V + 31) // 32
        lo, hi = self.rank * own, (self.rank + 1) * own
        send = torch.zeros((B, words), dtype=torch.int32, device=sp.device)
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

    def locate_topk(self, sp, ep, k):
        """First k rows of every range located on this shard, all-gathered as global positions."""
        rows = sp.unsqueeze(1) + torch.arange(k, device=sp.device, dtype=sp.dtype).unsqueeze(0)
        valid = rows < ep.unsqueeze(1)
        pos = self.index.locate_batch(torch.where(valid, rows, torch.zeros_like(rows)).contiguous().view(-1)).view(-1, k)
        return gather_hits(pos, valid, self.pos_base, self.group)


class _RawCuda:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can view it."""

    def __init__(self, ptr, n_words):
        self.__cuda_array_interface__ = {"shape": (n_words,), "typestr": "<i4", "data": (int(ptr), False), "version": 2}


class PeerMaskExchange:
    """OR-reduce-scatter of the per-shard bitmaps by PULLING over NVLink peer memory -- no NCCL in the loop.

    The dense all-to-all moves 16 KB per prefix per shard although a prefix with a few tokens of history has a
    handful of continuations.  Here every rank keeps its [B, words] bitmaps in a peer-visible buffer (CUDA IPC,
    fmidx_ipc_*) together with one occupancy word per row (bit s: 4096-token slice s may be non-empty, produced
    by the mask kernels themselves).  Per step (all kernels of this library, data and flags written straight
    into peer memory):
      publish  this shard's occupancy words + occurrence counts -> slot `rank` of every shard's inbox, then
               flag[rank] = step on every shard                                (fmidx_xchg_publish)
      wait     a one-CTA kernel spins until every shard's flag reached this step   (fmidx_xchg_wait)
      pull     for the rows this rank owns, read only the occupied slices out of the peers' bitmaps and OR
               them                                                              (fmidx_mask_pull_or)
      ack      ack[rank] = step on every shard                                  (fmidx_xchg_signal)
    `n_buffers` buffer sets, each with its own flags and step counter: the sets are independent, so set i's kernels
    and exchange can run back to back on "its" stream while the other sets' steps overlap them on theirs (bench.py:
    4 sets, 4 streams).  Before a set is overwritten, compute_local waits for every shard's ack of the step that
    last used it.
    """

    def __init__(self, index, B, vocab, n_buffers=3, group=None):
        import ctypes as C
        from . import _native as N
        self.index, self.group = index, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        assert B % self.world == 0
        self.B, self.vocab, self.own = B, vocab, B // self.world
        self.words = (vocab + 31) // 32
        self.ow = ((vocab + 4095) // 4096 + 31) // 32
        self.n = n_buffers
        self.dev = dev = torch.device("cuda", index.device_index)
        self._local, self._peers = [], []
        # side data of one shard for one step: occupancy words, then the occurrence counts (ep - sp) as int64
        self.xwords = B * self.ow + 2 * B
        self.mask_ptr, self.mask_tables, self.local_views = [], [], []
        self.inbox_ptr, self.inbox_tables, self.inbox_views = [], [], []
        for _ in range(n_buffers):
            ptr, table = self._shared(B * self.words * 4)
            self.mask_ptr.append(ptr)
            self.mask_tables.append(table)
            self.local_views.append(torch.as_tensor(_RawCuda(ptr, B * self.words), device=dev).view(B, self.words))
            ptr, table = self._shared(self.world * self.xwords * 4)
            self.inbox_ptr.append(ptr)
            self.inbox_tables.append(table)
            self.inbox_views.append(torch.as_tensor(_RawCuda(ptr, self.world * self.xwords), device=dev).view(self.world, self.xwords))
        # every buffer set has its OWN data / ack flags and step counter, so the exchanges of different sets are
        # independent and may run on different streams (set i: flags[i][0:world] data of step s published by rank r,
        # flags[i][world:2*world] step s pulled by rank r)
        fptr, ftable = self._shared(n_buffers * 2 * self.world * 4)
        self.flags = torch.as_tensor(_RawCuda(fptr, n_buffers * 2 * self.world), device=dev)
        self.flags.zero_()
        self.data_flag_tables = [ftable + 8 * self.world * i for i in range(n_buffers)]
        self.ack_flag_tables = [ftable + 8 * self.world * i + 4 * self.world for i in range(n_buffers)]
        self.xlocal = [torch.empty(self.xwords, dtype=torch.int32, device=dev) for _ in range(n_buffers)]
        self.counts = [torch.empty(B, dtype=torch.int64, device=dev) for _ in range(n_buffers)]
        self.own_views = [v[self.rank * self.own:(self.rank + 1) * self.own] for v in self.local_views]
        self.seq = [0] * n_buffers
        self.last_seq = [0] * n_buffers
        # ctypes views of the fixed pointers (built once: two host calls per step is all the Python there is)
        self._c_mask = [C.c_void_p(p) for p in self.mask_ptr]
        self._c_inbox = [C.c_void_p(p) for p in self.inbox_ptr]
        self._c_xlocal = [C.c_void_p(t.data_ptr()) for t in self.xlocal]
        self._c_counts = [C.c_void_p(t.data_ptr()) for t in self.counts]
        self._c_mask_tables = [C.c_void_p(t.data_ptr()) for t in self.mask_tables]
        self._c_inbox_tables = [C.c_void_p(t.data_ptr()) for t in self.inbox_tables]
        self._c_flags = [C.c_void_p(self.flags.data_ptr() + 8 * self.world * i) for i in range(n_buffers)]
        self._c_ack = [C.c_void_p(self.flags.data_ptr() + 8 * self.world * i + 4 * self.world) for i in range(n_buffers)]
        self._c_data_flag_table = [C.c_void_p(t.data_ptr()) for t in self.data_flag_tables]
        self._c_ack_flag_table = [C.c_void_p(t.data_ptr()) for t in self.ack_flag_tables]
        torch.cuda.synchronize()
        dist.barrier(group=group)

    def _shared(self, nbytes):
        """Allocate a peer-visible buffer; returns (local pointer, device table of every rank's pointer to its copy)."""
        import ctypes as C
        from . import _native as N
        L = N.lib()
        ptr = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        N.check(L.fmidx_ipc_alloc(self.index._h, nbytes, C.byref(ptr), handle))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=self.group)
        ptrs = []
        for r in range(self.world):
            if r == self.rank:
                ptrs.append(ptr.value)
            else:
                q = C.c_void_p()
                hb = (C.c_ubyte * 64).from_buffer_copy(handles[r])
                N.check(L.fmidx_ipc_open(self.index._h, hb, C.byref(q)))
                ptrs.append(q.value)
                self._peers.append(q.value)
        self._local.append(ptr.value)
        return ptr.value, torch.tensor(ptrs, dtype=torch.int64, device=self.dev)

    @staticmethod
    def _s():
        import ctypes as C
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def compute_local(self, i, sp, ep, shift=10, ws_slot=-1, stream=None):
        """K2 of this shard for all B prefixes into buffer set i (bitmaps + occupancy words + counts); first waits
        (on the GPU) until every shard has pulled the step that last used set i.  ws_slot: the enumeration's scratch
        workspace (give every stream that may run concurrently its own, 0..7).  stream: a torch stream (default: the
        current one; passing it avoids a stream context switch per step -- the host is the bottleneck here)."""
        import ctypes as C
        from . import _native as N
        sptr = self._s() if stream is None else C.c_void_p(stream.cuda_stream)
        N.check(N.lib().fmidx_xchg_local(
            self.index._h, C.c_void_p(sp.data_ptr()), C.c_void_p(ep.data_ptr()), self.B, shift, self.vocab,
            self._c_mask[i], self._c_xlocal[i], self._c_ack[i], self.world, self.rank, self.last_seq[i], ws_slot, sptr))

    def exchange(self, i, stream=None):
        """publish -> wait -> pull + OR (in place) -> total counts -> ack for buffer set i, ONE host call on `stream`
        (default: the current one).  Returns the final [B/world, words] bitmaps of the rows this rank owns (a view
        of set i)."""
        import ctypes as C
        from . import _native as N
        sptr = self._s() if stream is None else C.c_void_p(stream.cuda_stream)
        self.seq[i] += 1
        N.check(N.lib().fmidx_xchg_round(
            self.index._h, self._c_mask_tables[i], self._c_inbox_tables[i], self._c_data_flag_table[i], self._c_ack_flag_table[i],
            self._c_flags[i], self._c_inbox[i], self._c_xlocal[i],
            self.world, self.rank, self.B, self.vocab, self.seq[i], self._c_counts[i], sptr))
        self.last_seq[i] = self.seq[i]
        return self.own_views[i]

    def total_counts(self, i):
        """Sum over shards of the occurrence counts of all B prefixes (they rode in the published side data and
        were summed by exchange())."""
        return self.counts[i]

    def close(self):
        import ctypes as C
        from . import _native as N
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for q in self._peers:
            N.lib().fmidx_ipc_close(self.index._h, C.c_void_p(q))
        dist.barrier(group=self.group)
        for p in self._local:
            N.lib().fmidx_ipc_free(self.index._h, C.c_void_p(p))
        self._peers, self._local = [], []
