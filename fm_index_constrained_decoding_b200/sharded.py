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
