"""Multi-GPU parity worker of the document-sharded path (launched by tests/test_multigpu.py, or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tests/multigpu_check.py

Every rank builds the index of its document shard on its GPU; the expectation is one CPU oracle per shard with
results OR/sum-reduced on the host (sharded parity).  The peer-memory step is driven exactly like bench.py
drives it: 4 buffer sets, each on its own stream, different prefix batches rotating, every set wrapped around >= 3
times before anything is checked, results copied out per step and compared afterwards.  Also: hybrid groups
(shards x replicas) when the world size allows, the fused decode step over the exchange, located hits."""
import os
import random
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

V, EOS = 128256, 128009
WORDS = (V + 31) // 32


def make_docs(seed=3, n_docs=400):
    rng = random.Random(seed)
    docs = [[rng.randrange(3000) for _ in range(rng.randint(10, 200))] + [EOS] for _ in range(n_docs)]
    for k in range(n_docs // 4):
        docs[n_docs - n_docs // 4 + k] = docs[k][3:40] + docs[n_docs - n_docs // 4 + k]
    return docs, rng


def make_batch(docs, rng, B):
    seqs = []
    for _ in range(B):
        d = rng.choice(docs)
        i = rng.randrange(len(d) - 1)
        seqs.append(d[i:i + rng.randint(1, 8)])
    return seqs


def expected(docs, parts, beg, batches):
    """Per-shard oracle results, OR / sum reduced on the host."""
    from oracle.oracle import OracleFMIndex
    from fm_index_constrained_decoding_b200.index import ragged
    oracles = []
    for a, b in parts:
        o = OracleFMIndex()
        o.initialize(docs[a:b])
        oracles.append((o, beg[a]))
    out = []
    for seqs in batches:
        toks, offs = ragged(seqs)
        B = len(seqs)
        mask, cnt, per = np.zeros((B, WORDS), np.uint32), np.zeros(B, np.int64), []
        for o, _ in oracles:
            osp, oep, om, _ = o.query_batch(toks, offs, want_mask=True, vocab=V, nthreads=4)
            mask |= om
            cnt += (oep - osp).astype(np.int64)
            per.append((osp, oep, om))
        out.append((mask, cnt, per))
    return out, oracles


def check_group(ix_docs_parts, group, grank, gworld, tag, n_sets=4, rounds=4, B_per=64):
    """Drive ShardStep like bench.py (n_sets sets x n_sets streams, `rounds` wrap-arounds) inside `group`."""
    from fm_index_constrained_decoding_b200 import sharded
    from fm_index_constrained_decoding_b200.index import ragged
    ix, docs, parts, beg = ix_docs_parts
    B = B_per * gworld
    own = B // gworld
    rng = random.Random(1000 + 7 * gworld + (0 if group is None else 1))
    n_batches = 6
    batches = [make_batch(docs, rng, B) for _ in range(n_batches)]
    exp, _ = expected(docs, parts, beg, batches)
    dev = torch.device("cuda", ix.device_index)
    dev_batches = []
    for seqs in batches:
        t, o = ragged(seqs)
        dev_batches.append((torch.from_numpy(t).to(dev), torch.from_numpy(o).to(dev)))
    st = sharded.ShardStep(ix, B, V, n_sets=n_sets, group=group)
    rng_bufs = [(torch.empty(B, dtype=torch.int64, device=dev), torch.empty(B, dtype=torch.int64, device=dev)) for _ in range(n_sets)]
    steps = n_sets * rounds
    got_mask = torch.empty((steps, own, WORDS), dtype=torch.int32, device=dev)
    got_cnt = torch.empty((steps, B), dtype=torch.int64, device=dev)
    got_sp = torch.empty((steps, B), dtype=torch.int64, device=dev)
    got_ep = torch.empty((steps, B), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    for k in range(steps):
        i = k % n_sets
        s = st.streams[i]
        toks, offs = dev_batches[k % n_batches]
        sp, ep = rng_bufs[i]
        masks, counts = st.step(i, toks, offs, sp, ep)  # the set's own stream
        with torch.cuda.stream(s):  # copy the results out before the set is reused
            got_mask[k].copy_(masks, non_blocking=True)
            got_cnt[k].copy_(counts, non_blocking=True)
            got_sp[k].copy_(sp, non_blocking=True)
            got_ep[k].copy_(ep, non_blocking=True)
    torch.cuda.synchronize()
    gm, gc = got_mask.cpu().numpy().view(np.uint32), got_cnt.cpu().numpy()
    gsp, gep = got_sp.cpu().numpy().astype(np.uint64), got_ep.cpu().numpy().astype(np.uint64)
    for k in range(steps):
        mask, cnt, per = exp[k % n_batches]
        assert (gsp[k] == per[grank][0]).all() and (gep[k] == per[grank][1]).all(), "%s: ranges differ at step %d" % (tag, k)
        assert (gm[k] == mask[grank * own:(grank + 1) * own]).all(), "%s: owned bitmaps differ at step %d" % (tag, k)
        assert (gc[k] == cnt).all(), "%s: total counts differ at step %d" % (tag, k)
    # host-buffer form on the same sets
    htoks = [(torch.from_numpy(ragged(b)[0]).pin_memory(), torch.from_numpy(ragged(b)[1]).pin_memory()) for b in batches]
    hout = [(torch.empty(B, dtype=torch.int64).pin_memory(), torch.empty(B, dtype=torch.int64).pin_memory(),
             torch.empty(B, dtype=torch.int64).pin_memory(), torch.empty(own, dtype=torch.int32).pin_memory()) for _ in range(n_sets)]
    for k in range(2 * n_sets):
        i = k % n_sets
        st.wait(i)
        if k >= n_sets:
            mask, cnt, per = exp[(k - n_sets) % n_batches]
            hsp, hep, hc, hn = hout[i]
            assert (hsp.numpy().astype(np.uint64) == per[grank][0]).all() and (hc.numpy() == cnt).all()
            nd = np.unpackbits(mask[grank * own:(grank + 1) * own].view(np.uint8), axis=1).sum(1)
            assert (hn.numpy() == nd).all(), "%s: distinct-token counts differ" % tag
        st.submit(i, htoks[k % n_batches][0], htoks[k % n_batches][1], *hout[i])
    for i in range(n_sets):
        st.wait(i)
    # captured form: every set's step recorded once for fixed device buffers whose CONTENT changes between launches, one graph
    # launch per step, mixed with eager steps on the same sets (both advance the same device-side step number)
    maxtok = max(t.numel() for t, _ in dev_batches)
    gbufs = [(torch.zeros(maxtok, dtype=torch.int32, device=dev), torch.zeros(B + 1, dtype=torch.int64, device=dev),
              torch.empty(B, dtype=torch.int64, device=dev), torch.empty(B, dtype=torch.int64, device=dev)) for _ in range(n_sets)]
    for i in range(n_sets):
        st.graph_prepare(i, *gbufs[i])
    torch.cuda.synchronize()
    steps2 = n_sets * 3
    for k in range(steps2):
        i = k % n_sets
        s = st.streams[i]
        toks, offs = dev_batches[(k + 1) % n_batches]
        gt, go, gsp_b, gep_b = gbufs[i]
        with torch.cuda.stream(s):
            gt[:toks.numel()].copy_(toks, non_blocking=True)
            go.copy_(offs, non_blocking=True)
        if k % 5 == 3:
            masks, counts = st.step(i, gt, go, gsp_b, gep_b)  # an eager step in between
        else:
            masks, counts = st.graph_launch(i)
        with torch.cuda.stream(s):
            got_mask[k].copy_(masks, non_blocking=True)
            got_cnt[k].copy_(counts, non_blocking=True)
            got_sp[k].copy_(gsp_b, non_blocking=True)
    torch.cuda.synchronize()
    gm, gc = got_mask.cpu().numpy().view(np.uint32), got_cnt.cpu().numpy()
    gsp = got_sp.cpu().numpy().astype(np.uint64)
    for k in range(steps2):
        mask, cnt, per = exp[(k + 1) % n_batches]
        assert (gsp[k] == per[grank][0]).all(), "%s: captured step %d: ranges differ" % (tag, k)
        assert (gm[k] == mask[grank * own:(grank + 1) * own]).all(), "%s: captured step %d: owned bitmaps differ" % (tag, k)
        assert (gc[k] == cnt).all(), "%s: captured step %d: total counts differ" % (tag, k)
    # fused decode step over the exchange (rows this rank owns)
    sh = sharded.ShardedFMIndex(ix, group=group)
    g = torch.Generator().manual_seed(grank)
    logits = torch.randn(own, V, generator=g)
    ovr = torch.full((B,), -1, dtype=torch.int32)
    ovr[grank * own + 1] = 77  # one override row among the owned ones
    out, counts = sh.decode_step(st, 0, dev_batches[0][0], dev_batches[0][1], logits.to(dev), override_tok=ovr.to(dev), always_allow=EOS)
    torch.cuda.synchronize()
    mask, cnt, _ = exp[0]
    bits = np.unpackbits(mask[grank * own:(grank + 1) * own].view(np.uint8), axis=1, bitorder="little")[:, :V].astype(bool)
    bits[1, :] = False
    bits[1, 77] = True
    bits[:, EOS] = True
    want = torch.where(torch.from_numpy(bits), logits + 0.0, torch.full_like(logits, float("-inf")))
    assert torch.equal(out.cpu(), want), "%s: fused decode rows differ" % tag
    assert (counts.cpu().numpy() == cnt).all()
    st.close()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ["FMIDX_DEVICE"] = str(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from fm_index_constrained_decoding_b200 import FMIndex, sharded
    from fm_index_constrained_decoding_b200.index import ragged

    docs, rng = make_docs()
    beg = [0]
    for d in docs:
        beg.append(beg[-1] + len(d))
    # ---- all ranks = shards -------------------------------------------------------------------------
    parts = sharded.shard_documents(beg, world)
    lo, hi = parts[rank]
    ix = FMIndex(device=local)
    ix.initialize(docs[lo:hi], in_memory=True)
    check_group((ix, docs, parts, beg), None, rank, world, "shards=%d" % world)
    # ---- NCCL plumbing of ShardedFMIndex (all-to-all + OR, counts, fused rows, hits) ------------------------
    sh = sharded.ShardedFMIndex(ix, doc_base=lo, pos_base=beg[lo])
    B = 64 * world
    own = B // world
    seqs = make_batch(docs, rng, B)
    exp1, oracles = expected(docs, parts, beg, [seqs])
    mask, cnt, per = exp1[0]
    toks, offs = ragged(seqs)
    dt, do = torch.from_numpy(toks).cuda(), torch.from_numpy(offs).cuda()
    sp, ep = sh.get_range_batch(dt, do)
    assert (sp.cpu().numpy().astype(np.uint64) == per[rank][0]).all()
    m_own = sh.distinct_mask_batch(sp, ep, V)
    assert (m_own.cpu().numpy().view(np.uint32) == mask[rank * own:(rank + 1) * own]).all()
    m_all = sh.distinct_mask_batch(sp, ep, V, everywhere=True)
    assert (m_all.cpu().numpy().view(np.uint32) == mask).all()
    assert (sh.count_batch(sp, ep).cpu().numpy() == cnt).all()
    g = torch.Generator().manual_seed(rank)
    logits = torch.randn(own, V, generator=g)
    for _ in range(2):  # twice: the send buffer is reused
        out = sh.mask_logits_rows(sp, ep, logits.cuda(), always_allow=EOS).cpu()
    bits = np.unpackbits(mask[rank * own:(rank + 1) * own].view(np.uint8), axis=1, bitorder="little")[:, :V].astype(bool)
    bits[:, EOS] = True
    assert torch.equal(out, torch.where(torch.from_numpy(bits), logits + 0.0, torch.full_like(logits, float("-inf"))))
    K = 4
    hits, hv = sh.locate_topk(sp, ep, K)
    hits, hv = hits.cpu().numpy(), hv.cpu().numpy()
    for s, ((osp, oep, _), (o, base)) in enumerate(zip(per, oracles)):
        for qi in range(B):
            for k in range(K):
                r = int(osp[qi]) + k
                if r < int(oep[qi]):
                    assert hv[s, qi, k] and hits[s, qi, k] == o.locate(r) + base
                else:
                    assert not hv[s, qi, k]
    # ---- hybrid: groups of 2 shards x world/2 replica groups ------------------------------------------------
    if world >= 4 and world % 2 == 0:
        g_size = 2
        groups = [dist.new_group(list(range(a, a + g_size))) for a in range(0, world, g_size)]
        mygroup = groups[rank // g_size]
        grank = rank % g_size
        gparts = sharded.shard_documents(beg, g_size)
        glo, ghi = gparts[grank]
        gix = FMIndex(device=local)
        gix.initialize(docs[glo:ghi], in_memory=True)
        check_group((gix, docs, gparts, beg), mygroup, grank, g_size, "hybrid 2x%d (replica %d)" % (world // 2, rank // g_size), rounds=3)
    dist.barrier()
    if rank == 0:
        print("multigpu_check ok: world=%d, sharded step (4 sets x 4 streams, wrapped), host-buffer form, fused decode rows, "
              "NCCL plumbing, hits%s bit-exact vs per-shard oracle" % (world, ", hybrid groups" if world >= 4 else ""))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
