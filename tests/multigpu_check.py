"""Multi-GPU parity check of the document-sharded path (NOT a pytest file: needs N GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tests/multigpu_check.py

Every rank builds the index of its document shard on its GPU; the expectation is one CPU oracle per
shard with results OR/sum-reduced on the host."""
import os
import random
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ["FMIDX_DEVICE"] = str(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from fm_index_constrained_decoding_b200 import FMIndex, sharded
    from fm_index_constrained_decoding_b200.index import ragged
    from oracle.oracle import OracleFMIndex

    V, eos = 128256, 128009
    rng = random.Random(3)
    docs = [[rng.randrange(3000) for _ in range(rng.randint(10, 200))] + [eos] for _ in range(400)]
    for k in range(100):
        docs[300 + k] = docs[k][3:40] + docs[300 + k]
    beg = [0]
    for d in docs:
        beg.append(beg[-1] + len(d))
    parts = sharded.shard_documents(beg, world)
    lo, hi = parts[rank]
    ix = FMIndex(device=local)
    ix.initialize(docs[lo:hi], in_memory=True)
    sh = sharded.ShardedFMIndex(ix, doc_base=lo, pos_base=beg[lo])
    B = 64 * world
    seqs = []
    for _ in range(B):
        d = rng.choice(docs)
        i = rng.randrange(len(d) - 1)
        seqs.append(d[i:i + rng.randint(1, 8)])
    toks, offs = ragged(seqs)
    dt, do = torch.from_numpy(toks).cuda(), torch.from_numpy(offs).cuda()
    sp, ep = sh.get_range_batch(dt, do)
    # host expectation
    words = (V + 31) // 32
    exp_mask = np.zeros((B, words), np.uint32)
    exp_cnt = np.zeros(B, np.int64)
    shard_res, shard_masks = [], []
    for a, b in parts:
        o = OracleFMIndex()
        o.initialize(docs[a:b])
        osp, oep, om, _ = o.query_batch(toks, offs, want_mask=True, vocab=V, nthreads=4)
        exp_mask |= om
        shard_masks.append(om)
        exp_cnt += (oep - osp).astype(np.int64)
        shard_res.append((osp, oep, o, beg[a]))
    assert (sp.cpu().numpy().astype(np.uint64) == shard_res[rank][0]).all()
    assert (ep.cpu().numpy().astype(np.uint64) == shard_res[rank][1]).all()
    own = B // world
    m_own = sh.distinct_mask_batch(sp, ep, V)
    assert (m_own.cpu().numpy().view(np.uint32) == exp_mask[rank * own:(rank + 1) * own]).all()
    m_all = sh.distinct_mask_batch(sp, ep, V, everywhere=True)
    assert (m_all.cpu().numpy().view(np.uint32) == exp_mask).all()
    assert (sh.count_batch(sp, ep).cpu().numpy() == exp_cnt).all()
    # pull-style exchange over peer memory (CUDA IPC): same result as the all-to-all path, three buffer sets
    px = sharded.PeerMaskExchange(ix, B, V, n_buffers=3)
    for it in range(4):
        i = it % 3
        px.compute_local(i, sp, ep)
        got = px.exchange(i)
        torch.cuda.synchronize()
        assert (got.cpu().numpy().view(np.uint32) == exp_mask[rank * own:(rank + 1) * own]).all(), "pull exchange mismatch"
        # rows owned by peers hold this shard's bits in their OCCUPIED slices (the rest is never read)
        loc = px.local_views[i].cpu().numpy().view(np.uint32)
        occ = px.xlocal[i][:B].cpu().numpy().view(np.uint32)
        for r in range(B):
            if rank * own <= r < (rank + 1) * own:
                continue
            for sl in range(32):
                w0, w1 = sl * 128, min((sl + 1) * 128, words)
                if (occ[r] >> sl) & 1:
                    assert (loc[r, w0:w1] == shard_masks[rank][r, w0:w1]).all()
                else:
                    assert not shard_masks[rank][r, w0:w1].any()
        assert (px.total_counts(i).cpu().numpy() == exp_cnt).all()
    px.close()
    # fused rows
    g = torch.Generator().manual_seed(rank)
    logits = torch.randn(own, V, generator=g)
    out = sh.mask_logits_rows(sp, ep, logits.cuda(), always_allow=eos).cpu()
    bits = np.unpackbits(exp_mask[rank * own:(rank + 1) * own].view(np.uint8), axis=1, bitorder="little")[:, :V].astype(bool)
    bits[:, eos] = True
    exp = torch.where(torch.from_numpy(bits), logits + 0.0, torch.full_like(logits, float("-inf")))
    assert torch.equal(out, exp)
    # hits
    K = 4
    hits, hv = sh.locate_topk(sp, ep, K)
    hits, hv = hits.cpu().numpy(), hv.cpu().numpy()
    for s, (osp, oep, o, base) in enumerate(shard_res):
        for qi in range(B):
            for k in range(K):
                r = int(osp[qi]) + k
                if r < int(oep[qi]):
                    assert hv[s, qi, k] and hits[s, qi, k] == o.locate(r) + base
                else:
                    assert not hv[s, qi, k]
    dist.barrier()
    if rank == 0:
        print("multigpu_check ok: world=%d, %d prefixes, masks/counts/fused rows/hits bit-exact vs per-shard oracle" % (world, B))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
