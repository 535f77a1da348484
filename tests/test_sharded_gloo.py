"""CPU suite: the N>1 exchange logic on a world_size-2 gloo group.  The per-shard results are
produced by the CPU oracle (one oracle index per shard, as DESIGN.md section 6 defines sharded
parity); the collectives are the product's (fm_index_constrained_decoding_b200/sharded.py)."""
import os
import random

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fm_index_constrained_decoding_b200 import sharded
from fm_index_constrained_decoding_b200.index import ragged

VOCAB = 700


def make_corpus():
    rng = random.Random(12)
    docs = [[rng.randrange(600) for _ in range(rng.randint(5, 60))] + [699] for _ in range(40)]
    for k in range(10):  # repeats across shards
        docs[30 + k] = docs[k][2:20] + docs[30 + k]
    seqs = []
    for _ in range(64):
        d = rng.choice(docs)
        i = rng.randrange(len(d) - 1)
        seqs.append(d[i:i + rng.randint(1, 6)])
    return docs, seqs


def worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.oracle import OracleFMIndex
        docs, seqs = make_corpus()
        beg = [0]
        for d in docs:
            beg.append(beg[-1] + len(d))
        parts = sharded.shard_documents(beg, world)
        lo, hi = parts[rank]
        o = OracleFMIndex()
        o.initialize(docs[lo:hi])
        toks, offs = ragged(seqs)
        sp, ep, m, _ = o.query_batch(toks, offs, want_mask=True, vocab=VOCAB, nthreads=1)
        sp_t = torch.from_numpy(sp.astype(np.int64))
        ep_t = torch.from_numpy(ep.astype(np.int64))
        mask_t = torch.from_numpy(m.view(np.int32).copy())
        pieces, own = sharded.or_reduce_scatter(mask_t)
        allm = sharded.or_all(mask_t)
        cnt = sharded.sum_counts(sp_t, ep_t)
        K = 4
        rows = sp_t.unsqueeze(1) + torch.arange(K).unsqueeze(0)
        valid = rows < ep_t.unsqueeze(1)
        pos = torch.tensor([[o.locate(int(r)) if v else 0 for r, v in zip(rr, vv)] for rr, vv in zip(rows.tolist(), valid.tolist())])
        hits, hv = sharded.gather_hits(pos, valid, beg[lo])
        q.put((rank, own.numpy(), allm.numpy(), cnt.numpy(), hits.numpy(), hv.numpy(), parts))
    finally:
        dist.destroy_process_group()


def test_shard_documents_balanced():
    beg = [0]
    rng = random.Random(1)
    for _ in range(1000):
        beg.append(beg[-1] + rng.randint(10, 500))
    for w in (1, 2, 3, 8):
        parts = sharded.shard_documents(beg, w)
        assert parts[0][0] == 0 and parts[-1][1] == 1000
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        sizes = [beg[b] - beg[a] for a, b in parts]
        assert max(sizes) - min(sizes) <= 1000


@pytest.mark.timeout(300)
def test_world2_exchange_matches_host_reduction():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # host-side expectation: one oracle per shard, OR / sum reduced here
    from oracle.oracle import OracleFMIndex
    docs, seqs = make_corpus()
    parts = res[0][6]
    exp_mask = np.zeros((len(seqs), (VOCAB + 31) // 32), dtype=np.uint32)
    exp_cnt = np.zeros(len(seqs), dtype=np.int64)
    toks, offs = ragged(seqs)
    shard_or = []
    for lo, hi in parts:
        o = OracleFMIndex()
        o.initialize(docs[lo:hi])
        sp, ep, m, _ = o.query_batch(toks, offs, want_mask=True, vocab=VOCAB, nthreads=1)
        exp_mask |= m
        exp_cnt += (ep - sp).astype(np.int64)
        shard_or.append((sp, ep, o))
    own_rows = len(seqs) // world
    for rank, own, allm, cnt, hits, hv, _ in res:
        assert (own.view(np.uint32) == exp_mask[rank * own_rows:(rank + 1) * own_rows]).all()
        assert (allm.view(np.uint32) == exp_mask).all()
        assert (cnt == exp_cnt).all()
    # gathered hits are global positions whose text really is the pattern (reversed doc coordinates)
    hits, hv = res[0][4], res[0][5]
    assert (hits == res[1][4]).all()
    beg = [0]
    for d in docs:
        beg.append(beg[-1] + len(d))
    flatT = []
    for d in docs:
        flatT.extend(d[::-1])
    for s in range(world):
        for qi, pat in enumerate(seqs):
            for k in range(4):
                if hv[s, qi, k]:
                    p = int(hits[s, qi, k])
                    assert flatT[p:p + len(pat)] == pat[::-1]
    # a monolithic index agrees whenever the pattern cannot straddle a shard boundary (no EOS inside)
    mono = OracleFMIndex()
    mono.initialize(docs)
    for qi, pat in enumerate(seqs):
        if 699 not in pat[:-1]:
            assert mono.get_count(pat) == exp_cnt[qi]
