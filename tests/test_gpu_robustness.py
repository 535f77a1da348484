"""GPU suite: the failure modes the round-1 review found.
  * service: concurrent mixed endpoints against ONE index handle with a small coalescing window (every engine call must run
    on the service's single worker thread);
  * shared mask workspace: an asynchronous mask_logits / distinct_mask_batch on a torch stream followed by host-style calls
    (private stream) and by calls on a second torch stream;
  * a prefix longer than the text (the resident-layout tail of k_search must not index outside the text);
  * the remote index facade with CUDA scores."""
import random
import threading

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
V, EOS = 128256, 128009


def corpus(seed=17, n_docs=60, vocab=3000):
    rng = random.Random(seed)
    docs = [[rng.randrange(vocab) for _ in range(rng.randint(10, 120))] + [EOS] for _ in range(n_docs)]
    for k in range(n_docs // 4):
        docs[n_docs - 1 - k] = docs[k][2:30] + docs[n_docs - 1 - k]
    return rng, docs


def test_service_concurrent_mixed_endpoints(cuda):
    from starlette.testclient import TestClient
    from fm_index_constrained_decoding_b200 import FMIndex
    from fm_index_constrained_decoding_b200.service import create_app
    from oracle.oracle import OracleFMIndex
    rng, docs = corpus()
    o = OracleFMIndex()
    o.initialize(docs)
    ix = FMIndex()
    ix.initialize(docs, in_memory=True)
    jobs = []
    for k in range(160):
        d = rng.choice(docs)
        i = rng.randrange(len(d) - 1)
        q = d[i:i + rng.randint(1, 6)]
        jobs.append((k % 5, q, rng.randrange(len(docs))))
    out, err = [None] * len(jobs), []
    with TestClient(create_app(ix, window_ms=0.2)) as client:
        def hit(k):
            kind, q, di = jobs[k]
            try:
                if kind == 0:
                    out[k] = tuple(client.post("/get_range", json={"sequence": q}).json()["range"])
                elif kind == 1:
                    out[k] = client.post("/continuations", json={"sequence": q}).json()["continuations"]
                elif kind == 2:
                    r = client.post("/search", json={"sequence": q, "limit": 6}).json()
                    out[k] = (r["count"], r["doc_indices"])
                elif kind == 3:
                    out[k] = client.get("/docs/%d" % di).json()["tokens"]
                else:
                    lo, hi = o.get_range(q)
                    r = client.post("/get_distinct_count_multi", json={"lows": [lo, 0], "highs": [hi, 50]}).json()["distinct_list"]
                    out[k] = [tuple(map(list, x)) for x in r]
            except Exception as e:  # noqa: BLE001
                err.append(repr(e))
        for a in range(0, len(jobs), 32):  # 32 requests in flight at a time
            th = [threading.Thread(target=hit, args=(k,)) for k in range(a, min(a + 32, len(jobs)))]
            [t.start() for t in th]
            [t.join() for t in th]
    assert not err, err[:3]
    for k, (kind, q, di) in enumerate(jobs):
        if kind == 0:
            assert out[k] == o.get_range(q)
        elif kind == 1:
            assert out[k] == o.get_continuations(q)
        elif kind == 2:
            assert out[k] == (o.get_count(q), list(o.get_doc_indices(q))[:6])
        elif kind == 3:
            assert out[k] == o.get_doc(di)
        else:
            lo, hi = o.get_range(q)
            assert out[k] == [tuple(map(list, x)) for x in o.get_distinct_count_multi([lo, 0], [hi, 50])]


def test_mask_workspace_is_safe_across_streams(cuda):
    """distinct_mask_batch / mask_logits on torch streams interleaved with host-style calls on the same index."""
    from fm_index_constrained_decoding_b200 import FMIndex
    from fm_index_constrained_decoding_b200.index import ragged
    from oracle.oracle import OracleFMIndex
    rng, docs = corpus(23, n_docs=200, vocab=300)  # small alphabet: wide ranges, real tree walks
    o = OracleFMIndex()
    o.initialize(docs)
    ix = FMIndex()
    ix.initialize(docs, in_memory=True)
    seqs = [[rng.randrange(300)] for _ in range(256)] + [rng.choice(docs)[:2] for _ in range(256)]
    toks, offs = ragged(seqs)
    dt, do = torch.from_numpy(toks).cuda(), torch.from_numpy(offs).cuda()
    sp, ep = ix.get_range_batch(dt, do)
    torch.cuda.synchronize()
    osp, oep, om, _ = o.query_batch(toks, offs, want_mask=True, vocab=V, nthreads=4)
    want = torch.from_numpy(om.view(np.int32))
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for it in range(6):
        with torch.cuda.stream(s1):
            m1 = ix.distinct_mask_batch(sp, ep, V)
        host = [ix.get_distinct_count(int(osp[i]), int(oep[i])) for i in (it, 300 + it)]  # host-style, private stream
        with torch.cuda.stream(s2):
            m2 = ix.distinct_mask_batch(sp, ep, V)
        torch.cuda.synchronize()
        assert torch.equal(m1.cpu(), want) and torch.equal(m2.cpu(), want), it
        assert [tuple(map(list, h)) for h in host] == [tuple(map(list, o.get_distinct_count(int(osp[i]), int(oep[i])))) for i in (it, 300 + it)]


@pytest.mark.parametrize("layout", ["resident", "sampled"])
def test_prefix_longer_than_the_text(cuda, layout):
    from fm_index_constrained_decoding_b200 import FMIndex
    from fm_index_constrained_decoding_b200.index import ragged
    from oracle.oracle import OracleFMIndex
    for docs in ([[7, 8, 9] * 30], [[5, 6] * 40 + [7]], [[3] * 70]):  # n >= 64 symbols, very repetitive
        o = OracleFMIndex()
        o.initialize(docs)
        ix = FMIndex(layout=layout)
        ix.initialize(docs, in_memory=True)
        n = len(docs[0])
        seqs = [docs[0][:k] for k in (1, 5, 33, 34, 65, n - 1, n)] + [docs[0] + docs[0][:5], docs[0] * 3, [docs[0][0]] * (4 * n)]
        toks, offs = ragged(seqs)
        sp, ep = ix.get_range_batch(torch.from_numpy(toks).cuda(), torch.from_numpy(offs).cuda())
        got = list(zip(sp.tolist(), ep.tolist()))
        assert got == [o.get_range(s) for s in seqs], (layout, docs[0][:4])


def test_remote_index_with_cuda_scores(cuda):
    from starlette.testclient import TestClient
    from fm_index_constrained_decoding_b200.client import RemoteFMIndex
    from fm_index_constrained_decoding_b200.processors import SoftBoostLogitsProcessor
    from fm_index_constrained_decoding_b200.service import create_app
    from oracle.oracle import OracleFMIndex, soft_boost_reference
    rng, docs = corpus(5, n_docs=30, vocab=500)
    docs = [[t for t in d if t != EOS] for d in docs]  # no EOS 128009: the vocabulary of this test is 600
    o = OracleFMIndex()
    o.initialize(docs)
    with TestClient(create_app(o)) as client:
        remote = RemoteFMIndex(session=client)
        Vs, eos, pad = 600, 590, 591
        marker = [78191, 128007, 271]
        rows = [[1, 2] + marker + rng.choice(docs)[3:7] for _ in range(5)]
        scores = torch.randn(len(rows), Vs, generator=torch.Generator().manual_seed(4))
        proc = SoftBoostLogitsProcessor(remote, 1, variant="causal", model_name="llama", pad_token_id=pad, eos_token_id=eos,
                                        always_allow_eos=True)
        exp = soft_boost_reference(o, rows, scores, 1, "causal", eos, pad, marker, 0, True, None, None, 2.0, 5)
        got = proc(torch.tensor(rows).cuda(), scores.cuda())
        assert got.is_cuda and torch.equal(got.cpu(), exp)


@pytest.mark.parametrize("setting", ["table", "no_table", "chunk_scan", "inline_only", "sampled"])
def test_wide_range_paths_agree_with_the_oracle(cuda, setting):
    """The same wide ranges through every enumeration path: unigram continuation table, chunked BWT scan, inline BWT scan,
    tree walk (resident without table / sampled layout) -- bitmaps, fused logits rows (fp32 + bf16) and sparse lists."""
    from fm_index_constrained_decoding_b200 import FMIndex
    from fm_index_constrained_decoding_b200.index import ragged
    from oracle.oracle import OracleFMIndex
    rng = random.Random(31)
    vocab = 180
    weights = [1.0 / (r + 1) for r in range(vocab)]
    docs = [rng.choices(range(vocab), weights=weights, k=rng.randint(200, 900)) + [EOS] for _ in range(300)]
    o = OracleFMIndex()
    o.initialize(docs)
    ix = FMIndex(layout="sampled" if setting == "sampled" else "resident")
    if setting in ("table", "no_table"):
        ix.tune(scan_max=1024)          # table entries for every token with more than 1,024 occurrences
    elif setting == "chunk_scan":
        ix.tune(scan_max=1 << 20)       # everything wider than 1,024 rows is scanned in chunks
    elif setting == "inline_only":
        ix.tune(scan_max=0)             # > 1,024 rows: tree walk, no table entry below the widest
    ix.initialize(docs, in_memory=True)
    if setting == "no_table":
        ix.tune(unigram_table=0)
    seqs = [[t] for t in range(vocab)] + [[EOS]] + [rng.choice(docs)[i:i + 2] for i in range(60)] + [rng.choice(docs)[i:i + 3] for i in range(60)]
    toks, offs = ragged(seqs)
    dt, do = torch.from_numpy(toks).cuda(), torch.from_numpy(offs).cuda()
    osp, oep, om, _ = o.query_batch(toks, offs, want_mask=True, vocab=V, nthreads=4)
    assert int((oep - osp).max()) > 5000  # the workload really has wide ranges
    sp, ep, mask = ix.query_batch(dt, do, V)
    assert (sp.cpu().numpy().astype(np.uint64) == osp).all() and (ep.cpu().numpy().astype(np.uint64) == oep).all()
    assert (mask.cpu().numpy().view(np.uint32) == om).all(), setting
    # a shifted / clipped token window (shift 3, vocab 100): the table rows are cut out with a different bit offset
    m2 = ix.distinct_mask_batch(sp, ep, 100, shift=3)
    exp2 = np.zeros((len(seqs), 4), np.uint32)
    for i in range(len(seqs)):
        for sym in o.distinct(int(osp[i]), int(oep[i])):
            t = sym - 3
            if 0 <= t < 100:
                exp2[i, t >> 5] |= np.uint32(1 << (t & 31))
    assert (m2.cpu().numpy().view(np.uint32) == exp2).all(), setting
    bits = torch.from_numpy(np.unpackbits(om.view(np.uint8), axis=1, bitorder="little")[:, :V].astype(bool))
    bits[:, EOS] = True
    for dtp in (torch.float32, torch.bfloat16):
        lg = torch.randn(len(seqs), V, generator=torch.Generator().manual_seed(1)).to(dtp)
        got = ix.mask_logits(sp, ep, lg.cuda(), always_allow=EOS).cpu()
        assert torch.equal(got, torch.where(bits, lg + 0.0, torch.full_like(lg, float("-inf")))), (setting, dtp)
    for i in (0, 5, 100, vocab + 3):
        assert tuple(map(list, ix.get_distinct_count(int(osp[i]), int(oep[i])))) == tuple(map(list, o.get_distinct_count(int(osp[i]), int(oep[i]))))


def test_graph_captured_host_pipeline(cuda):
    """fmidx_query_graph_*: the captured step gives the oracle's ranges and distinct-token counts, over many replays (the two
    captures alternate between the work-list counter pairs), with different batches per slot, and wide ranges in the mix."""
    from fm_index_constrained_decoding_b200 import FMIndex
    from fm_index_constrained_decoding_b200.index import GraphQuerySlots, ragged
    from oracle.oracle import OracleFMIndex
    rng = random.Random(41)
    vocab = 400
    weights = [1.0 / (r + 1) for r in range(vocab)]
    docs = [rng.choices(range(vocab), weights=weights, k=rng.randint(100, 600)) + [EOS] for _ in range(200)]
    o = OracleFMIndex()
    o.initialize(docs)
    ix = FMIndex()
    ix.tune(scan_max=1024)
    ix.initialize(docs, in_memory=True)
    B, MAXTOK, NS = 256, 256 * 8, 3
    gs = GraphQuerySlots(ix, NS, B, MAXTOK, V)
    batches = []
    for k in range(5):
        seqs = [[rng.randrange(vocab)] if j % 7 == 0 else (lambda d, i: d[i:i + rng.randint(1, 8)])(rng.choice(docs), rng.randrange(50)) for j in range(B)]
        toks, offs = ragged(seqs)
        osp, oep, _, ond = o.query_batch(toks, offs, want_mask=True, vocab=V, nthreads=4)
        batches.append((toks, offs, osp, oep, ond.astype(np.uint32)))
    pending = [None] * NS
    for step in range(4 * NS + 2):
        i = step % NS
        gs.wait(i)
        if pending[i] is not None:
            _, _, osp, oep, ond = batches[pending[i]]
            assert (gs.sp[i] == osp).all() and (gs.ep[i] == oep).all() and (gs.n_distinct[i] == ond).all(), step
        toks, offs = batches[step % len(batches)][:2]
        gs.offs[i][:] = offs
        gs.toks[i][:toks.size] = toks
        gs.launch(i)
        pending[i] = step % len(batches)
    for i in range(NS):
        gs.wait(i)
        _, _, osp, oep, ond = batches[pending[i]]
        assert (gs.sp[i] == osp).all() and (gs.ep[i] == oep).all() and (gs.n_distinct[i] == ond).all()
    with pytest.raises(Exception):  # a captured slot refuses the eager submit
        toks, offs = batches[0][:2]
        sp = np.empty(B, np.uint64)
        from fm_index_constrained_decoding_b200 import _native
        _native.check(_native.lib().fmidx_query_submit(ix._h, 0, toks.ctypes.data, offs.ctypes.data, B, 10, V, sp.ctypes.data, sp.ctypes.data, None))


def test_graph_captured_device_step(cuda):
    """fmidx_query_graph_prepare_device / _launch_on: the device-resident step captured for fixed buffers gives the oracle's
    ranges and bitmaps over many replays, with the inputs rewritten in place between replays, on the slot's own stream and on
    a caller's stream."""
    from fm_index_constrained_decoding_b200 import FMIndex
    from fm_index_constrained_decoding_b200.index import DeviceGraphSteps, ragged
    from oracle.oracle import OracleFMIndex
    rng = random.Random(43)
    vocab = 400
    weights = [1.0 / (r + 1) for r in range(vocab)]
    docs = [rng.choices(range(vocab), weights=weights, k=rng.randint(100, 600)) + [EOS] for _ in range(200)]
    o = OracleFMIndex()
    o.initialize(docs)
    ix = FMIndex()
    ix.tune(scan_max=1024)
    ix.initialize(docs, in_memory=True)
    B, MAXTOK, NS = 256, 256 * 8, 2
    gs = DeviceGraphSteps(ix, NS, B, MAXTOK, V, first_slot=20)
    batches = []
    for k in range(4):
        seqs = [[rng.randrange(vocab)] if j % 5 == 0 else (lambda d, i: d[i:i + rng.randint(1, 8)])(rng.choice(docs), rng.randrange(50)) for j in range(B)]
        toks, offs = ragged(seqs)
        osp, oep, om, _ = o.query_batch(toks, offs, want_mask=True, vocab=V, nthreads=4)
        batches.append((torch.from_numpy(toks), torch.from_numpy(offs), osp.astype(np.int64), oep.astype(np.int64), om.view(np.int32)))
    side = torch.cuda.Stream()
    for step in range(9):
        i = step % NS
        toks, offs, osp, oep, om = batches[step % len(batches)]
        st = side if step % 3 == 2 else None
        if st is None:
            gs.wait(i)
        else:
            torch.cuda.synchronize()
        with torch.cuda.stream(st if st is not None else torch.cuda.current_stream()):
            gs.toks[i][:toks.numel()].copy_(toks)
            gs.offs[i].copy_(offs)
        torch.cuda.synchronize()
        gs.launch(i, st)
        torch.cuda.synchronize()
        assert (gs.sp[i].cpu().numpy() == osp).all() and (gs.ep[i].cpu().numpy() == oep).all(), step
        assert (gs.mask[i].cpu().numpy() == om.reshape(B, -1)).all(), step


@pytest.mark.parametrize("layout", ["resident", "sampled"])
def test_scalar_fast_paths_match_batched_paths_and_oracle(cuda, layout):
    """The scalar (batch 1) calls -- arguments as kernel parameters, results polled from mapped pinned memory -- against
    the batched code paths (scalar_fast=0) and the oracle: get_range, backward_search_multi, backward_search_step, locate,
    distinct, distinct_count; ranges around the 4096-row limit of the scalar distinct kernel and prefixes around its
    96-token limit take the fallback."""
    from fm_index_constrained_decoding_b200 import FMIndex
    from oracle.oracle import OracleFMIndex
    rng = random.Random(47)
    vocab = 600
    weights = [1.0 / (r + 1) for r in range(vocab)]
    docs = [rng.choices(range(vocab), weights=weights, k=rng.randint(50, 400)) + [EOS] for _ in range(120)]
    o = OracleFMIndex()
    o.initialize(docs)
    ix = FMIndex(layout=layout)
    ix.initialize(docs, in_memory=True)
    n = ix.size()
    cases = []
    for _ in range(60):
        d = rng.choice(docs)
        i = rng.randrange(len(d) - 1)
        cases.append(d[i:i + rng.randint(1, 12)])
    cases += [[0], [vocab + 5], [], docs[3][:95], docs[3][:96], docs[3][:97], docs[3][:140], [docs[0][0], 2 ** 31 - 1]]
    # get_range followed by get_distinct_count of the returned range (the decode loop's pair): the enumeration is launched
    # speculatively behind the search; other ranges in between, repeated requests and a rebuilt index must not see stale answers
    for spec in (1, 0):
        ix.tune(scalar_speculate=spec)
        prev = None
        for q in cases:
            lo, hi = ix.get_range(q)
            if prev is not None and rng.random() < 0.3:
                assert ix.get_distinct_count(*prev) == o.get_distinct_count(*prev)   # not the speculated range
            assert ix.get_distinct_count(lo, hi) == o.get_distinct_count(lo, hi), (spec, q[:5])
            assert ix.get_distinct_count(lo, hi) == o.get_distinct_count(lo, hi)     # asked twice
            assert ix.distinct(lo, hi) == o.distinct(lo, hi)
            prev = (lo, hi)
    ix.tune(scalar_speculate=1)
    for fast in (1, 0):
        ix.tune(scalar_fast=fast)
        for q in cases:
            lo, hi = ix.get_range(q)
            assert (lo, hi) == o.get_range(q), (fast, q[:5])
            sym = [t + 10 for t in q][::-1]
            assert ix.backward_search_multi(sym) == o.backward_search_multi(sym), (fast, q[:5])
            assert ix.distinct(lo, hi) == o.distinct(lo, hi)
            assert ix.distinct_count(lo, hi) == o.distinct_count(lo, hi)
        for lo, hi in [(0, n), (0, n + 7), (5, 5), (9, 3), (0, 4096), (0, 4097), (1, 4097), (n - 3, n), (n - 1, n + 100), (n, n + 1),
                       (17, 18), (100, 4000)]:
            assert ix.distinct(lo, hi) == o.distinct(lo, hi), (fast, lo, hi)
            assert ix.distinct_count(lo, hi) == o.distinct_count(lo, hi), (fast, lo, hi)
        for _ in range(40):
            s, lo = rng.randrange(vocab + 20), rng.randrange(n)
            hi = rng.choice([lo, lo + rng.randrange(50), n - 1, n + 5, max(0, lo - 1)])
            assert ix.backward_search_step(s, lo, hi) == o.backward_search_step(s, lo, hi), (fast, s, lo, hi)
        for r in [0, 1, n - 1, n, n + 9, 2 ** 40] + [rng.randrange(n) for _ in range(30)]:
            assert ix.locate(r) == o.locate(r), (fast, r)


@pytest.mark.parametrize("layout", ["resident", "sampled"])
def test_lazy_clearing_of_mask_rows(cuda, layout):
    """fmidx_tune("lazy_clear", 2): only the slices a buffer's previous use dirtied are zeroed.  The same two buffers see a
    rotation of very different batches (wide one-token prefixes, narrow histories, empty and absent prefixes) through
    fmidx_query_batch and through the captured device step; every mask must equal the oracle's, whatever was there before;
    switching the feature off and on again, another buffer on the same slot and a different batch size must not leave
    stale bits either."""
    from fm_index_constrained_decoding_b200 import FMIndex
    from fm_index_constrained_decoding_b200.index import DeviceGraphSteps, ragged
    from oracle.oracle import OracleFMIndex
    rng = random.Random(53)
    vocab = 50000
    weights = [1.0 / (r + 1) ** 1.05 for r in range(vocab)]
    docs = [rng.choices(range(vocab), weights=weights, k=rng.randint(100, 500)) + [EOS] for _ in range(150)]
    o = OracleFMIndex()
    o.initialize(docs)
    ix = FMIndex(layout=layout)
    ix.tune(scan_max=2048, lazy_clear=2)
    ix.initialize(docs, in_memory=True)
    B = 192

    def batch(kind):
        seqs = []
        for j in range(B):
            d = rng.choice(docs)
            i = rng.randrange(len(d) - 1)
            if kind == "wide":
                seqs.append([d[i]] if j % 3 else [rng.randrange(5)])
            elif kind == "narrow":
                seqs.append(d[i:i + rng.randint(3, 9)])
            elif kind == "empty":
                seqs.append([] if j % 2 else [vocab + 7, 3])
            else:
                seqs.append(d[i:i + rng.randint(1, 6)])
        toks, offs = ragged(seqs)
        osp, oep, om, _ = o.query_batch(toks, offs, want_mask=True, vocab=V, nthreads=4)
        return torch.from_numpy(toks).cuda(), torch.from_numpy(offs).cuda(), om.view(np.int32)
    kinds = ["wide", "narrow", "mixed", "empty", "narrow", "wide", "mixed", "narrow", "narrow", "empty", "mixed"]
    batches = [batch(k) for k in kinds]
    bufs = [torch.full((B, (V + 31) // 32), -1, dtype=torch.int32, device="cuda") for _ in range(2)]  # start dirty
    spb, epb = torch.empty(B, dtype=torch.int64, device="cuda"), torch.empty(B, dtype=torch.int64, device="cuda")
    for step, (toks, offs, om) in enumerate(batches):
        i = step % 2
        if step == 6:
            ix.tune(lazy_clear=1)   # one call without book-keeping ...
        if step == 7:
            ix.tune(lazy_clear=2)   # ... must not leave a stale state behind
        sp, ep, m = ix.query_batch(toks, offs, V, ws_slot=i, out=(spb, epb, bufs[i]))
        torch.cuda.synchronize()
        assert (m.cpu().numpy() == om).all(), (layout, step, kinds[step])
    # another buffer and another batch size on the same slot
    other = torch.full((B, (V + 31) // 32), -1, dtype=torch.int32, device="cuda")
    toks, offs, om = batches[1]
    _, _, m = ix.query_batch(toks, offs, V, ws_slot=0, out=(spb, epb, other))
    assert (m.cpu().numpy() == om).all()
    half = ragged([[int(t) for t in toks[offs[j]:offs[j + 1]].tolist()] for j in range(B // 2)])
    _, _, m = ix.query_batch(torch.from_numpy(half[0]).cuda(), torch.from_numpy(half[1]).cuda(), V, ws_slot=0, out=(spb[:B // 2], epb[:B // 2], bufs[0][:B // 2]))
    assert (m.cpu().numpy() == om[:B // 2]).all()
    # captured device step with in-place input rewrites
    maxtok = max(t.numel() for t, _, _ in batches)
    gs = DeviceGraphSteps(ix, 1, B, maxtok, V, first_slot=9)
    for step, (toks, offs, om) in enumerate(batches):
        gs.toks[0][:toks.numel()].copy_(toks)
        gs.offs[0].copy_(offs)
        torch.cuda.synchronize()
        gs.launch(0)
        gs.wait(0)
        assert (gs.mask[0].cpu().numpy() == om).all(), (layout, "graph", step, kinds[step])
