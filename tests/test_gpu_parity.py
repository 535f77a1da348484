"""GPU suite (-m gpu): the CUDA path, called through the C ABI (ctypes -> libfmidx.so), must be
bit-exact against the CPU oracle and the golden fixtures."""
import ctypes as C
import os
import random

import numpy as np
import pytest
import torch

from conftest import golden_cases

pytestmark = pytest.mark.gpu
U64_MAX = (1 << 64) - 1


def make(docs, **kw):
    from fm_index_constrained_decoding_b200 import FMIndex
    ix = FMIndex(**kw)
    ix.initialize(docs, in_memory=True)
    return ix


def oracle(docs):
    from oracle.oracle import OracleFMIndex
    o = OracleFMIndex()
    o.initialize(docs)
    return o


def random_docs(rng, V, ndocs, lo, hi, repeats=True):
    docs = [[rng.randrange(V) for _ in range(rng.randint(lo, hi))] for _ in range(ndocs)]
    if repeats and ndocs > 2:  # planted repeats so long prefixes have several matches
        for _ in range(ndocs // 2):
            a, b = rng.randrange(ndocs), rng.randrange(ndocs)
            if len(docs[a]) > 8:
                i = rng.randrange(len(docs[a]) - 4)
                docs[b] = docs[b] + docs[a][i:i + rng.randint(4, 40)]
    return docs


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: "n%d" % c["size"])
@pytest.mark.parametrize("dens", [(32, 64, False), (1, 1, True), (4, 2, False)])
def test_golden_scalar_surface(cuda, case, dens):
    ix = make(case["docs"], sa_dens=dens[0], isa_dens=dens[1], keep_text=dens[2])
    N = case["size"]
    assert ix.size() == N and ix.beginnings == case["beginnings"] and len(ix) == N - 1
    assert [ix.locate(r) for r in range(N)] == case["SA"]
    assert ix.locate(N) == case["locate_oob"] == U64_MAX
    for q in case["queries"]:
        r = ix.get_range(q["q"])
        assert list(r) == q["range"], q
        assert ix.get_distinct_count(*r) == (q["distinct"], q["counts"])
        assert ix.get_count(q["q"]) == q["range"][1] - q["range"][0]
        assert ix.get_continuations(q["q"]) == q["distinct"]
    assert [list(x) for x in (ix.occurring_distinct, ix.occurring_counts)] == case["occurring"]
    assert [ix.get_doc(i) for i in range(len(case["docs"]))] == case["get_doc"]
    for s in case["steps"]:
        assert list(ix.backward_search_step(s["sym"], s["low"], s["high"])) == s["out"], s
    for d in case["distinct_count"]:
        assert list(ix.distinct_count(d["low"], d["high"])) == d["out"], d
        assert list(ix.distinct(d["low"], d["high"])) == d["out"][0::2]
    for e in case["extract"]:
        assert list(ix.extract_text(e["begin"], e["end"])) == e["out"], e


def test_builder_suffix_array_vs_oracle(cuda):
    """device prefix-doubling SA == oracle SA-IS, incl. long repeats and runs."""
    from fm_index_constrained_decoding_b200 import fm_index, _native
    from oracle.oracle import lib
    rng = np.random.default_rng(0)
    base = rng.integers(1, 2000, size=30_000)
    data = np.concatenate([base, rng.integers(1, 4, size=20_000), base[:12_000], np.full(3000, 7),
                           rng.integers(1, 130_000, size=200_000), base[5000:9000]]).astype(np.uint64)
    ix = fm_index.FMIndex(sa_dens=1, isa_dens=1)
    ix.initialize(data)
    got = np.empty(data.size + 1, dtype=np.uint32)
    _native.check(_native.lib().fmidx_debug_copy_sa(ix._h, got.ctypes.data_as(C.c_void_p), got.size))
    ref = np.empty(data.size + 1, dtype=np.uint32)
    assert lib().fmo_suffix_array(data.ctypes.data, data.size, ref.ctypes.data) == 0
    assert (got == ref).all()


@pytest.mark.parametrize("layout", ["resident", "sampled"])
@pytest.mark.parametrize("V,ndocs,lo,hi", [(3, 40, 1, 200), (300, 200, 5, 300), (128256, 300, 20, 400)])
def test_batched_search_and_masks_vs_oracle(cuda, V, ndocs, lo, hi, layout):
    from fm_index_constrained_decoding_b200.index import ragged
    rng = random.Random(V)
    docs = random_docs(rng, V, ndocs, lo, hi)
    ix, o = make(docs, layout=layout), oracle(docs)
    seqs = [[]]
    for _ in range(1500):
        d = rng.choice(docs)
        i = rng.randrange(len(d))
        q = d[i:i + rng.randint(1, 32)]
        if rng.random() < 0.25:
            q = list(q)
            q[rng.randrange(len(q))] = rng.randrange(V)
        if rng.random() < 0.02:
            q = q + [V + 5]  # symbol beyond the alphabet
        seqs.append(q)
    toks, offs = ragged(seqs)
    vocab = max(V, 64)
    sp_o, ep_o, m_o, _ = o.query_batch(toks, offs, want_mask=True, vocab=vocab, nthreads=4)
    dt, do = torch.from_numpy(toks).cuda(), torch.from_numpy(offs).cuda()
    sp, ep = ix.get_range_batch(dt, do)
    assert (sp.cpu().numpy().astype(np.uint64) == sp_o).all()
    assert (ep.cpu().numpy().astype(np.uint64) == ep_o).all()
    mask, counts = ix.distinct_mask_batch(sp, ep, vocab, want_counts=True)
    assert (mask.cpu().numpy().view(np.uint32) == m_o).all()
    # counts at set bits and the sparse list API
    cn = counts.cpu().numpy()
    lows, highs = sp_o[:200].tolist(), ep_o[:200].tolist()
    res = ix.get_distinct_count_multi(lows, highs)
    exp = o.get_distinct_count_multi(lows, highs)
    assert res == exp
    for k in range(200):
        for t, c in zip(*exp[k]):
            assert cn[k, t] == c
    # host-buffer entry point (what a non-CUDA caller binds)
    from fm_index_constrained_decoding_b200 import _native
    B = len(seqs)
    words = (vocab + 31) // 32
    hsp, hep = np.empty(B, np.uint64), np.empty(B, np.uint64)
    hm = np.empty((B, words), np.uint32)
    _native.check(_native.lib().fmidx_query_host(ix._h, toks.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p), B, 10,
                                                 vocab, hsp.ctypes.data_as(C.c_void_p), hep.ctypes.data_as(C.c_void_p),
                                                 hm.ctypes.data_as(C.c_void_p)))
    assert (hsp == sp_o).all() and (hep == ep_o).all() and (hm == m_o).all()


def test_step_batch_and_raw_multi(cuda):
    rng = random.Random(5)
    docs = random_docs(rng, 50, 60, 5, 120)
    ix, o = make(docs), oracle(docs)
    N = o.size()
    syms = [rng.randrange(0, 70) for _ in range(3000)]
    lows = [rng.randint(0, N + 1) for _ in range(3000)]
    highs = [rng.randint(0, N + 2) for _ in range(3000)]
    t = lambda x: torch.tensor(x, dtype=torch.int64, device="cuda")
    sp, ep = ix.step_batch(t(syms), t(lows), t(highs))
    exp = [o.backward_search_step(s, l, h) for s, l, h in zip(syms, lows, highs)]
    assert list(zip(sp.tolist(), ep.tolist())) == exp
    for _ in range(100):
        q = [rng.randrange(8, 64) for _ in range(rng.randint(0, 6))]
        assert ix.backward_search_multi(q) == o.backward_search_multi(q), q


@pytest.mark.parametrize("path", ["regs", "tma", "pipe", "row"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_fused_mask_logits_vs_reference_arithmetic(cuda, dtype, path, monkeypatch):
    """document_constrained_generation_vllm.py:46-137 on a [B, V] batch (config 4 semantics);
    both logits staging variants (bulk-async-copy through shared memory / registers)."""
    from oracle.oracle import hard_mask_reference
    monkeypatch.setenv("FMIDX_LOGITS_PATH", path)
    rng = random.Random(9)
    V, eos, pad = 128256, 128009, 128004
    docs = random_docs(rng, 5000, 150, 30, 300)
    for d in docs:
        d.append(eos)
    ix, o = make(docs), oracle(docs)
    rows = []
    for k in range(48):
        d = rng.choice(docs)
        i = rng.randrange(len(d) - 1)
        s = d[i:i + rng.randint(1, 12)]
        if k % 7 == 3:
            s = s + [eos]
        if k % 11 == 5:
            s = s + [pad]
        if k % 5 == 1:
            s = s + [rng.randrange(5000)]  # likely dead end
        rows.append(s)
    g = torch.Generator().manual_seed(3)
    scores = torch.randn(len(rows), V, generator=g).to(dtype)
    scores[0, 5] = float("inf")
    scores[1, 7] = float("-inf")
    scores[2, 9] = -0.0
    scores[3, rows[3][0]] = float("nan")
    for stop_at_count, allow_eos in [(0, True), (0, False), (3, True)]:
        # every row has its own length here, so run the reference row by row (its batch form needs equal lengths)
        exp = torch.cat([hard_mask_reference(o, [r], scores[k:k + 1], 1, eos, pad, stop_at_count, allow_eos)
                         for k, r in enumerate(rows)])
        # GPU: ranges, then override rows exactly as vllm.py:81-84,105-118 decide them
        from fm_index_constrained_decoding_b200.index import ragged
        toks, offs = ragged(rows)
        sp, ep = ix.get_range_batch(torch.from_numpy(toks).cuda(), torch.from_numpy(offs).cuda())
        ptoks, poffs = ragged([r[:-1] for r in rows])
        psp, pep = ix.get_range_batch(torch.from_numpy(ptoks).cuda(), torch.from_numpy(poffs).cuda())
        last = torch.tensor([r[-1] for r in rows], device="cuda")
        finished = (last == eos) | (last == pad)
        count = torch.where(finished, torch.zeros_like(psp), pep - psp)
        override = torch.full((len(rows),), -1, dtype=torch.int32, device="cuda")
        override[finished] = pad
        if stop_at_count > 0:
            override[count <= stop_at_count] = eos
        sp = torch.where(finished, torch.zeros_like(sp), sp)
        ep = torch.where(finished, torch.zeros_like(ep), ep)
        got = ix.mask_logits(sp, ep, scores.cuda(), out=torch.empty_like(scores, device="cuda"), override_tok=override,
                             always_allow=eos if allow_eos else -1).cpu()
        assert torch.equal(torch.isnan(got), torch.isnan(exp))
        a, b = torch.nan_to_num(got.float(), nan=0.0), torch.nan_to_num(exp.float(), nan=0.0)
        assert torch.equal(a, b)
        assert torch.equal(torch.signbit(a), torch.signbit(b))  # -0.0 + 0.0 -> +0.0
    # empty history: S = occurring_distinct (vllm.py:62-66) == range [0, len) -> same kernel
    exp = hard_mask_reference(o, [[]], scores[:1], 1, eos, pad, 0, False)
    z = torch.zeros(1, dtype=torch.int64, device="cuda")
    got = ix.mask_logits(z, z + len(ix), scores[:1].cuda(), out=torch.empty_like(scores[:1], device="cuda")).cpu()
    assert torch.equal(torch.nan_to_num(got.float(), nan=0.0), torch.nan_to_num(exp.float(), nan=0.0))


def test_mask_logits_unaligned_and_inplace(cuda):
    rng = random.Random(2)
    docs = random_docs(rng, 900, 50, 10, 100)
    ix, o = make(docs), oracle(docs)
    V = 1001  # not a multiple of 4/8/32 -> scalar path, ragged last mask word
    rows = [rng.choice(docs)[:rng.randint(1, 3)] for _ in range(9)]
    from fm_index_constrained_decoding_b200.index import ragged
    toks, offs = ragged(rows)
    sp, ep = ix.get_range_batch(torch.from_numpy(toks).cuda(), torch.from_numpy(offs).cuda())
    x = torch.randn(9, V)
    y = x.clone().cuda()
    mo = torch.empty((9, (V + 31) // 32), dtype=torch.int32, device="cuda")
    ix.mask_logits(sp, ep, y, mask_out=mo)  # in place
    for k, r in enumerate(rows):
        allowed, _ = o.get_distinct_count(*o.get_range(r))
        exp = torch.full((V,), float("-inf"))
        exp[allowed] = x[k, allowed] + 0.0
        assert torch.equal(y[k].cpu(), exp)
        bits = [v for v in range(V) if (int(mo[k, v >> 5]) >> (v & 31)) & 1]
        assert bits == allowed


@pytest.mark.parametrize("dens", [(32, 64, False), (1, 1, True), (8, 16, False), (1, 1, False)])
def test_locate_extract_batches(cuda, dens):
    rng = random.Random(4)
    docs = random_docs(rng, 2000, 120, 10, 200)
    ix, o = make(docs, sa_dens=dens[0], isa_dens=dens[1], keep_text=dens[2]), oracle(docs)
    N = o.size()
    rows = [rng.randrange(N + 3) for _ in range(4000)]
    pos = ix.locate_batch(torch.tensor(rows, dtype=torch.int64, device="cuda")).cpu().numpy().astype(np.uint64)
    assert (pos == o.locate_batch(rows, nthreads=4)).all()
    docs_idx = ix.doc_index_batch(torch.from_numpy(pos[pos < N].astype(np.int64)).cuda()).tolist()
    assert docs_idx == [o.get_doc_index(int(p)) for p in pos[pos < N]]
    begins = [rng.randrange(N) for _ in range(300)]
    ends = [b + rng.randint(-2, 40) for b in begins]
    ends = [max(e, 0) for e in ends]
    out, lens = ix.extract_batch(torch.tensor(begins, dtype=torch.int64, device="cuda"),
                                 torch.tensor(ends, dtype=torch.int64, device="cuda"), 48)
    out, lens = out.cpu().tolist(), lens.cpu().tolist()
    for k in range(300):
        exp = list(o.extract_text(begins[k], ends[k]))
        assert lens[k] == len(exp) and out[k][:lens[k]] == exp, (begins[k], ends[k])
    assert [ix.get_doc(i) for i in range(0, len(docs), 17)] == [o.get_doc(i) for i in range(0, len(docs), 17)]
    assert list(ix.get_doc_indices(docs[3][:2])) == list(o.get_doc_indices(docs[3][:2]))


def test_keep_text_extract(cuda):
    rng = random.Random(6)
    docs = random_docs(rng, 500, 40, 5, 80)
    ix, o = make(docs, keep_text=True), oracle(docs)
    for i in range(len(docs)):
        assert ix.get_doc(i) == o.get_doc(i)


def test_save_load_roundtrip(cuda, tmp_path):
    from fm_index_constrained_decoding_b200 import FMIndex
    rng = random.Random(8)
    docs = random_docs(rng, 4000, 80, 10, 150)
    ix, o = make(docs), oracle(docs)
    ix.labels = ["d%d" % i for i in range(len(docs))]
    path = str(tmp_path / "corpus.fm_index")
    ix.save(path)
    assert os.path.exists(path + ".fmi") and os.path.exists(path + ".oth")
    ld = FMIndex.load(path)
    assert ld.size() == o.size() and ld.beginnings == ix.beginnings and ld.labels == ix.labels
    assert (ld.occurring_distinct, ld.occurring_counts) == (o.occurring_distinct, o.occurring_counts)
    for _ in range(50):
        d = rng.choice(docs)
        q = d[3:3 + rng.randint(1, 5)]
        assert ld.get_range(q) == o.get_range(q)
    assert [ld.locate(r) for r in range(0, o.size(), 97)] == [o.locate(r) for r in range(0, o.size(), 97)]


def test_edge_cases(cuda):
    from fm_index_constrained_decoding_b200 import FMIndex, _native, fm_index
    # zero symbol rejected like sdsl construct
    raw = fm_index.FMIndex()
    with pytest.raises(_native.FmidxError) as e:
        raw.initialize([4, 0, 5])
    assert e.value.code == _native.FMIDX_EZERO
    # empty corpus: only the sentinel
    raw.initialize([])
    assert raw.size() == 1 and raw.distinct_count(0, 1) == (0, 1) and raw.locate(0) == 0
    # single-symbol text with a long run
    ix, o = make([[7] * 700]), oracle([[7] * 700])
    assert ix.get_range([7] * 650) == o.get_range([7] * 650)
    assert ix.get_distinct_count(0, 701) == o.get_distinct_count(0, 701)
    # uninitialised handle
    with pytest.raises(_native.FmidxError):
        FMIndex().size()
    # SWIG-style argument errors
    with pytest.raises(OverflowError):
        ix.backward_search_step(-1, 0, 5)
    with pytest.raises(TypeError):
        ix.locate(1.5)
    # empty batch
    z = torch.zeros(0, dtype=torch.int32, device="cuda")
    sp, ep = ix.get_range_batch(z, torch.zeros(1, dtype=torch.int64, device="cuda"))
    assert sp.numel() == 0


def test_large_properties(cuda):
    """4M-token synthetic corpus (too slow for exhaustive oracle checks): size-independent
    properties + oracle spot checks through build-from-device."""
    from fm_index_constrained_decoding_b200 import FMIndex, synth
    toks, beg = synth.zipf_corpus(4_000_000, seed=0, device="cuda")
    ix = FMIndex(layout="sampled")
    ix.initialize_from_device_tokens(toks, beg)
    n = toks.numel()
    assert ix.size() == n + 1 and len(ix) == n
    q, offs = synth.sample_queries(toks, 4096, seed=1, beginnings=beg)
    sp, ep = ix.get_range_batch(q, offs)
    assert (ep > sp).all()  # positive queries always match
    # (1) range width == brute-force occurrence count for short patterns (direct scan of the stream)
    t64 = toks.long()
    lens = (offs[1:] - offs[:-1]).tolist()
    checked = 0
    for k in range(4096):
        if lens[k] <= 3 and checked < 40:
            pat = q[offs[k]:offs[k + 1]].long()
            hit = torch.ones(n - len(pat) + 1, dtype=torch.bool, device="cuda")
            for j in range(len(pat)):
                hit &= t64[j:n - len(pat) + 1 + j] == pat[j]
            assert int(hit.sum()) == int(ep[k] - sp[k])
            checked += 1
    # (2) every located occurrence really is the pattern; (3) mask bits == tokens following the occurrences
    beg_l = beg.tolist()
    for k in range(0, 4096, 131):
        w = int(ep[k] - sp[k])
        if w > 64:
            continue
        rows = torch.arange(int(sp[k]), int(ep[k]), device="cuda")
        pos = ix.locate_batch(rows)
        pat = q[offs[k]:offs[k + 1]].tolist()
        rev = [x + 10 for x in pat[::-1]]
        nxt = set()
        for p in pos.tolist():
            got = ix.extract_text(p + 1, p + len(pat) + 1)  # as-written: yields T[p+len-1 .. p]
            assert list(got) == rev[::-1], (k, p)
            prev = ix.extract_text(p, p + 1)[0] if p > 0 else 0  # symbol before the occurrence in T
            if prev > 0:
                nxt.add(prev - 10)
        m = ix.distinct_mask_batch(sp[k:k + 1], ep[k:k + 1], 128256)[0].cpu().numpy().view(np.uint32)
        bits = set(np.flatnonzero(np.unpackbits(m.view(np.uint8), bitorder="little")).tolist())
        assert bits == nxt
    # (4) masked logits: a row is finite exactly on the mask
    logits = torch.randn(64, 128256, device="cuda")
    m = ix.distinct_mask_batch(sp[:64], ep[:64], 128256)
    out = ix.mask_logits(sp[:64], ep[:64], logits.clone())
    bits = torch.from_numpy(np.unpackbits(m.cpu().numpy().view(np.uint8), axis=1, bitorder="little").astype(bool))
    assert torch.equal(torch.isfinite(out).cpu(), bits)
    assert torch.equal(out.cpu()[bits], (logits.cpu() + 0.0)[bits])
    assert beg_l[0] == 0


@pytest.mark.parametrize("layout", ["resident", "sampled"])
def test_sample_corpus_config0(cuda, layout):
    """BASELINE configs[0]: the reference's sample corpus (stand-in tokenizer), single-prefix search +
    distinct next-token query, batch 1, through the reference-shaped API; index built via the
    temporary int32 file exactly like index.py:57-66."""
    from conftest import load_golden
    from fm_index_constrained_decoding_b200 import FMIndex
    g = load_golden("sample_corpus.json")
    ix = FMIndex(layout=layout)
    ix.initialize(g["docs"], in_memory=False)
    ix.labels = g["labels"]
    assert ix.size() == g["size"] and ix.n_docs == 10
    for q in g["queries"]:
        r = ix.get_range(q["q"])
        assert list(r) == q["range"]
        assert ix.get_distinct_count(*r) == (q["distinct"], q["counts"])
        assert [ix.locate(x) for x in range(*r)][:16] == q["locate"]
        assert list(ix.get_doc_indices(q["q"]))[:16] == [ix.get_doc_index(p) for p in q["locate"]]
    assert [list(x) for x in (ix.occurring_distinct, ix.occurring_counts)] == g["occurring"]
    assert [ix.get_doc(i) for i in range(10)] == g["get_doc"]


@pytest.mark.parametrize("V", [70_000, 1_100_000, 9_000_000])
def test_large_alphabets_other_level_counts(cuda, V):
    """3, 4 and 4 nibble levels above the byte level (17-, 21- and 24-bit symbols): the top-descent depth and
    the node tables change."""
    from fm_index_constrained_decoding_b200.index import ragged
    rng = random.Random(V)
    docs = [[rng.randrange(V) for _ in range(rng.randint(5, 80))] for _ in range(60)]
    docs[7] = docs[3][2:30] + docs[7]
    docs.append([V - 1, V - 2, V - 1])
    ix, o = make(docs), oracle(docs)
    bits = (V + 9).bit_length()
    assert ix.info()["levels"] == 1 + (bits - 8 + 3) // 4
    seqs = [[]] + [d[i:i + k] for d in docs[:40] for i, k in ((0, 1), (1, 2), (2, 5))]
    toks, offs = ragged(seqs)
    sp_o, ep_o, m_o, _ = o.query_batch(toks, offs, want_mask=True, vocab=V, nthreads=4)
    sp, ep = ix.get_range_batch(torch.from_numpy(toks).cuda(), torch.from_numpy(offs).cuda())
    assert (sp.cpu().numpy().astype(np.uint64) == sp_o).all() and (ep.cpu().numpy().astype(np.uint64) == ep_o).all()
    mask = ix.distinct_mask_batch(sp, ep, V)
    assert (mask.cpu().numpy().view(np.uint32) == m_o).all()
    assert ix.get_distinct_count(0, ix.size()) == o.get_distinct_count(0, o.size())
    x = torch.randn(len(seqs), V)
    y = ix.mask_logits(sp, ep, x.cuda(), out=torch.empty_like(x, device="cuda")).cpu()
    bits = torch.from_numpy(np.unpackbits(m_o.view(np.uint8), axis=1, bitorder="little")[:, :V].astype(bool))
    assert torch.equal(torch.isfinite(y), bits)
    assert [ix.locate(r) for r in range(0, o.size(), 37)] == [o.locate(r) for r in range(0, o.size(), 37)]


def test_pipelined_host_api(cuda):
    """fmidx_query_submit / fmidx_query_wait: several host batches in flight, each bit-exact."""
    from fm_index_constrained_decoding_b200 import _native
    from fm_index_constrained_decoding_b200.index import ragged
    rng = random.Random(31)
    docs = random_docs(rng, 3000, 150, 10, 200)
    ix, o = make(docs), oracle(docs)
    L = _native.lib()
    batches, outs = [], []
    for s in range(6):
        seqs = [rng.choice(docs)[:rng.randint(1, 20)] for _ in range(300 + 17 * s)]
        toks, offs = ragged(seqs)
        B = len(seqs)
        res = (np.empty(B, np.uint64), np.empty(B, np.uint64), np.empty(B, np.uint32))
        batches.append((toks, offs, B))
        outs.append(res)
    for s, (toks, offs, B) in enumerate(batches):
        slot = s % 3
        _native.check(L.fmidx_query_wait(ix._h, slot))
        sp, ep, nd = outs[s]
        _native.check(L.fmidx_query_submit(ix._h, slot, toks.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p), B, 10, 3000,
                                           sp.ctypes.data_as(C.c_void_p), ep.ctypes.data_as(C.c_void_p), nd.ctypes.data_as(C.c_void_p)))
    for slot in range(3):
        _native.check(L.fmidx_query_wait(ix._h, slot))
    for s, (toks, offs, B) in enumerate(batches):
        osp, oep, _, ond = o.query_batch(toks, offs, want_mask=True, vocab=3000, nthreads=4)
        assert (outs[s][0] == osp).all() and (outs[s][1] == oep).all() and (outs[s][2] == ond.astype(np.uint32)).all()
    with pytest.raises(_native.FmidxError):
        _native.check(L.fmidx_query_wait(ix._h, 32))


_OVERFLOW_SCRIPT = r"""
import random, sys
import numpy as np, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
from test_gpu_parity import make, oracle, random_docs
from fm_index_constrained_decoding_b200.index import ragged
rng = random.Random(77)
V = 128256
docs = random_docs(rng, V, 300, 20, 400)
ix, o = make(docs), oracle(docs)
seqs = [[]] + [[rng.choice(rng.choice(docs))] for _ in range(40)] + [[]] * 3
toks, offs = ragged(seqs)
sp_o, ep_o, m_o, _ = o.query_batch(toks, offs, want_mask=True, vocab=V, nthreads=2)
sp, ep = ix.get_range_batch(torch.from_numpy(toks).cuda(), torch.from_numpy(offs).cuda())
mask, counts = ix.distinct_mask_batch(sp, ep, V, want_counts=True)
assert (mask.cpu().numpy().view(np.uint32) == m_o).all()
mask2 = ix.distinct_mask_batch(sp, ep, V)
assert (mask2.cpu().numpy().view(np.uint32) == m_o).all()
exp = o.get_distinct_count_multi(sp_o[:8].tolist(), ep_o[:8].tolist())
cn = counts.cpu().numpy()
for k in range(8):
    for t, c in zip(*exp[k]):
        assert cn[k, t] == c
print("OVERFLOW-OK")
"""


def test_mask_byte_node_list_overflow(cuda):
    """k_mask_expand evaluates byte-level nodes itself once the node list is full: force that path with a
    7-entry list (FMIDX_BNODE_CAP is read once per process, hence the subprocess)."""
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, FMIDX_BNODE_CAP="7")
    r = subprocess.run([sys.executable, "-c", _OVERFLOW_SCRIPT % {"root": os.path.dirname(here), "tests": here}],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OVERFLOW-OK" in r.stdout, r.stdout + r.stderr


def test_query_batch_two_streams(cuda):
    """fmidx_query_batch: search + masks in one call; two batches in flight on two streams with their own
    workspace slots must not disturb each other."""
    from fm_index_constrained_decoding_b200.index import ragged
    rng = random.Random(31)
    V = 128256
    docs = random_docs(rng, 3000, 400, 20, 400)
    ix, o = make(docs), oracle(docs)
    batches = []
    for k in range(2):
        seqs = [[]] + [rng.choice(docs)[:rng.randint(1, 3)] for _ in range(300)]
        for _ in range(300):
            d = rng.choice(docs)
            i = rng.randrange(len(d))
            seqs.append(d[i:i + rng.randint(1, 6)])
        toks, offs = ragged(seqs)
        batches.append((toks, offs, o.query_batch(toks, offs, want_mask=True, vocab=V, nthreads=4)))
    streams = [torch.cuda.Stream() for _ in range(2)]
    dev = [(torch.from_numpy(t).cuda(), torch.from_numpy(f).cuda()) for t, f, _ in batches]
    torch.cuda.synchronize()
    for rep in range(6):
        outs = []
        for k in range(2):
            with torch.cuda.stream(streams[k]):
                outs.append(ix.query_batch(dev[k][0], dev[k][1], V, ws_slot=k))
        torch.cuda.synchronize()
        for k in range(2):
            sp_o, ep_o, m_o, _ = batches[k][2]
            sp, ep, mask = outs[k]
            assert (sp.cpu().numpy().astype(np.uint64) == sp_o).all() and (ep.cpu().numpy().astype(np.uint64) == ep_o).all()
            assert (mask.cpu().numpy().view(np.uint32) == m_o).all()


def test_gather_peak_probe(cuda):
    """fmidx_gather_peak (measurement aid): runs on a small index and returns a positive rate."""
    from fm_index_constrained_decoding_b200 import _native
    rng = random.Random(5)
    ix = make(random_docs(rng, 5000, 200, 50, 400))
    for unit in (32, 128):
        v = C.c_double()
        _native.check(_native.lib().fmidx_gather_peak(ix._h, unit, 1 << 20, C.byref(v)))
        assert v.value > 0.0
    with pytest.raises(_native.FmidxError):
        _native.check(_native.lib().fmidx_gather_peak(ix._h, 64, 1 << 20, C.byref(C.c_double())))


def test_build_index_cli_roundtrip(cuda, tmp_path):
    """build_index (reference build_fm_index.py command line): TSV -> index files -> load -> same answers as an index
    initialised from the same token lists."""
    from fm_index_constrained_decoding_b200 import FMIndex, build_index as bi
    tsv = tmp_path / "c.tsv"
    tsv.write_text("".join("%d\tT%d\tdoc %d says hello world number %d again hello\n" % (i, i, i % 3, i % 5) for i in range(12)),
                   encoding="utf-8")
    bi.main([str(tsv), str(tmp_path / "idx"), "--standin_tokenizer", "--include_title"])
    ix = FMIndex.load(str(tmp_path / "idx"))
    docs = [bi.standin_encode(t) for _, t in bi.iter_documents(str(tsv), include_title=True)]
    ref = make(docs)
    assert ix.labels == [str(i) for i in range(12)] and len(ix) == len(ref) and ix.beginnings == ref.beginnings
    for q in (docs[0][:1], docs[0][1:3], docs[5][2:6], [123]):
        assert ix.get_range(q) == ref.get_range(q)
        assert ix.get_distinct_count(*ix.get_range(q)) == ref.get_distinct_count(*ref.get_range(q))
