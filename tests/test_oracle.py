"""CPU suite: pins the C oracle (oracle/fm_oracle.c) against the golden fixtures and against the
independent brute-force model on fresh random corpora."""
import random

import numpy as np
import pytest

from conftest import golden_cases
from oracle.brute import BruteFMIndex
from oracle.oracle import OracleFMIndex, NativeOracle, lib

U64_MAX = (1 << 64) - 1


def test_worked_example_constants():
    """SURVEY.md section 9, by hand."""
    o = OracleFMIndex()
    o.initialize([[5, 6, 7], [5, 6, 8], [6, 7]])
    assert o.size() == 9 and len(o) == 8 and o.beginnings == [0, 3, 6, 8]
    assert o.get_range([5]) == (1, 3)
    assert o.get_range([5, 6]) == (4, 6)
    assert o.get_range([6, 7]) == (6, 8)
    assert o.get_range([7, 5]) == (1, 2)     # cross-document match
    assert o.get_range([9]) == (0, 0)
    assert o.get_range([]) == (0, 9)
    assert o.get_distinct_count(3, 6) == ([7, 8], [2, 1])
    assert o.get_distinct_count(7, 8) == ([], [])    # only the sentinel follows
    assert (o.occurring_distinct, o.occurring_counts) == ([5, 6, 7, 8], [1, 3, 2, 1])
    assert [o.get_doc(i) for i in range(3)] == [[6, 7, -10], [6, 8, 5], [7, 5]]  # as-written off-by-one
    assert o.locate(9) == U64_MAX
    assert [o.locate(r) for r in range(9)] == [8, 5, 2, 7, 4, 1, 6, 0, 3]


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: "n%d" % c["size"])
def test_oracle_matches_golden(case):
    o = OracleFMIndex()
    o.initialize(case["docs"])
    assert o.size() == case["size"] and o.beginnings == case["beginnings"]
    assert [o.locate(r) for r in range(case["size"])] == case["SA"]
    assert o.locate(case["size"]) == case["locate_oob"]
    for q in case["queries"]:
        r = o.get_range(q["q"])
        assert list(r) == q["range"], q
        assert o.get_distinct_count(*r) == (q["distinct"], q["counts"])
        assert [o.locate(x) for x in range(*r)] == q["locate"]
    assert [list(x) for x in (o.occurring_distinct, o.occurring_counts)] == case["occurring"]
    assert [o.get_doc(i) for i in range(len(case["docs"]))] == case["get_doc"]
    for s in case["steps"]:
        assert list(o.backward_search_step(s["sym"], s["low"], s["high"])) == s["out"], s
    for d in case["distinct_count"]:
        assert list(o.distinct_count(d["low"], d["high"])) == d["out"], d
    for e in case["extract"]:
        assert list(o.extract_text(e["begin"], e["end"])) == e["out"], e


def test_oracle_vs_brute_random():
    rng = random.Random(7)
    for _ in range(60):
        V = rng.choice([2, 4, 30, 5000, 128256])
        docs = [[rng.randrange(V) for _ in range(rng.randint(1, 50))] for _ in range(rng.randint(1, 8))]
        b, o = BruteFMIndex(docs), OracleFMIndex()
        o.initialize(docs)
        N = b.size()
        for _ in range(20):
            d = rng.choice(docs)
            i = rng.randrange(len(d))
            q = d[i:i + rng.randint(1, 8)]
            if rng.random() < 0.25:
                q = q + [rng.randrange(V)]
            r = o.get_range(q)
            assert r == b.get_range(q) == b.get_range_by_definition(q)
            assert o.distinct_count(*r) == b.distinct_count(*r)
            lo, hi = rng.randint(0, N + 1), rng.randint(0, N + 2)
            assert o.distinct_count(lo, hi) == b.distinct_count(lo, hi)
            assert o.distinct(lo, hi) == b.distinct(lo, hi)


def test_suffix_array_medium():
    """SA-IS against numpy on a text with long repeats (periodic + planted)."""
    rng = np.random.default_rng(3)
    base = rng.integers(1, 50, size=2000)
    data = np.concatenate([base, base, rng.integers(1, 5, size=3000), base[:700], np.full(500, 3)]).astype(np.uint64)
    out = np.empty(data.size + 1, dtype=np.uint32)
    assert lib().fmo_suffix_array(data.ctypes.data, data.size, out.ctypes.data) == 0
    T = list(data.astype(np.int64)) + [0]
    ref = sorted(range(len(T)), key=lambda i: T[i:])
    assert out.tolist() == ref


def test_zero_symbol_rejected():
    n = NativeOracle()
    with pytest.raises(ValueError):
        n.initialize([3, 0, 4])


def test_backward_search_multi_matches_steps():
    rng = random.Random(11)
    docs = [[rng.randrange(6) for _ in range(40)] for _ in range(4)]
    o = OracleFMIndex()
    o.initialize(docs)
    for _ in range(50):
        d = rng.choice(docs)
        i = rng.randrange(len(d) - 4)
        q = [t + 10 for t in d[i:i + rng.randint(1, 4)]]
        l, r1 = o.backward_search_multi(q)
        exp = o.get_range([t - 10 for t in q])
        if exp != (0, 0):
            assert (l, r1) == exp
        else:
            assert l == r1  # raw arithmetic keeps an empty range empty (fm_index.cpp:55-65)


def test_query_batch_threads_agree():
    rng = np.random.default_rng(5)
    docs = [rng.integers(0, 300, size=rng.integers(5, 200)).tolist() for _ in range(30)]
    o = OracleFMIndex()
    o.initialize(docs)
    seqs = []
    for _ in range(200):
        d = docs[rng.integers(len(docs))]
        i = rng.integers(len(d))
        seqs.append(d[i:i + rng.integers(1, 6)])
    from fm_index_constrained_decoding_b200.index import ragged
    toks, offs = ragged(seqs)
    sp1, ep1, m1, n1 = o.query_batch(toks, offs, want_mask=True, vocab=320, nthreads=1)
    sp4, ep4, m4, n4 = o.query_batch(toks, offs, want_mask=True, vocab=320, nthreads=4)
    assert (sp1 == sp4).all() and (ep1 == ep4).all() and (m1 == m4).all() and (n1 == n4).all()
    for k, s in enumerate(seqs):
        assert (int(sp1[k]), int(ep1[k])) == o.get_range(s)
        d, _ = o.get_distinct_count(int(sp1[k]), int(ep1[k]))
        bits = [v for v in range(320) if (m1[k, v >> 5] >> (v & 31)) & 1]
        assert bits == d


def test_sample_corpus_config0_golden():
    """BASELINE configs[0]: test_data/sample_corpus.tsv through a stand-in tokenizer (see
    tests/golden/make_sample_corpus_golden.py); expectations from the brute-force model."""
    from conftest import load_golden
    g = load_golden("sample_corpus.json")
    o = OracleFMIndex()
    o.initialize(g["docs"])
    assert o.size() == g["size"] and o.n_docs == 10
    for q in g["queries"]:
        r = o.get_range(q["q"])
        assert list(r) == q["range"]
        assert o.get_distinct_count(*r) == (q["distinct"], q["counts"])
        assert [o.locate(x) for x in range(*r)][:16] == q["locate"]
    assert [list(x) for x in (o.occurring_distinct, o.occurring_counts)] == g["occurring"]
    assert [o.get_doc(i) for i in range(10)] == g["get_doc"]
