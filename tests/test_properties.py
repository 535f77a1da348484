"""CPU suite: hypothesis property tests -- the C oracle against the brute-force model and against the
defining properties of an FM-index, on arbitrary small corpora (incl. tiny alphabets and long runs)."""
from hypothesis import given, settings, strategies as st

from oracle.brute import BruteFMIndex
from oracle.oracle import OracleFMIndex

docs_st = st.lists(st.lists(st.integers(0, 6), min_size=1, max_size=25), min_size=1, max_size=6)
wide_docs_st = st.lists(st.lists(st.integers(0, 128255), min_size=1, max_size=12), min_size=1, max_size=4)


@settings(max_examples=120, deadline=None)
@given(docs_st, st.data())
def test_ranges_and_distinct_match_definition(docs, data):
    b, o = BruteFMIndex(docs), OracleFMIndex()
    o.initialize(docs)
    N = o.size()
    assert sorted(o.locate(r) for r in range(N)) == list(range(N))  # SA is a permutation
    d = data.draw(st.sampled_from(docs))
    i = data.draw(st.integers(0, len(d) - 1))
    L = data.draw(st.integers(1, 8))
    q = d[i:i + L] + ([data.draw(st.integers(0, 7))] if data.draw(st.booleans()) else [])
    r = o.get_range(q)
    assert r == b.get_range(q) == b.get_range_by_definition(q)
    toks, cnts = o.get_distinct_count(*r)
    assert toks == sorted(toks) and all(c > 0 for c in cnts)
    raw = o.distinct_count(*r)
    assert sum(raw[1::2]) == r[1] - r[0]                             # counts partition the range
    for t, c in zip(toks, cnts):
        assert o.get_count(q + [t]) == c if r != (0, 0) else True    # each continuation extends the match
    lo = data.draw(st.integers(0, N + 1))
    hi = data.draw(st.integers(0, N + 2))
    assert o.distinct_count(lo, hi) == b.distinct_count(lo, hi)
    assert o.extract_text(min(lo, N - 1), hi) == b.extract_text(min(lo, N - 1), hi)


@settings(max_examples=40, deadline=None)
@given(wide_docs_st)
def test_llama_sized_alphabet(docs):
    b, o = BruteFMIndex(docs), OracleFMIndex()
    o.initialize(docs)
    assert [o.locate(r) for r in range(o.size())] == b.SA
    assert o.get_distinct_count(0, len(o)) == b.get_distinct_count(0, len(b))
    assert [o.get_doc(i) for i in range(len(docs))] == [b.get_doc(i) for i in range(len(docs))]
    for d in docs:
        assert o.get_range(d) == b.get_range_by_definition(d)
