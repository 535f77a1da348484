"""Generates the golden fixtures in this directory from the brute-force model (oracle/brute.py).

    python tests/golden/make_golden.py

The reference ships no golden vectors for the FM-index path and cannot be built here (no
sdsl-lite), so these vectors are DERIVED: sorted suffixes + list counting, i.e. the definition
of every quantity.  worked_example.json is the hand-checkable case of SURVEY.md section 9.
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.brute import BruteFMIndex  # noqa: E402


def case(docs, queries, rng):
    b = BruteFMIndex(docs)
    N = b.size()
    out = {"docs": docs, "size": N, "SA": b.SA, "BWT": b.BWT, "beginnings": b.beginnings, "queries": []}
    for q in queries:
        r = b.get_range(q)
        d, c = b.get_distinct_count(*r)
        out["queries"].append({"q": q, "range": list(r), "distinct": d, "counts": c,
                               "locate": [b.locate(x) for x in range(*r)]})
    out["occurring"] = [list(x) for x in b.occurring_distinct_counts()]
    out["get_doc"] = [b.get_doc(i) for i in range(len(docs))]
    out["locate_oob"] = b.locate(N)
    steps = []
    syms = sorted(set(b.T)) + [1, 9, max(b.T) + 1, max(b.T) + 1000]
    for _ in range(40):
        s = rng.choice(syms)
        lo = rng.randint(0, N + 1)
        hi = rng.randint(0, N + 2)
        steps.append({"sym": s, "low": lo, "high": hi, "out": list(b.backward_search_step(s, lo, hi))})
    out["steps"] = steps
    dcs = []
    for _ in range(30):
        lo = rng.randint(0, N + 1)
        hi = rng.randint(0, N + 3)
        dcs.append({"low": lo, "high": hi, "out": list(b.distinct_count(lo, hi))})
    out["distinct_count"] = dcs
    ext = []
    for _ in range(30):
        be = rng.randint(0, N - 1)
        en = rng.randint(0, N + 3)
        ext.append({"begin": be, "end": en, "out": list(b.extract_text(be, en))})
    out["extract"] = ext
    return out


def main():
    rng = random.Random(20261018)
    docs = [[5, 6, 7], [5, 6, 8], [6, 7]]
    qs = [[5], [5, 6], [6], [6, 7], [5, 6, 7], [7, 5], [8, 5], [9], []]
    with open(os.path.join(HERE, "worked_example.json"), "w") as f:
        json.dump(case(docs, qs, rng), f)
    cases = []
    for V in [2, 3, 7, 40, 300, 128256]:
        for _ in range(3):
            nd = rng.randint(1, 7)
            docs = [[rng.randrange(V) for _ in range(rng.randint(1, 40))] for _ in range(nd)]
            qs = [[]]
            for _ in range(25):
                d = rng.choice(docs)
                i = rng.randrange(len(d))
                q = d[i:i + rng.randint(1, 6)]
                if rng.random() < 0.3:
                    q = q + [rng.randrange(V)]
                qs.append(q)
            cases.append(case(docs, qs, rng))
    with open(os.path.join(HERE, "random_small.json"), "w") as f:
        json.dump(cases, f)
    print("wrote", len(cases) + 1, "cases")


if __name__ == "__main__":
    main()
