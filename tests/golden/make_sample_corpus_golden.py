"""BASELINE.json configs[0]: the reference's own CPU-runnable case, test_data/sample_corpus.tsv.

    python tests/golden/make_sample_corpus_golden.py      (needs /root/reference; run in the build container)

The Llama-3.2 tokenizer files are not available offline, so a documented STAND-IN tokenizer is used:
whitespace split, 64-bit FNV-1a hash of each word mod 128,000 (ids below the special-token range),
"Paraphrase: " prepended and EOS 128009 appended per document exactly like
/root/reference/build_fm_index.py:46-48.  The shipped TSV uses the two characters backslash-t
instead of TAB, which the reference's own loader would drop; they are
converted to TABs here.  Expected values come from the brute-force model (oracle/brute.py).
Only token ids and expectations are written -- no reference text is copied into the repo.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.brute import BruteFMIndex  # noqa: E402

SRC = "/root/reference/test_data/sample_corpus.tsv"
EOS = 128009


def fnv1a(word):
    h = 0xCBF29CE484222325
    for b in word.encode("utf-8"):
        h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h % 128000


def main():
    docs, labels = [], []
    for line in open(SRC, encoding="utf-8"):
        pp = line.strip().replace("\\t", "\t").split("\t", 2)
        if len(pp) != 3:
            continue
        labels.append(pp[0].strip())
        text = " ".join(pp[2].split())
        docs.append([fnv1a(w) for w in ("Paraphrase: " + text).split()] + [EOS])
    b = BruteFMIndex(docs)
    queries = []
    for d in docs[:4]:
        for start in (1, 7, 40):
            for L in (1, 2, 3, 5, 8):
                q = d[start:start + L]
                r = b.get_range(q)
                dist, cnt = b.get_distinct_count(*r)
                queries.append({"q": q, "range": list(r), "distinct": dist, "counts": cnt,
                                "locate": [b.locate(x) for x in range(*r)][:16]})
    out = {"source": "test_data/sample_corpus.tsv (stand-in tokenizer: FNV-1a word hash mod 128000)", "labels": labels,
           "docs": docs, "size": b.size(), "queries": queries,
           "occurring": [list(x) for x in b.occurring_distinct_counts()],
           "get_doc": [b.get_doc(i) for i in range(len(docs))]}
    with open(os.path.join(HERE, "sample_corpus.json"), "w") as f:
        json.dump(out, f)
    print("docs", len(docs), "tokens", b.size() - 1, "queries", len(queries))


if __name__ == "__main__":
    main()
