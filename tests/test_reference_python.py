"""CPU suite: the vectors in tests/golden/reference_python.json are OUTPUTS OF THE REFERENCE'S OWN PYTHON
(/root/reference/index.py and the document_constrained_generation_* processors imported unmodified over oracle/ref_shim;
generator: tests/golden/make_reference_golden.py).  Here:
  * when /root/reference is present (the build container) the vectors are regenerated and must equal the committed file;
  * the oracle's restatement of that Python layer (oracle/oracle.py: OracleFMIndex, hard_mask_reference,
    soft_boost_reference) must reproduce every vector -- so the GPU parity tests, which use those restatements as
    their checker at other sizes, are anchored to the reference's own answers.
The product is checked against the same vectors in tests/test_reference_golden_gpu.py."""
import json
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import refgolden_common as G  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "reference_python.json")))
DT = {"float32": torch.float32, "bfloat16": torch.bfloat16}


def corpus_docs(name):
    if name == "worked_example":
        return [[5, 6, 7], [5, 6, 8], [6, 7]]
    if name == "unicorn_paragraph":
        return GOLD["corpora"]["unicorn_paragraph"]
    if name == "sample_corpus":
        return json.load(open(os.path.join(HERE, "golden", "sample_corpus.json")))["docs"]
    if name == "random_4000":
        return G.random_corpus(11)
    if name == "random_wide_vocab":
        return G.random_corpus(12, n_docs=40, vocab=128000)
    raise KeyError(name)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree only exists in the build container")
def test_committed_vectors_are_what_the_reference_answers():
    import subprocess
    code = ("import sys, json; sys.path.insert(0, %r); import make_reference_golden as m; "
            "print(json.dumps(m.generate(), separators=(',', ':')))" % os.path.join(HERE, "golden"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout
    assert json.loads(out.strip().splitlines()[-1]) == GOLD


def check_index_vectors(ix, v, check_get_doc=True):
    """Every method of index.py:20-207 against one golden case (used for the oracle here and for the product on the GPU)."""
    assert ix.size() == v["size"] and len(ix) == v["len"] and ix.n_docs == v["n_docs"]
    assert list(ix.beginnings) == v["beginnings"] and sorted(ix.occurring) == v["occurring_sorted"]
    assert list(ix.occurring_distinct) == v["occurring_distinct"] and list(ix.occurring_counts) == v["occurring_counts"]
    for q in v["queries"]:
        lo, hi = ix.get_range(q["q"])
        assert [lo, hi] == q["range"], q
        assert ix.get_count(q["q"]) == q["count"]
        assert list(ix.get_continuations(q["q"])) == q["continuations"]
        d, c = ix.get_distinct_count(lo, hi)
        assert list(d) == q["distinct"] and list(c) == q["counts"]
        assert list(ix.get_distinct(lo, hi)) == q["distinct"]
        assert list(ix.get_doc_indices(q["q"]))[:24] == q["doc_indices"]
        assert [ix.get_token_index_from_row(r) for r in range(lo, min(hi, lo + 8))] == q["token_index"]
        assert [ix.get_doc_index_from_row(r) for r in range(lo, min(hi, lo + 8))] == q["doc_indices"][:8]
    multi = ix.get_distinct_count_multi([q["range"][0] for q in v["queries"]], [q["range"][1] for q in v["queries"]])
    assert [list(m[0]) for m in multi] == [q["distinct"] for q in v["queries"]]
    assert [list(m[1]) for m in multi] == [q["counts"] for q in v["queries"]]
    if check_get_doc:
        assert [list(ix.get_doc(i)) for i in range(len(v["get_doc"]))] == v["get_doc"]
    assert [ix.get_doc_length(i) for i in range(len(v["doc_length"]))] == v["doc_length"]
    for sym, lo, hi, want in v["steps"]:
        assert list(ix.backward_search_step(sym, lo, hi)) == want, (sym, lo, hi)
    assert ix.locate(v["size"]) == v["locate_out_of_range"] == 2 ** 64 - 1
    for lo, hi, want in v["distinct_clamps"]:
        assert list(ix.distinct(lo, hi)) == want


@pytest.mark.parametrize("case", range(len(GOLD["indexes"])))
def test_oracle_index_layer_matches_reference_python(case):
    from oracle.oracle import OracleFMIndex
    v = GOLD["indexes"][case]
    o = OracleFMIndex()
    o.initialize(corpus_docs(v["name"]))
    check_index_vectors(o, v)


def _oracle_index():
    from oracle.oracle import OracleFMIndex
    o = OracleFMIndex()
    o.initialize(G.random_corpus(11))
    return o


def hard_mask_rows(v, style):
    """(history rows, scores) of one golden hard-mask case, with forced-BOS handled like the reference does."""
    dt = DT[v["dtype"]]
    if style == "vllm":
        hs = [list(h) for h in v["histories"]]
        scores = torch.cat([G.make_scores(1, dt, v["salt0"] + i) for i in range(len(hs))])
        return hs, scores
    ids = v["input_ids"]
    return [r[1:] for r in ids], G.make_scores(len(ids), dt, v["salt"])


@pytest.mark.parametrize("style", ["vllm", "seq"])
def test_oracle_hard_mask_restatement_matches_reference_python(style):
    from oracle.oracle import hard_mask_reference
    o = _oracle_index()
    for v in GOLD["processors"][style]:
        cfg = v["cfg"]
        hs, scores = hard_mask_rows(v, style)
        bos = cfg.get("forced_bos_token_id")
        got = []
        for i, h in enumerate(hs):  # row by row: the vllm variant is per request, and forced-BOS rows differ in length
            s = scores[i:i + 1]
            if bos is not None and len(h) == 0:
                m = torch.full_like(s, float("-inf"))
                m[:, bos] = 0.0
                got.append(s + m)
                continue
            if bos is not None:
                h = h[1:]
            got.append(hard_mask_reference(o, [h], s, 1, G.EOS, G.PAD, cfg.get("stop_at_count", 0), cfg.get("always_allow_eos", False),
                                           cfg.get("force_decoding_from")))
        got = torch.cat(got)
        assert G.row_digest(got) == v["digest"], (style, cfg, v["dtype"])
        assert G.finite_positions(got) == v["allowed"]


@pytest.mark.parametrize("variant", ["causal", "qwen"])
def test_oracle_soft_boost_restatement_matches_reference_python(variant):
    from oracle.oracle import soft_boost_reference
    o = _oracle_index()
    for v in GOLD["processors"][variant]:
        cfg = v["cfg"]
        ids = v["input_ids"]
        scores = G.make_scores(len(ids), DT[v["dtype"]], v["salt"], specials=False)
        marker = G.QWEN_MARK if variant == "qwen" else G.LLAMA_MARK
        got = soft_boost_reference(o, ids, scores, cfg.get("num_beams", 1), variant, G.EOS, G.PAD, marker, cfg.get("stop_at_count", 0),
                                   cfg.get("always_allow_eos", False), cfg.get("force_decoding_from"), cfg.get("forced_bos_token_id"),
                                   2.0, cfg.get("min_new_tokens", 5))
        assert G.row_digest(got) == v["digest"], (variant, cfg, v["dtype"])


def replay_beam_search(case):
    """Yields (step, input_ids rows incl. the decoder-start column, scores, parents, expected digest) of a stored search."""
    ci = GOLD["beam_search"].index(case)
    R = case["rows"]
    ids = [[2] for _ in range(R)]
    for step, st in enumerate(case["steps"]):
        yield step, ids, G.make_scores(R, torch.float32, 6000 + 100 * ci + step), st
        ids = [ids[p] + [t] for p, t in zip(st["parents"], st["tokens"])]


def test_oracle_hard_mask_restatement_replays_the_reference_beam_search():
    from oracle.oracle import hard_mask_reference
    o = _oracle_index()
    for case in GOLD["beam_search"]:
        cfg = case["cfg"]
        for step, ids, scores, st in replay_beam_search(case):
            got = hard_mask_reference(o, [r[1:] for r in ids], scores, case["num_beams"], G.EOS, G.PAD, cfg.get("stop_at_count", 0),
                                      cfg.get("always_allow_eos", False), cfg.get("force_decoding_from"))
            assert G.row_digest(got) == st["digest"], (cfg, step)
