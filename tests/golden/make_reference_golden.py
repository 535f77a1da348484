"""Runs the REFERENCE'S OWN PYTHON, unmodified, and commits what it answers as golden vectors.

    python tests/golden/make_reference_golden.py        (build container only: needs /root/reference)

/root/reference/index.py and the document_constrained_generation_{vllm,seq,causal,causal_qwen}.py processors are imported
as they are; the two modules they need and this image lacks -- the SWIG extension `cpp_modules.fm_index` (sdsl-lite) and
`logits_processor_zoo` -- come from oracle/ref_shim (the C oracle behind the SWIG surface; see its README).  So every
convention of the reference's Python layer (SHIFT, reversal, `d > 0`, occurring_distinct over [0, n), `[::-1]` / `pop()`
bookkeeping, `scores + mask`) in tests/golden/reference_python.json is an output of the reference itself; what stays
restated underneath is sdsl-lite's C++.  Inputs are rebuilt from integer formulas (refgolden_common.py), so only
results are stored; no reference text or code is copied.
"""
import importlib
import json
import os
import sys
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
sys.path.insert(0, HERE)
import refgolden_common as G  # noqa: E402


def import_reference():
    """Reference modules, unmodified, over the oracle-backed shim."""
    for p in (os.path.join(ROOT, "oracle", "ref_shim"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    for m in ("cpp_modules", "cpp_modules.fm_index", "index"):
        sys.modules.pop(m, None)
    mods = {name: importlib.import_module(name) for name in
            ("index", "document_constrained_generation_vllm", "document_constrained_generation_seq",
             "document_constrained_generation_causal", "document_constrained_generation_causal_qwen")}
    assert mods["index"].__file__.startswith(REF), mods["index"].__file__
    assert sys.modules["cpp_modules.fm_index"].__file__.startswith(os.path.join(ROOT, "oracle", "ref_shim"))
    return mods


def corpora():
    unicorn_words = open(os.path.join(REF, "index_for_vllm.py"), encoding="utf-8").read()
    # index_for_vllm.py:6-24 (commented-out demo corpus): the paragraph between the triple quotes
    a = unicorn_words.index('"""') + 3
    b = unicorn_words.index('"""', a)
    para = " ".join(l.lstrip("# ").strip() for l in unicorn_words[a:b].splitlines()).split()
    sample = json.load(open(os.path.join(HERE, "sample_corpus.json")))["docs"]
    return {
        "worked_example": [[5, 6, 7], [5, 6, 8], [6, 7]],
        "unicorn_paragraph": [[G.fnv1a(w) for w in para] + [G.EOS]],                    # index_for_vllm.py:30-45 shape
        "sample_corpus": sample,                                                       # BASELINE configs[0]
        "random_4000": G.random_corpus(11),
        "random_wide_vocab": G.random_corpus(12, n_docs=40, vocab=128000),
    }


def queries_for(docs, seed):
    r = G.lcg(seed)
    qs = [[]]
    for _ in range(40):
        d = docs[r(len(docs))]
        i = r(max(1, len(d) - 1))
        L = 1 + r(8)
        q = d[i:i + L]
        kind = r(6)
        if kind == 0 and q:
            q = q[:-1] + [r(128256)]            # corrupted tail -> mostly (0, 0)
        elif kind == 1 and len(docs) > 1:
            q = docs[(r(len(docs)))][-2:] + q[:1]  # straddles a document boundary in forward order
        qs.append(q)
    return qs


def index_vectors(ref_index, name, docs, in_memory):
    ix = ref_index.FMIndex()
    ix.initialize([list(d) for d in docs], in_memory=in_memory)
    out = {"name": name, "in_memory": in_memory, "size": ix.size(), "len": len(ix), "n_docs": ix.n_docs,
           "beginnings": list(ix.beginnings), "occurring_sorted": sorted(ix.occurring),
           "occurring_distinct": list(ix.occurring_distinct), "occurring_counts": list(ix.occurring_counts)}
    qs = queries_for(docs, len(name))
    qv = []
    for q in qs:
        lo, hi = ix.get_range(q)
        dist, cnt = ix.get_distinct_count(lo, hi)
        qv.append({"q": q, "range": [lo, hi], "count": ix.get_count(q), "continuations": ix.get_continuations(q),
                   "distinct": dist, "counts": cnt, "doc_indices": list(ix.get_doc_indices(q))[:24],
                   "token_index": [ix.get_token_index_from_row(r) for r in range(lo, min(hi, lo + 8))]})
    out["queries"] = qv
    multi = ix.get_distinct_count_multi([v["range"][0] for v in qv], [v["range"][1] for v in qv])
    assert [list(m[0]) for m in multi] == [v["distinct"] for v in qv] and [list(m[1]) for m in multi] == [v["counts"] for v in qv]
    n_show = min(ix.n_docs, 6)
    out["get_doc"] = [ix.get_doc(i) for i in range(n_show)]
    out["doc_length"] = [ix.get_doc_length(i) for i in range(n_show)]
    r = G.lcg(99 + len(name))
    steps = []
    N = ix.size()
    for _ in range(24):
        sym, lo = 10 + r(200 if name == "worked_example" else 4100), r(N + 2)
        hi = lo + r(N) if r(4) else r(N + 3)
        steps.append([sym, lo, hi, list(ix.backward_search_step(sym, lo, hi))])
    out["steps"] = steps
    out["locate_out_of_range"] = ix.locate(N)
    out["distinct_clamps"] = [[lo, hi, list(ix.distinct(lo, hi))] for lo, hi in ((0, N + 5), (3, 3), (5, 2), (N, N + 1))]
    # save / load round trip through index.py:189-207 (the shim's own container underneath)
    with tempfile.TemporaryDirectory() as tmp:
        ix.save(os.path.join(tmp, "x"))
        ld = ref_index.FMIndex.load(os.path.join(tmp, "x"))
        assert ld.occurring_distinct == ix.occurring_distinct and ld.beginnings == ix.beginnings and ld.size() == ix.size()
    return out, ix


def histories(docs, n, L, seed, finish=True):
    r = G.lcg(seed)
    rows = []
    for k in range(n):
        d = docs[r(len(docs))]
        i = r(max(1, len(d) - L))
        h = list(d[i:i + L])
        if len(h) < L:
            h = (h + [G.PAD] * L)[:L]
        if finish and k % 7 == 1:
            h[-1] = G.EOS
        elif finish and k % 7 == 2:
            h[-1] = G.PAD
        elif k % 7 == 3:
            h[-1] = 3999 - (k % 5)  # (almost surely) not a continuation
        rows.append(h)
    return rows


def processor_vectors(mods, ix, docs):
    vllm_cls = mods["document_constrained_generation_vllm"].IndexBasedLogitsProcessor
    seq_cls = mods["document_constrained_generation_seq"].IndexBasedLogitsProcessor
    causal_cls = mods["document_constrained_generation_causal"].IndexBasedLogitsProcessor
    qwen_cls = mods["document_constrained_generation_causal_qwen"].IndexBasedLogitsProcessor
    out = {"vllm": [], "seq": [], "causal": [], "qwen": []}
    dts = {"float32": torch.float32, "bfloat16": torch.bfloat16}
    force = docs[3][10:11]
    # ---- document_constrained_generation_vllm.py: one request per call, num_beams = 1 (its .view needs that) ----------
    cfgs = [dict(), dict(always_allow_eos=True), dict(always_allow_eos=True, stop_at_count=2),
            dict(always_allow_eos=True, force_decoding_from=force), dict(forced_bos_token_id=77, always_allow_eos=True)]
    for ci, cfg in enumerate(cfgs):
        proc = vllm_cls(index=ix, num_beams=1, pad_token_id=G.PAD, eos_token_id=G.EOS, **cfg)
        hs = [[]] + histories(docs, 7, 1 + ci, 300 + ci)
        if cfg.get("force_decoding_from"):
            hs = [[]] + [docs[3][11:11 + k] for k in range(1, 6)]
        if cfg.get("forced_bos_token_id") is not None:
            hs = [[]] + [[77] + h for h in hs[1:]]
        for dn, dt in dts.items():
            rows = []
            for hi_, h in enumerate(hs):
                s = G.make_scores(1, dt, 1000 + 10 * ci + hi_)[0]
                rows.append(proc(tuple(h), s).reshape(1, -1))  # the reference returns [1, V] for a [V] input
            res = torch.cat(rows)
            out["vllm"].append({"cfg": cfg, "dtype": dn, "histories": hs, "salt0": 1000 + 10 * ci,
                                "digest": G.row_digest(res), "allowed": G.finite_positions(res)})
    # ---- document_constrained_generation_seq.py: HF batch style, decoder-start token in column 0, beams > 1 -----------
    cfgs = [dict(num_beams=2), dict(num_beams=3, always_allow_eos=True), dict(num_beams=2, always_allow_eos=True, stop_at_count=2),
            dict(num_beams=2, always_allow_eos=True, force_decoding_from=force), dict(num_beams=2, forced_bos_token_id=77)]
    for ci, cfg in enumerate(cfgs):
        nb = cfg["num_beams"]
        proc = seq_cls(index=ix, pad_token_id=G.PAD, eos_token_id=G.EOS, **cfg)
        for L in (0, 1, 4):
            n = nb * 2
            hs = histories(docs, n, L, 500 + ci * 7 + L) if L else [[] for _ in range(n)]
            if cfg.get("force_decoding_from") and L:
                hs = [docs[3][11:11 + L] for _ in range(n)]
            lead = [2] + ([77] if cfg.get("forced_bos_token_id") is not None and L else [])
            ids = torch.tensor([lead + h for h in hs], dtype=torch.long)
            for dn, dt in dts.items():
                s = G.make_scores(n, dt, 2000 + 10 * ci + L)
                res = proc(ids, s)
                out["seq"].append({"cfg": cfg, "dtype": dn, "input_ids": ids.tolist(), "salt": 2000 + 10 * ci + L,
                                   "digest": G.row_digest(res), "allowed": G.finite_positions(res)})
    # ---- soft boost: document_constrained_generation_causal.py / _causal_qwen.py ---------------------------------------
    cfgs = [dict(), dict(always_allow_eos=True), dict(always_allow_eos=True, stop_at_count=3),
            dict(always_allow_eos=True, force_decoding_from=force), dict(always_allow_eos=True, forced_bos_token_id=7)]
    for ci, cfg in enumerate(cfgs):
        proc = causal_cls(index=ix, num_beams=2, model_name="llama-3", pad_token_id=G.PAD, eos_token_id=G.EOS, **cfg)
        for L in (0, 1, 5):
            n = 8
            conts = histories(docs, n, L, 700 + ci * 5 + L) if L else [[] for _ in range(n)]
            if cfg.get("force_decoding_from") and L:
                conts = [docs[3][11:11 + L] for _ in range(n)]
            lead = [7] if cfg.get("forced_bos_token_id") is not None and L else []
            r = G.lcg(800 + ci)
            ids = torch.tensor([[r(2000) for _ in range(6)] + G.LLAMA_MARK + lead + c for c in conts], dtype=torch.long)
            for dn, dt in dts.items():
                s = G.make_scores(n, dt, 3000 + 10 * ci + L, specials=False)
                res = proc(ids, s)
                out["causal"].append({"cfg": cfg, "dtype": dn, "input_ids": ids.tolist(), "salt": 3000 + 10 * ci + L,
                                      "digest": G.row_digest(res)})
    cfgs = [dict(min_new_tokens=5), dict(min_new_tokens=5, always_allow_eos=True), dict(min_new_tokens=50, always_allow_eos=True),
            dict(min_new_tokens=3, always_allow_eos=True, stop_at_count=2)]
    for ci, cfg in enumerate(cfgs):
        proc = qwen_cls(index=ix, num_beams=3, pad_token_id=G.PAD, eos_token_id=G.EOS, **cfg)
        width, n = 9, 12
        conts = histories(docs, n, width - len(G.THINK), 900 + ci)
        r = G.lcg(950 + ci)
        rows = []
        for k, c in enumerate(conts):
            if k % 5 == 0:
                body = G.THINK[:2] + [r(2000) for _ in range(width - 2)]       # still thinking: row untouched
            elif k % 5 == 1:
                body = [r(2000) for _ in range(width - 2)] + G.THINK[2:]       # thinking ends at the very end
            else:
                body = G.THINK + c
            rows.append([r(2000) for _ in range(4)] + G.QWEN_MARK + body)
        ids = torch.tensor(rows, dtype=torch.long)
        for dn, dt in dts.items():
            s = G.make_scores(n, dt, 4000 + 10 * ci, specials=False)
            res = proc(ids, s)
            out["qwen"].append({"cfg": cfg, "dtype": dn, "input_ids": ids.tolist(), "salt": 4000 + 10 * ci, "digest": G.row_digest(res)})
        ids2 = torch.tensor([[5, 6] + G.QWEN_MARK + G.THINK for _ in range(3)], dtype=torch.long)  # only the thinking prefix
        s2 = G.make_scores(3, torch.float32, 4500 + ci, specials=False)
        out["qwen"].append({"cfg": cfg, "dtype": "float32", "input_ids": ids2.tolist(), "salt": 4500 + ci, "digest": G.row_digest(proc(ids2, s2))})
    return out


def beam_search_vectors(mods, ix, docs):
    """A 32-step beam search (2 prompts x 4 beams) driven through the reference's seq processor, with beam reordering:
    every new row continues a pseudo-randomly chosen parent row of its prompt with a token taken from that parent's
    allowed set (so beams are duplicated and dropped like in a real search; rows that drew EOS continue with PAD).
    Stored per step: parents, the new tokens and the digest of the reference's output rows."""
    seq_cls = mods["document_constrained_generation_seq"].IndexBasedLogitsProcessor
    force = docs[3][10:11]
    out = []
    for ci, cfg in enumerate([dict(always_allow_eos=True), dict(always_allow_eos=True, stop_at_count=1),
                              dict(always_allow_eos=True, force_decoding_from=force), dict()]):
        nb, nprompt = 4, 2
        R = nb * nprompt
        proc = seq_cls(index=ix, num_beams=nb, pad_token_id=G.PAD, eos_token_id=G.EOS, **cfg)
        r = G.lcg(6000 + ci)
        ids = [[2] for _ in range(R)]
        steps = []
        for step in range(32):
            s = G.make_scores(R, torch.float32, 6000 + 100 * ci + step)
            res = proc(torch.tensor(ids, dtype=torch.long), s)
            allowed = [torch.nonzero(~torch.isneginf(res[i]) & ~torch.isnan(res[i])).flatten().tolist() for i in range(R)]
            parents, toks = [], []
            for row in range(R):
                base = (row // nb) * nb
                p = base + r(nb)
                last = ids[p][-1]
                if last in (G.EOS, G.PAD) and len(ids[p]) > 1:
                    t = G.PAD
                elif allowed[p]:
                    live = [a for a in allowed[p] if a != G.EOS]
                    cand = live if live and r(16) else allowed[p]   # mostly keep the beam alive
                    t = cand[r(len(cand))]
                else:
                    t = G.PAD
                parents.append(p)
                toks.append(t)
            steps.append({"digest": G.row_digest(res), "n_allowed": [len(a) for a in allowed], "parents": parents, "tokens": toks})
            ids = [ids[p] + [t] for p, t in zip(parents, toks)]
        out.append({"cfg": cfg, "num_beams": nb, "rows": R, "steps": steps})
    return out


def generate():
    mods = import_reference()
    cs = corpora()
    res = {"generated_by": "tests/golden/make_reference_golden.py: /root/reference/index.py + document_constrained_generation_"
                           "{vllm,seq,causal,causal_qwen}.py imported unmodified over oracle/ref_shim",
           "indexes": [], "processors": None}
    keep = {}
    for name, docs in cs.items():
        for in_memory in ((True, False) if name in ("worked_example", "random_4000") else (True,)):
            v, ix = index_vectors(mods["index"], name, docs, in_memory)
            res["indexes"].append(v)
            keep[name] = ix
    res["processors"] = processor_vectors(mods, keep["random_4000"], cs["random_4000"])
    res["beam_search"] = beam_search_vectors(mods, keep["random_4000"], cs["random_4000"])
    res["corpora"] = {"unicorn_paragraph": cs["unicorn_paragraph"]}  # token ids of the stand-in tokenizer (no text)
    return res


if __name__ == "__main__":
    r = generate()
    path = os.path.join(HERE, "reference_python.json")
    with open(path, "w") as f:
        json.dump(r, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes;", len(r["indexes"]), "index cases,",
          {k: len(v) for k, v in r["processors"].items()}, "processor cases")
