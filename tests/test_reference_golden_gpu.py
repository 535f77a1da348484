"""GPU suite: the PRODUCT against the vectors the reference's own Python produced (tests/golden/reference_python.json,
written by tests/golden/make_reference_golden.py from /root/reference/index.py and the document_constrained_generation_*
processors run unmodified).  Every method of the `index.FMIndex` facade on five corpora (worked example, the unicorn
paragraph of index_for_vllm.py and test_data/sample_corpus.tsv with the stand-in tokenizer, two random corpora), and the
hard-mask / soft-boost logits processors in fp32 and bf16 (stop_at_count, force_decoding_from, forced_bos_token_id,
finished beams, num_beams > 1), bit for bit.  When /root/reference is present the reference's index.py itself is also
run on top of this repo's `cpp_modules.fm_index` (the zero-change integration route of INTEGRATION.md)."""
import importlib
import json
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
import refgolden_common as G  # noqa: E402
from test_reference_python import GOLD, DT, check_index_vectors, corpus_docs, hard_mask_rows, replay_beam_search  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("layout", ["resident", "sampled"])
@pytest.mark.parametrize("case", range(len(GOLD["indexes"])))
def test_index_facade_matches_reference_python(cuda, case, layout):
    from fm_index_constrained_decoding_b200 import FMIndex
    v = GOLD["indexes"][case]
    ix = FMIndex(layout=layout)
    ix.initialize(corpus_docs(v["name"]), in_memory=v["in_memory"])
    check_index_vectors(ix, v)


def test_save_load_matches_reference_python(cuda, tmp_path):
    from fm_index_constrained_decoding_b200 import FMIndex
    v = next(x for x in GOLD["indexes"] if x["name"] == "random_4000")
    ix = FMIndex()
    ix.initialize(corpus_docs("random_4000"), in_memory=True)
    ix.save(str(tmp_path / "idx"))
    check_index_vectors(FMIndex.load(str(tmp_path / "idx")), v)


@pytest.fixture(scope="module")
def index4000():
    from fm_index_constrained_decoding_b200 import FMIndex
    ix = FMIndex()
    ix.initialize(G.random_corpus(11), in_memory=True)
    return ix


@pytest.mark.parametrize("style", ["vllm", "seq"])
def test_hard_mask_processor_matches_reference_python(cuda, index4000, style):
    from fm_index_constrained_decoding_b200.processors import IndexBasedLogitsProcessor
    for v in GOLD["processors"][style]:
        cfg = dict(v["cfg"])
        nb = cfg.pop("num_beams", 1)
        proc = IndexBasedLogitsProcessor(index4000, num_beams=nb, pad_token_id=G.PAD, eos_token_id=G.EOS,
                                         skip_first_token=(style == "seq"), **cfg)
        if style == "vllm":
            hs, scores = hard_mask_rows(v, style)
            got = torch.stack([proc(tuple(h), scores[i].cuda()) for i, h in enumerate(hs)])  # one request per call
        else:
            ids = torch.tensor(v["input_ids"], dtype=torch.long)
            got = proc(ids.cuda(), G.make_scores(ids.shape[0], DT[v["dtype"]], v["salt"]).cuda())
        assert got.dtype == DT[v["dtype"]]
        assert G.finite_positions(got) == v["allowed"], (style, v["cfg"], v["dtype"])
        assert G.row_digest(got) == v["digest"], (style, v["cfg"], v["dtype"])


@pytest.mark.parametrize("variant", ["causal", "qwen"])
def test_soft_boost_processor_matches_reference_python(cuda, index4000, variant):
    from fm_index_constrained_decoding_b200.processors import CausalIndexBasedLogitsProcessor, QwenIndexBasedLogitsProcessor
    for v in GOLD["processors"][variant]:
        cfg = dict(v["cfg"])
        if variant == "causal":
            proc = CausalIndexBasedLogitsProcessor(index4000, 2, "llama-3", pad_token_id=G.PAD, eos_token_id=G.EOS, **cfg)
        else:
            proc = QwenIndexBasedLogitsProcessor(index4000, 3, pad_token_id=G.PAD, eos_token_id=G.EOS, **cfg)
        ids = torch.tensor(v["input_ids"], dtype=torch.long)
        scores = G.make_scores(ids.shape[0], DT[v["dtype"]], v["salt"], specials=False)
        got = proc(ids.cuda(), scores.cuda())
        assert got.dtype == DT[v["dtype"]] and G.row_digest(got) == v["digest"], (variant, v["cfg"], v["dtype"])


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree only exists in the build container")
def test_reference_index_py_unmodified_over_the_gpu_shim(cuda):
    """The reference's own index.py, imported as it is, with `cpp_modules.fm_index` resolving to this repo's ctypes
    module (cpp_modules/): the drop-in route of INTEGRATION.md section 1."""
    for m in ("cpp_modules", "cpp_modules.fm_index", "index"):
        sys.modules.pop(m, None)
    sys.path.insert(0, "/root/reference")
    sys.path.insert(0, ROOT)
    try:
        ref_index = importlib.import_module("index")
        assert ref_index.__file__.startswith("/root/reference")
        assert sys.modules["cpp_modules.fm_index"].__file__.startswith(os.path.join(ROOT, "cpp_modules"))
        for v in GOLD["indexes"]:
            ix = ref_index.FMIndex()
            ix.initialize([list(d) for d in corpus_docs(v["name"])], in_memory=v["in_memory"])
            check_index_vectors(ix, v)
    finally:
        sys.path.remove("/root/reference")
        for m in ("cpp_modules", "cpp_modules.fm_index", "index"):
            sys.modules.pop(m, None)


@pytest.mark.parametrize("mode", ["beam_idx", "matched_parents", "searched"])
def test_incremental_decode_replays_the_reference_beam_search(cuda, index4000, mode):
    """32 steps x 8 rows with beam reordering, device-resident state: one backward-search step per beam per token
    (`fmidx_decode_step`), parents given (beam_idx), matched on the device, or no state at all (every step searched) --
    every step's rows equal the reference seq processor's, bit for bit."""
    from fm_index_constrained_decoding_b200.processors import IndexBasedLogitsProcessor
    for case in GOLD["beam_search"]:
        cfg = dict(case["cfg"])
        proc = IndexBasedLogitsProcessor(index4000, num_beams=case["num_beams"], pad_token_id=G.PAD, eos_token_id=G.EOS,
                                         skip_first_token=True, incremental=(mode != "searched"), **cfg)
        prev_parents = None
        for step, ids, scores, st in replay_beam_search(case):
            bi = torch.tensor(prev_parents, dtype=torch.int32).cuda() if (mode == "beam_idx" and prev_parents is not None) else None
            got = proc(torch.tensor(ids, dtype=torch.long).cuda(), scores.cuda(), beam_idx=bi)
            assert G.row_digest(got) == st["digest"], (mode, case["cfg"], step)
            prev_parents = st["parents"]
        if mode == "searched":
            assert proc.steps_incremental == 0
        else:
            assert proc.steps_incremental == 30 and proc.steps_searched == 1  # step 0 has no history, step 1 searches, 2..31 step
