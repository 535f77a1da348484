"""GPU suite: the hard-mask IndexBasedLogitsProcessor over the fused kernel against the restated
reference arithmetic (oracle.hard_mask_reference = document_constrained_generation_vllm.py:46-137)."""
import random

import pytest
import torch

pytestmark = pytest.mark.gpu
V, EOS, PAD = 128256, 128009, 128004


def setup(seed=21):
    from fm_index_constrained_decoding_b200 import FMIndex
    from oracle.oracle import OracleFMIndex
    rng = random.Random(seed)
    docs = [[rng.randrange(4000) for _ in range(rng.randint(20, 150))] + [EOS] for _ in range(120)]
    for k in range(30):
        docs[60 + k] = docs[k][5:50] + docs[60 + k]
    ix = FMIndex()
    ix.initialize(docs, in_memory=True)
    o = OracleFMIndex()
    o.initialize(docs)
    return rng, docs, ix, o


def same(a, b):
    return torch.equal(torch.nan_to_num(a.float().cpu(), nan=0.0), torch.nan_to_num(b.float().cpu(), nan=0.0))


@pytest.mark.parametrize("stop_at_count,allow_eos,force", [(0, False, None), (0, True, None), (2, True, None), (0, True, [17])])
def test_batch_style_matches_reference(cuda, stop_at_count, allow_eos, force):
    from fm_index_constrained_decoding_b200.processors import IndexBasedLogitsProcessor
    from oracle.oracle import hard_mask_reference
    rng, docs, ix, o = setup()
    if force is not None:
        force = docs[3][10:11]
    Lh, B = 6, 24
    rows = []
    for k in range(B):
        d = rng.choice(docs)
        i = rng.randrange(1 if force is None else 11, len(d) - Lh)
        h = d[i:i + Lh] if force is None else docs[3][11:11 + Lh]
        if k % 6 == 1:
            h = h[:-1] + [EOS]
        if k % 6 == 2:
            h = h[:-1] + [PAD]
        if k % 6 == 3:
            h = h[:-1] + [rng.randrange(4000)]
        rows.append(h)
    g = torch.Generator().manual_seed(5)
    scores = torch.randn(B, V, generator=g)
    proc = IndexBasedLogitsProcessor(ix, num_beams=4, pad_token_id=PAD, eos_token_id=EOS, stop_at_count=stop_at_count,
                                     always_allow_eos=allow_eos, force_decoding_from=force)
    got = proc(torch.tensor(rows), scores.cuda())
    exp = hard_mask_reference(o, rows, scores, 4, EOS, PAD, stop_at_count, allow_eos, force)
    assert same(got, exp)
    # seq2seq style: decoder-start token in column 0
    proc2 = proc.clone()
    proc2.skip_first_token = True
    got2 = proc2(torch.tensor([[2] + r for r in rows]), scores.cuda())
    assert same(got2, exp)


def test_vllm_per_request_style_and_empty_history(cuda):
    from fm_index_constrained_decoding_b200.processors import IndexBasedLogitsProcessor
    from oracle.oracle import hard_mask_reference
    rng, docs, ix, o = setup(4)
    proc = IndexBasedLogitsProcessor(ix, num_beams=1, pad_token_id=PAD, eos_token_id=EOS, always_allow_eos=True)
    g = torch.Generator().manual_seed(6)
    for k in range(6):
        d = rng.choice(docs)
        h = tuple(d[3:3 + k])  # k == 0: empty history -> occurring_distinct
        s = torch.randn(V, generator=g)
        got = proc(h, s.cuda())
        exp = hard_mask_reference(o, [list(h)], s.unsqueeze(0), 1, EOS, PAD, 0, True)[0]
        assert got.shape == (V,) and same(got, exp)


def test_forced_bos(cuda):
    from fm_index_constrained_decoding_b200.processors import IndexBasedLogitsProcessor
    rng, docs, ix, o = setup(8)
    proc = IndexBasedLogitsProcessor(ix, num_beams=1, pad_token_id=PAD, eos_token_id=EOS, always_allow_eos=True,
                                     forced_bos_token_id=77)
    s = torch.randn(2, V)
    got = proc.process_batch([[], []], s.cuda()).cpu()
    exp = torch.full_like(s, float("-inf"))
    exp[:, 77] = s[:, 77] + 0.0
    assert torch.equal(got, exp)
    # non-empty: first token dropped, then the usual step
    from oracle.oracle import hard_mask_reference
    h = [[5] + docs[0][2:6], [5] + docs[1][4:8]]
    got = proc.process_batch(h, s.cuda())
    exp = hard_mask_reference(o, [x[1:] for x in h], s, 1, EOS, PAD, 0, True)
    assert same(got, exp)


@pytest.mark.parametrize("variant", ["causal", "qwen"])
def test_soft_boost_on_gpu(cuda, variant):
    """Soft-boost processors (causal.py / causal_qwen.py semantics) over the real index: search + enumeration on
    the GPU, boosts added in the dtype of the scores; bit-exact against the row-by-row restatement."""
    from fm_index_constrained_decoding_b200.processors import CausalIndexBasedLogitsProcessor, QwenIndexBasedLogitsProcessor
    from oracle.oracle import soft_boost_reference
    rng, docs, ix, o = setup(33)
    think = [151667, 271, 151668, 271]
    marker = [151645, 198, 151644, 77091, 198] if variant == "qwen" else [78191, 128007, 271]
    width, n = 6, 24
    rows = []
    for k in range(n):
        d = rng.choice(docs)
        i = rng.randrange(0, len(d) - width)
        c = d[i:i + width]
        if k % 6 == 1:
            c = c[:-1] + [EOS]
        if k % 6 == 2:
            c = c[:-1] + [PAD]
        if k % 6 == 3:
            c = c[:-1] + [4500]  # not in the corpus
        if variant == "qwen":
            c = (think[:2] + c + [5, 6]) if k % 6 == 4 else (think + c)
        rows.append([rng.randrange(4000) for _ in range(5)] + marker + c)
    g = torch.Generator().manual_seed(9)
    for dtype in (torch.float32, torch.bfloat16):
        scores = torch.randn(n, V, generator=g).to(dtype)
        if variant == "causal":
            proc = CausalIndexBasedLogitsProcessor(ix, 4, "llama", pad_token_id=PAD, eos_token_id=EOS, stop_at_count=2, always_allow_eos=True)
        else:
            proc = QwenIndexBasedLogitsProcessor(ix, 4, min_new_tokens=3, pad_token_id=PAD, eos_token_id=EOS, stop_at_count=2, always_allow_eos=True)
        got = proc(torch.tensor(rows), scores.cuda())
        exp = soft_boost_reference(o, rows, scores, 4, variant, EOS, PAD, marker, 2, True, None, None, 2.0, 3)
        assert got.dtype == dtype and torch.equal(got.cpu(), exp)
