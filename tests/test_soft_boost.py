"""CPU suite: the soft-boost processors' host logic and tensor arithmetic against the restated reference
(oracle.soft_boost_reference = document_constrained_generation_causal.py / _causal_qwen.py, row by row).
The index behind the processor is a stand-in with the facade's batched methods answered by the CPU oracle, so the
whole step runs without a GPU; tests/test_processors.py::test_soft_boost_on_gpu runs it over the real index."""
import random

import numpy as np
import pytest
import torch

V, EOS, PAD, ABSENT = 5000, 4990, 4991, 2500
LLAMA_MARK = [78191, 128007, 271]
QWEN_MARK = [151645, 198, 151644, 77091, 198]
THINK = [151667, 271, 151668, 271]


class OracleBackedIndex:
    """get_range_batch / distinct_mask_batch / occurring_distinct of the FMIndex facade on CPU tensors."""

    def __init__(self, o):
        self.o = o
        self.occurring_distinct = list(o.occurring_distinct)

    def __len__(self):
        return len(self.o)

    def get_range_batch(self, toks, offs):
        t, f = toks.tolist(), offs.tolist()
        r = [self.o.get_range(t[f[i]:f[i + 1]]) for i in range(len(f) - 1)]
        return torch.tensor([x[0] for x in r], dtype=torch.int64), torch.tensor([x[1] for x in r], dtype=torch.int64)

    def distinct_mask_batch(self, sp, ep, vocab):
        words = (vocab + 31) // 32
        m = np.zeros((sp.numel(), words), np.uint32)
        for i, (lo, hi) in enumerate(zip(sp.tolist(), ep.tolist())):
            for t in self.o.get_distinct_count(lo, hi)[0]:
                if 0 <= t < vocab:
                    m[i, t >> 5] |= np.uint32(1 << (t & 31))
        return torch.from_numpy(m.view(np.int32))


def corpus(seed=3):
    from oracle.oracle import OracleFMIndex
    rng = random.Random(seed)
    docs = [[rng.randrange(2000) for _ in range(rng.randint(15, 80))] + [EOS] for _ in range(60)]
    for k in range(20):
        docs[30 + k] = docs[k][3:25] + docs[30 + k]
    o = OracleFMIndex()
    o.initialize(docs)
    return rng, docs, o


def continuations(rng, docs, width, n):
    rows = []
    for k in range(n):
        d = rng.choice(docs)
        i = rng.randrange(0, len(d) - width)
        c = d[i:i + width]
        if k % 7 == 1:
            c = c[:-1] + [EOS]
        elif k % 7 == 2:
            c = c[:-1] + [PAD]
        elif k % 7 == 3:
            c = c[:-1] + [ABSENT]                      # last token not in the corpus: no trailing n-gram
        elif k % 7 == 4:
            c = [rng.randrange(2000) for _ in c[:-1]] + [c[-1]]  # last token known, n-gram as a whole (most likely) not
        rows.append(c)
    return rows


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("stop_at_count,allow_eos,force,bos", [(0, False, False, None), (0, True, False, None),
                                                               (3, True, False, None), (0, True, True, None), (0, True, False, 7)])
def test_causal_variant(dtype, stop_at_count, allow_eos, force, bos):
    from fm_index_constrained_decoding_b200.processors import CausalIndexBasedLogitsProcessor
    from oracle.oracle import soft_boost_reference
    rng, docs, o = corpus()
    ix = OracleBackedIndex(o)
    fdf = docs[4][8:9] if force else None
    for width in (0, 1, 5):
        n = 16
        if force and width:
            conts = [docs[4][9:9 + width] for _ in range(n)]
        else:
            conts = continuations(rng, docs, width, n) if width else [[] for _ in range(n)]
        lead = [bos] if bos is not None and width else []
        rows = [[rng.randrange(2000) for _ in range(6)] + LLAMA_MARK + lead + c for c in conts]
        scores = torch.randn(n, V, generator=torch.Generator().manual_seed(width)).to(dtype)
        proc = CausalIndexBasedLogitsProcessor(ix, 4, "llama-3", pad_token_id=PAD, eos_token_id=EOS, force_decoding_from=fdf,
                                               stop_at_count=stop_at_count, always_allow_eos=allow_eos, forced_bos_token_id=bos)
        got = proc(torch.tensor(rows), scores)
        exp = soft_boost_reference(o, rows, scores, 4, "causal", EOS, PAD, LLAMA_MARK, stop_at_count, allow_eos, fdf, bos)
        assert got.dtype == dtype and torch.equal(got, exp), (width, stop_at_count, allow_eos, force, bos)


@pytest.mark.parametrize("allow_eos,min_new", [(False, 5), (True, 5), (True, 50)])
def test_qwen_variant(allow_eos, min_new):
    from fm_index_constrained_decoding_b200.processors import QwenIndexBasedLogitsProcessor
    from oracle.oracle import soft_boost_reference
    rng, docs, o = corpus(5)
    ix = OracleBackedIndex(o)
    width, n = 9, 21
    conts = continuations(rng, docs, width - len(THINK), n)
    rows = []
    for k, c in enumerate(conts):
        if k % 5 == 0:
            body = THINK[:2] + [rng.randrange(2000) for _ in range(width - 2)]     # still thinking -> row untouched
        elif k % 5 == 1:
            body = [rng.randrange(2000) for _ in range(width - 2)] + THINK[2:]     # thinking ends at the very end
        else:
            body = THINK + c
        rows.append([rng.randrange(2000) for _ in range(4)] + QWEN_MARK + body)
    scores = torch.randn(n, V, generator=torch.Generator().manual_seed(1))
    proc = QwenIndexBasedLogitsProcessor(ix, 3, min_new_tokens=min_new, pad_token_id=PAD, eos_token_id=EOS,
                                         stop_at_count=2, always_allow_eos=allow_eos)
    got = proc(torch.tensor(rows), scores)
    exp = soft_boost_reference(o, rows, scores, 3, "qwen", EOS, PAD, QWEN_MARK, 2, allow_eos, None, None, 2.0, min_new)
    assert torch.equal(got, exp)
    # a row that is only the thinking prefix: candidate [] -> every corpus unigram boosted
    rows2 = [[5, 6] + QWEN_MARK + THINK for _ in range(3)]
    s2 = torch.randn(3, V, generator=torch.Generator().manual_seed(2))
    assert torch.equal(proc(torch.tensor(rows2), s2),
                       soft_boost_reference(o, rows2, s2, 3, "qwen", EOS, PAD, QWEN_MARK, 2, allow_eos, None, None, 2.0, min_new))


def test_end_marker_must_be_unique():
    from fm_index_constrained_decoding_b200.processors import CausalIndexBasedLogitsProcessor
    _, _, o = corpus()
    proc = CausalIndexBasedLogitsProcessor(OracleBackedIndex(o), 1, "llama")
    with pytest.raises(ValueError):
        proc(torch.tensor([[1, 2, 3, 4, 5, 6]]), torch.zeros(1, V))


def test_random_histories_property():
    """Randomised comparison (hypothesis): any mix of corpus n-grams, unknown tokens, EOS / PAD endings, thinking
    prefixes and parameters gives exactly the reference's scores."""
    from hypothesis import given, settings, strategies as st
    from fm_index_constrained_decoding_b200.processors import SoftBoostLogitsProcessor
    from oracle.oracle import soft_boost_reference
    rng0, docs, o = corpus(8)
    ix = OracleBackedIndex(o)
    tok = st.one_of(st.integers(0, 1999), st.sampled_from([EOS, PAD, ABSENT, 271, 151668, 151667]))

    @st.composite
    def case(draw):
        variant = draw(st.sampled_from(["causal", "qwen"]))
        width = draw(st.integers(0, 8))
        n = draw(st.integers(1, 5))
        rows = []
        for _ in range(n):
            kind = draw(st.integers(0, 3))
            if kind == 0 and width:
                d = docs[draw(st.integers(0, len(docs) - 1))]
                i = draw(st.integers(0, max(0, len(d) - width)))
                c = (d[i:i + width] + [EOS] * width)[:width]
            else:
                c = [draw(tok) for _ in range(width)]
            if variant == "qwen" and width >= 4 and draw(st.booleans()):
                c = (THINK + c)[:width]
            rows.append(c)
        return (variant, rows, draw(st.integers(0, 3)), draw(st.booleans()), draw(st.integers(0, 6)),
                draw(st.sampled_from([1.5, 2.0])), draw(st.integers(0, 2 ** 31)))

    @settings(max_examples=60, deadline=None)
    @given(case())
    def run(c):
        variant, conts, stop, allow_eos, min_new, factor, seed = c
        marker = QWEN_MARK if variant == "qwen" else LLAMA_MARK
        rows = [[3, 4] + marker + r for r in conts]
        scores = torch.randn(len(rows), V, generator=torch.Generator().manual_seed(seed))
        proc = SoftBoostLogitsProcessor(ix, 1, variant=variant, model_name="llama", min_new_tokens=min_new, pad_token_id=PAD,
                                        eos_token_id=EOS, stop_at_count=stop, always_allow_eos=allow_eos,
                                        length_reward_factor=factor)
        exp = soft_boost_reference(o, rows, scores, 1, variant, EOS, PAD, marker, stop, allow_eos, None, None, factor, min_new)
        assert torch.equal(proc(torch.tensor(rows), scores), exp)

    run()
