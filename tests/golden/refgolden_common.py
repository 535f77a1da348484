"""Shared by tests/golden/make_reference_golden.py (runs the REFERENCE's Python, in the build container) and the tests
that check the oracle / the product against the vectors it wrote.  Everything here is deterministic integer arithmetic,
so the same inputs can be rebuilt anywhere without storing them."""
import hashlib

import torch

V, EOS, PAD = 128256, 128009, 128004
LLAMA_MARK = [78191, 128007, 271]                    # document_constrained_generation_causal.py:40-41
QWEN_MARK = [151645, 198, 151644, 77091, 198]
THINK = [151667, 271, 151668, 271]                   # document_constrained_generation_causal_qwen.py:60


def lcg(seed):
    """Tiny deterministic generator (no dependence on any library's RNG stream)."""
    state = [(seed * 6364136223846793005 + 1442695040888963407) & (2 ** 64 - 1)]

    def nxt(n):
        state[0] = (state[0] * 6364136223846793005 + 1442695040888963407) & (2 ** 64 - 1)
        return (state[0] >> 33) % n
    return nxt


def fnv1a(word):
    """Stand-in tokenizer (the Llama-3.2 tokenizer files are not available offline): FNV-1a word hash mod 128000."""
    h = 0xCBF29CE484222325
    for b in word.encode("utf-8"):
        h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h % 128000


def random_corpus(seed, n_docs=80, vocab=4000, lo=12, hi=90, repeats=20):
    r = lcg(seed)
    docs = [[r(vocab) for _ in range(lo + r(hi - lo))] + [EOS] for _ in range(n_docs)]
    for k in range(repeats):  # repeated passages: prefixes with several continuations
        docs[n_docs - repeats + k] = docs[k][3:3 + 20] + docs[n_docs - repeats + k]
    return docs


def make_scores(B, dtype, salt, vocab=V, specials=True):
    """[B, vocab] scores from integer arithmetic (exact in fp32; rounded once to `dtype`), with +inf, -inf, NaN and -0.0
    planted in row 0 so that the `x + 0.0` / `x + (-inf)` / `x + boost` corner cases are part of every vector."""
    v = torch.arange(vocab, dtype=torch.int64).unsqueeze(0)
    b = torch.arange(B, dtype=torch.int64).unsqueeze(1)
    x = (((v * 2654435761 + b * 40503 + salt * 977) % 4001) - 2000).to(torch.float32) / 128.0
    if specials:
        x[0, 0], x[0, 1], x[0, 2], x[0, 3] = float("inf"), float("-inf"), float("nan"), -0.0
    return x.to(dtype)


def row_digest(t):
    """sha256 of every row's bytes with NaNs canonicalised (NaN payloads differ between x86 and the GPU; -0.0 / +0.0 and
    the infinities are kept as they are)."""
    t = torch.nan_to_num(t.detach().cpu().float(), nan=12345.0, posinf=float("inf"), neginf=float("-inf")).contiguous()
    return [hashlib.sha256(t[i].numpy().tobytes()).hexdigest()[:24] for i in range(t.shape[0])]


def finite_positions(t, cap=48):
    """Per row: the positions that are not -inf (the allowed set), capped; for diagnostics next to the digest."""
    t = t.detach().cpu().float()
    out = []
    for i in range(t.shape[0]):
        idx = torch.nonzero(~torch.isneginf(t[i])).flatten().tolist()
        out.append(idx if len(idx) <= cap else {"n": len(idx), "first": idx[:cap]})
    return out
