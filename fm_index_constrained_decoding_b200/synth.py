"""Synthetic token corpora and query batches of the shapes named in BASELINE.json (there is no
network for real corpora or tokenizers).  Pure torch ops, runs on CPU or CUDA; deterministic for
a given (seed, device type).

Corpus: token ids Zipf(s) over the vocabulary (rank -> id through a seeded permutation),
documents of length ~U[doc_min, doc_max] each ending in EOS, plus a fraction of planted repeated
passages so that long prefixes keep more than one match.
"""
import torch

VOCAB = 128256      # Llama-3.2 vocabulary
EOS = 128009        # <|eot_id|>


def zipf_corpus(n_tokens, vocab=VOCAB, seed=0, device="cpu", eos=EOS, doc_min=64, doc_max=1024, zipf_s=1.1,
                planted_frac=0.01, passage_min=16, passage_max=256):
    """Returns (tokens int32 [n_tokens], beginnings int64 [n_docs+1]) on `device`."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    # Zipf over ranks, inverse-CDF sampling
    ranks = torch.arange(1, vocab, device=dev, dtype=torch.float64)  # vocab-1 non-EOS ids
    cdf = torch.cumsum(ranks.pow(-zipf_s), 0)
    cdf = (cdf / cdf[-1]).to(torch.float32)
    perm = torch.randperm(vocab, generator=g, device=dev)
    perm = perm[perm != eos]  # rank r -> token id perm[r]; EOS only at document ends
    toks = torch.empty(n_tokens, dtype=torch.int32, device=dev)
    CH = 1 << 26
    for a in range(0, n_tokens, CH):
        b = min(n_tokens, a + CH)
        u = torch.rand(b - a, generator=g, device=dev)
        r = torch.searchsorted(cdf, u).clamp_(max=vocab - 2)
        toks[a:b] = perm[r].to(torch.int32)
    # planted repeats: copy passages from the first half into random later places
    n_plant = int(n_tokens * planted_frac / ((passage_min + passage_max) / 2))
    if n_plant > 0 and n_tokens > 4 * passage_max:
        plen = torch.randint(passage_min, passage_max + 1, (n_plant,), generator=g, device=dev)
        src = torch.randint(0, n_tokens // 2 - passage_max, (n_plant,), generator=g, device=dev)
        dst = torch.randint(n_tokens // 2, n_tokens - passage_max, (n_plant,), generator=g, device=dev)
        off = torch.arange(int(plen.sum()), device=dev) - torch.repeat_interleave(torch.cumsum(plen, 0) - plen, plen)
        si = torch.repeat_interleave(src, plen) + off
        di = torch.repeat_interleave(dst, plen) + off
        toks[di] = toks[si]
    # documents
    mean = (doc_min + doc_max) / 2
    n_guess = int(n_tokens / mean * 1.2) + 16
    lens = torch.randint(doc_min, doc_max + 1, (n_guess,), generator=g, device=dev)
    ends = torch.cumsum(lens, 0)
    ends = ends[ends < n_tokens]
    beginnings = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), ends.to(torch.int64),
                            torch.tensor([n_tokens], dtype=torch.int64, device=dev)])
    if beginnings[-1] == beginnings[-2]:
        beginnings = beginnings[:-1]
    toks[beginnings[1:] - 1] = eos
    return toks, beginnings


def sample_queries(tokens, n_queries, len_min=1, len_max=32, seed=1, negative_frac=0.0, vocab=VOCAB, beginnings=None):
    """Ragged query batch: each prefix is a substring of one document (positive; with
    `beginnings` the length is clipped at the document end, because the index concatenates the
    documents REVERSED, so a forward substring crossing a boundary need not occur); a fraction
    has one position replaced by a random token (mostly negative -> early exit).
    Returns (toks int32 [total], offs int64 [n_queries+1]) on tokens.device."""
    dev = tokens.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    n = tokens.numel()
    lens = torch.randint(len_min, len_max + 1, (n_queries,), generator=g, device=dev)
    starts = torch.randint(0, max(1, n - len_max), (n_queries,), generator=g, device=dev)
    if beginnings is not None:
        beg = beginnings.to(dev)
        doc_end = beg[torch.searchsorted(beg, starts, right=True)]
        lens = torch.minimum(lens, doc_end - starts)
    offs = torch.zeros(n_queries + 1, dtype=torch.int64, device=dev)
    offs[1:] = torch.cumsum(lens, 0)
    total = int(offs[-1])
    rel = torch.arange(total, device=dev) - torch.repeat_interleave(offs[:-1], lens)
    q = tokens[torch.repeat_interleave(starts, lens) + rel].clone().to(torch.int32)
    if negative_frac > 0:
        bad = torch.rand(n_queries, generator=g, device=dev) < negative_frac
        where = (torch.rand(n_queries, generator=g, device=dev) * lens).long().clamp_(max=len_max - 1)
        where = torch.minimum(where, lens - 1)
        idx = (offs[:-1] + where)[bad]
        q[idx] = torch.randint(0, vocab, (idx.numel(),), generator=g, device=dev, dtype=torch.int32)
    return q, offs
