"""CPU suite: host-side helpers (ragged packing, synthetic corpus generator)."""
import numpy as np
import torch

from fm_index_constrained_decoding_b200.index import ragged
from fm_index_constrained_decoding_b200 import synth


def test_ragged():
    toks, offs = ragged([[1, 2, 3], [], [7]])
    assert toks.dtype == np.int32 and offs.dtype == np.int64
    assert toks.tolist() == [1, 2, 3, 7] and offs.tolist() == [0, 3, 3, 4]
    toks, offs = ragged([])
    assert toks.size == 0 and offs.tolist() == [0]


def test_zipf_corpus_shape_and_determinism():
    t1, b1 = synth.zipf_corpus(50_000, seed=0)
    t2, b2 = synth.zipf_corpus(50_000, seed=0)
    assert torch.equal(t1, t2) and torch.equal(b1, b2)
    assert t1.dtype == torch.int32 and t1.numel() == 50_000
    assert int(t1.min()) >= 0 and int(t1.max()) < synth.VOCAB
    assert b1[0] == 0 and b1[-1] == 50_000 and (b1[1:] > b1[:-1]).all()
    assert (t1[b1[1:] - 1] == synth.EOS).all()
    # Zipf: the most frequent token is much more frequent than the median one
    counts = torch.bincount(t1.long(), minlength=synth.VOCAB)
    assert counts.max() > 50 * counts[counts > 0].float().median()


def test_sample_queries():
    t, beg = synth.zipf_corpus(20_000, seed=2)
    q, offs = synth.sample_queries(t, 100, seed=1, beginnings=beg)
    lens = offs[1:] - offs[:-1]
    assert lens.min() >= 1 and lens.max() <= 32 and offs[-1] == q.numel()
    # positive queries are substrings of the stream
    s = t.tolist()
    for k in range(5):
        pat = q[offs[k]:offs[k + 1]].tolist()
        assert any(s[i:i + len(pat)] == pat for i in range(len(s) - len(pat) + 1))
    # and never cross a document boundary (EOS can only be the last token)
    for k in range(100):
        assert synth.EOS not in q[offs[k]:offs[k + 1] - 1].tolist()
    qn, _ = synth.sample_queries(t, 100, seed=1, negative_frac=0.5, beginnings=beg)
    assert (qn != q).any()
