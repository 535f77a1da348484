"""sdsl `csa_wt_int<>` dump reader (next-row f1).  The files come from the oracle's exporter, which
restates sdsl-lite's serialisation layout from memory -- the reader and the writer are therefore only
checked against EACH OTHER and against the text; neither is verified against a real sdsl-lite file
(none exists in this environment)."""
import random

import numpy as np
import pytest
import torch

from fm_index_constrained_decoding_b200 import sdsl_format
from oracle.oracle import OracleFMIndex


def corpus(seed, V=3000, ndocs=40):
    rng = random.Random(seed)
    docs = [[rng.randrange(V) for _ in range(rng.randint(3, 120))] for _ in range(ndocs)]
    docs[5] = docs[2][1:30] + docs[5]
    return docs


@pytest.mark.parametrize("seed,V", [(1, 4), (2, 300), (3, 128256)])
def test_parse_decode_invert_roundtrip(tmp_path, seed, V):
    docs = corpus(seed, V)
    o = OracleFMIndex()
    o.initialize(docs)
    path = str(tmp_path / "ref.fmi")
    o.save(path)
    assert not sdsl_format.is_own_container(path)
    parts = sdsl_format.read_csa_wt_int(path)
    n = o.size()
    assert parts["size"] == n and parts["sigma"] == o.sigma()
    # samples as sdsl keeps them: SA every 32nd row, ISA every 64th text position
    assert parts["sa_samples"].tolist() == [o.locate(r) for r in range(0, n, 32)]
    text, _ = sdsl_format.text_from_sdsl_file(path)
    expect = [t + 10 for d in docs for t in d[::-1]]
    assert text.tolist() == expect
    bwt = sdsl_format.decode_bwt(parts["tree"], n, parts["max_level"])
    T = expect + [0]
    sa = [o.locate(r) for r in range(n)]
    assert bwt.tolist() == [T[p - 1] for p in sa]


def test_reader_rejects_garbage(tmp_path):
    p = tmp_path / "x.fmi"
    p.write_bytes(np.random.default_rng(0).integers(0, 255, 5000, dtype=np.uint8).tobytes())
    with pytest.raises(sdsl_format.SdslFormatError):
        sdsl_format.read_csa_wt_int(str(p))
    o = OracleFMIndex()
    o.initialize(corpus(9))
    q = tmp_path / "y.fmi"
    o.save(str(q))
    data = q.read_bytes()
    q.write_bytes(data + b"\x00" * 8)  # trailing bytes
    with pytest.raises(sdsl_format.SdslFormatError):
        sdsl_format.read_csa_wt_int(str(q))
    q.write_bytes(data[:-9])  # truncated
    with pytest.raises(sdsl_format.SdslFormatError):
        sdsl_format.read_csa_wt_int(str(q))


def test_unpack_int_vector():
    vals = np.array([5, 0, 1023, 77, 512, 1], dtype=np.uint64)
    width = 10
    words = np.zeros(2, dtype=np.uint64)
    for i, v in enumerate(vals):
        bit = i * width
        words[bit >> 6] |= np.uint64(int(v) << (bit & 63) & 0xFFFFFFFFFFFFFFFF)
        if (bit & 63) + width > 64:
            words[(bit >> 6) + 1] |= np.uint64(int(v) >> (64 - (bit & 63)))
    assert sdsl_format.unpack_int_vector(words, len(vals), width).tolist() == vals.tolist()


@pytest.mark.gpu
def test_load_reference_built_index_on_gpu(cuda, tmp_path):
    """index.FMIndex.load on a `.fmi` in the reference's (recalled) sdsl layout + its `.oth` pickle."""
    import pickle
    from fm_index_constrained_decoding_b200 import FMIndex
    docs = corpus(4, 5000, 80)
    o = OracleFMIndex()
    o.initialize(docs)
    base = str(tmp_path / "wiki.fm_index")
    o.save(base + ".fmi")
    with open(base + ".oth", "wb") as f:
        pickle.dump((o.beginnings, o.occurring, None), f)  # index.py:193-194
    ix = FMIndex.load(base)
    assert ix.size() == o.size() and ix.beginnings == o.beginnings
    assert (ix.occurring_distinct, ix.occurring_counts) == (o.occurring_distinct, o.occurring_counts)
    rng = random.Random(0)
    for _ in range(60):
        d = rng.choice(docs)
        i = rng.randrange(len(d))
        q = d[i:i + rng.randint(1, 6)]
        r = ix.get_range(q)
        assert r == o.get_range(q)
        assert ix.get_distinct_count(*r) == o.get_distinct_count(*r)
    assert [ix.locate(r) for r in range(0, o.size(), 53)] == [o.locate(r) for r in range(0, o.size(), 53)]
    assert [ix.get_doc(i) for i in range(0, 80, 9)] == [o.get_doc(i) for i in range(0, 80, 9)]
