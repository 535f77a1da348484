"""CPU suite: the C-ABI library loads, exports every symbol include/fmidx.h declares, and fails
loudly (no fallback) without a CUDA device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fmidx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fmidx_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported():
    from fm_index_constrained_decoding_b200 import _native
    L = _native.lib()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), "libfmidx.so does not export %s" % s
    assert sorted(_native.SIGNATURES) == syms, "ctypes table and header disagree"
    assert L.fmidx_abi_version() == 1


def test_swig_surface_names():
    import cpp_modules.fm_index as m
    for name in ["FMIndex", "load_FMIndex", "IntVector", "IntVectorVector"]:
        assert hasattr(m, name)
    for meth in ["initialize", "initialize_from_file", "backward_search_multi", "backward_search_step", "distinct",
                 "distinct_count", "distinct_count_multi", "size", "locate", "extract_text", "save"]:
        assert callable(getattr(m.FMIndex, meth))
    v = m.IntVector([1, 2])
    v.push_back(3)
    assert list(v) == [1, 2, 3] and v.size() == 3
    with pytest.raises(OverflowError):
        v.push_back(-1)


def test_facade_surface_names():
    from fm_index_constrained_decoding_b200 import FMIndex
    for meth in ["initialize", "get_doc", "get_doc_index", "get_doc_length", "get_token_index_from_row",
                 "get_doc_index_from_row", "get_range", "get_count", "get_doc_indices", "get_continuations",
                 "get_distinct", "get_distinct_count", "get_distinct_count_multi", "save", "load", "__len__"]:
        assert hasattr(FMIndex, meth)
    assert isinstance(FMIndex.n_docs, property)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fm_index_constrained_decoding_b200 import FMIndex, _native
    with pytest.raises(_native.FmidxError) as e:
        FMIndex()
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fm_index_constrained_decoding_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), "%s mentions the oracle" % f


def test_arg_conversion_like_swig():
    from fm_index_constrained_decoding_b200.fm_index import _u64
    assert _u64(5) == 5 and _u64((1 << 64) - 1) == (1 << 64) - 1
    with pytest.raises(OverflowError):
        _u64(-1)
    with pytest.raises(OverflowError):
        _u64(1 << 64)
    with pytest.raises(TypeError):
        _u64(1.5)
    with pytest.raises(TypeError):
        _u64("3")
