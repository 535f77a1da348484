"""fm_index_constrained_decoding_b200 -- B200-native FM-index query path for document-constrained
decoding (the hot path of oksanatkach/fm-index-constrained-decoding, rebuilt for sm_100a).

    from fm_index_constrained_decoding_b200 import FMIndex      # index.py-compatible facade
    from fm_index_constrained_decoding_b200 import fm_index     # SWIG-module look-alike

Everything runs on the GPU through libfmidx.so (include/fmidx.h); there is no CPU fallback.
"""
from .index import FMIndex, SHIFT, ragged  # noqa: F401
from .fm_index import load_FMIndex, IntVector, IntVectorVector  # noqa: F401
from . import fm_index  # noqa: F401

__all__ = ["FMIndex", "SHIFT", "ragged", "load_FMIndex", "IntVector", "IntVectorVector", "fm_index"]
