"""`cpp_modules.fm_index` as the reference's index.py:13-14 imports it, backed by the C oracle (oracle/fm_oracle.c)
instead of SWIG + sdsl-lite.  TEST INFRASTRUCTURE: lets /root/reference/index.py run unmodified on the CPU so that its
outputs can be committed as golden vectors (tests/golden/make_reference_golden.py).  Surface: fm_index.i:7-21."""
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from oracle.oracle import NativeOracle  # noqa: E402


def _u64(x):
    """SWIG `unsigned long` in-typemap."""
    if isinstance(x, bool) or not isinstance(x, (int, np.integer)):
        raise TypeError("in method argument of type 'unsigned long'")
    if x < 0 or x > 2 ** 64 - 1:
        raise OverflowError("in method argument of type 'unsigned long'")
    return int(x)


class IntVector(list):
    pass


class IntVectorVector(list):
    pass


class FMIndex(NativeOracle):
    """fm_index.hpp:20-44.  Vectors come back as tuples (SWIG std_vector.i out-typemap)."""

    def initialize(self, data):
        self._text = [_u64(x) for x in data]
        NativeOracle.initialize(self, np.asarray(self._text, dtype=np.uint64))

    def initialize_from_file(self, file, width):  # fm_index.cpp:43-48: `width`-byte little-endian integers
        raw = np.fromfile(file, dtype={1: "<u1", 2: "<u2", 4: "<u4", 8: "<u8"}[int(width)])
        FMIndex.initialize(self, [int(x) for x in raw])  # not the subclass's (index.py:39)

    def backward_search_step(self, symbol, low, high):
        return NativeOracle.backward_search_step(self, _u64(symbol), _u64(low), _u64(high))

    def distinct(self, low, high):
        return NativeOracle.distinct(self, _u64(low), _u64(high))

    def distinct_count(self, low, high):
        return NativeOracle.distinct_count(self, _u64(low), _u64(high))

    def distinct_count_multi(self, lows, highs):
        return tuple(self.distinct_count(l, h) for l, h in zip(lows, highs))

    def locate(self, row):
        return NativeOracle.locate(self, _u64(row))

    def extract_text(self, begin, end):
        return NativeOracle.extract_text(self, _u64(begin), _u64(end))

    def save(self, path):  # fm_index.cpp:196-199; the shim's own container: the text
        np.save(path + ".npy", np.asarray(self._text, dtype=np.uint64))
        os.replace(path + ".npy", path)


def load_FMIndex(path):  # fm_index.cpp:201-209
    obj = FMIndex()
    with open(path, "rb") as f:
        FMIndex.initialize(obj, [int(x) for x in np.load(f)])
    return obj
