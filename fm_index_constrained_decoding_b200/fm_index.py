"""Drop-in replacement for the SWIG module `cpp_modules.fm_index` of the reference
(generated from cpp_modules_linux/fm_index.i:7-21 over cpp_modules_linux/fm_index.hpp:20-44).

Same Python-visible names and conventions -- `FMIndex`, `load_FMIndex`, `IntVector`,
`IntVectorVector`; vectors come back as tuples of ints; empty search = (0, 0); locate out of
range = 2**64-1 -- but every call runs on the GPU through libfmidx.so (include/fmidx.h).
The reference's own `index.py` works unmodified on top of this module (see cpp_modules/).
"""
import ctypes as C

import numpy as np

from . import _native as N

U64_MAX = (1 << 64) - 1


def _u64(x, name="value"):
    """SWIG `unsigned long` in-typemap: ints only, no negatives, no overflow."""
    if isinstance(x, bool) or not isinstance(x, (int, np.integer)):
        raise TypeError("in method argument '%s' of type 'unsigned long'" % name)
    x = int(x)
    if x < 0 or x > U64_MAX:
        raise OverflowError("in method argument '%s' of type 'unsigned long'" % name)
    return x


def _u64_array(seq, name="vector"):
    if isinstance(seq, np.ndarray) and seq.dtype == np.uint64:
        return np.ascontiguousarray(seq)
    return np.ascontiguousarray(np.fromiter((_u64(x, name) for x in seq), dtype=np.uint64))


class IntVector(list):
    """std::vector<unsigned long> proxy (fm_index.i:13)."""

    def push_back(self, x):
        self.append(_u64(x))

    def size(self):
        return len(self)


class IntVectorVector(list):
    """std::vector<std::vector<unsigned long>> proxy (fm_index.i:14)."""

    def push_back(self, x):
        self.append(IntVector(x))

    def size(self):
        return len(self)


class FMIndex:
    """class FMIndex of fm_index.hpp:20-44, device-resident."""

    # Build layouts.  "resident" (default) spends HBM to remove dependent loads: full SA, full ISA and
    # the text stay on the device (17.4 B/token instead of 5.45), so locate is one gather, extract is a
    # gather and backward search stops probing as soon as a range is a single row.  "sampled" keeps the
    # reference's sdsl defaults (csa_wt_int<>: SA sample every 32, ISA sample every 64, no text).
    # Query results are identical in every layout.
    LAYOUTS = {"resident": (1, 1, True), "sampled": (32, 64, False)}
    default_layout = None  # None -> $FMIDX_LAYOUT or "resident"

    def __init__(self, device=None, sa_dens=None, isa_dens=None, keep_text=None, layout=None):
        import os
        self._h = None
        L = N.lib()
        h = C.c_void_p()
        N.check(L.fmidx_new(N.default_device() if device is None else int(device), C.byref(h)))
        self._h = h
        layout = layout or self.default_layout or os.environ.get("FMIDX_LAYOUT", "resident")
        d_sa, d_isa, d_text = self.LAYOUTS[layout]
        N.check(L.fmidx_set_options(
            self._h,
            d_sa if sa_dens is None else int(sa_dens),
            d_isa if isa_dens is None else int(isa_dens),
            int(d_text if keep_text is None else keep_text)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                N.lib().fmidx_free(h)
            except Exception:
                pass
            self._h = None

    # ---- build / persist ------------------------------------------------------------------------
    def initialize(self, data):
        """fm_index.cpp:33-41."""
        a = _u64_array(data, "data")
        N.check(N.lib().fmidx_initialize(self._h, a.ctypes.data_as(C.c_void_p), a.size))

    def initialize_from_file(self, file, width):
        """fm_index.cpp:43-48."""
        N.check(N.lib().fmidx_initialize_from_file(self._h, str(file).encode(), int(width)))

    def save(self, path):
        """fm_index.cpp:196-199."""
        N.check(N.lib().fmidx_save(self._h, str(path).encode()))

    # ---- queries ----------------------------------------------------------------------------------
    def size(self):
        """fm_index.cpp:50-52."""
        out = C.c_uint64()
        N.check(N.lib().fmidx_size(self._h, C.byref(out)))
        return int(out.value)

    def backward_search_multi(self, query):
        """fm_index.cpp:55-65."""
        q = _u64_array(query, "query")
        out = (C.c_uint64 * 2)()
        N.check(N.lib().fmidx_backward_search_multi(self._h, q.ctypes.data_as(C.c_void_p), q.size, out))
        return (int(out[0]), int(out[1]))

    def backward_search_step(self, symbol, low, high):
        """fm_index.cpp:67-97."""
        out = (C.c_uint64 * 2)()
        N.check(N.lib().fmidx_backward_search_step(self._h, _u64(symbol, "symbol"), _u64(low, "low"), _u64(high, "high"), out))
        return (int(out[0]), int(out[1]))

    def _distinct(self, fn, low, high, per):
        low, high = _u64(low, "low"), _u64(high, "high")
        # one reusable output buffer per object (grown on demand: the C call reports the size it needs), instead of a
        # 2 * (max_sym + 1) u64 allocation and an fmidx_info call per scalar query
        buf = getattr(self, "_dbuf", None)
        if buf is None:
            buf = self._dbuf = (C.c_uint64 * 8192)()  # a ctypes array: no numpy / .ctypes round trip per call
        k = C.c_size_t()
        rc = fn(self._h, low, high, buf, len(buf), C.byref(k))
        if rc == N.FMIDX_ERANGE and k.value > len(buf):
            buf = self._dbuf = (C.c_uint64 * (int(k.value) + 1024))()
            rc = fn(self._h, low, high, buf, len(buf), C.byref(k))
        if rc != 0:
            N.check(rc)
        return tuple(buf[:k.value])

    def distinct(self, low, high):
        """fm_index.cpp:99-117."""
        return self._distinct(N.lib().fmidx_distinct, low, high, 1)

    def distinct_count(self, low, high):
        """fm_index.cpp:119-144: interleaved (sym0, cnt0, sym1, cnt1, ...)."""
        return self._distinct(N.lib().fmidx_distinct_count, low, high, 2)

    def distinct_count_multi(self, lows, highs):
        """fm_index.cpp:146-166: one device batch instead of one std::async per range."""
        lo, hi = _u64_array(lows, "lows"), _u64_array(highs, "highs")
        nr = min(lo.size, hi.size) if lo.size != hi.size else lo.size
        offs = np.zeros(nr + 1, dtype=np.uint64)
        L = N.lib()
        cap = 1 << 16
        while True:
            buf = np.empty(cap, dtype=np.uint64)
            rc = L.fmidx_distinct_count_multi(self._h, lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p), nr,
                                              buf.ctypes.data_as(C.c_void_p), cap, offs.ctypes.data_as(C.c_void_p))
            if rc == N.FMIDX_ERANGE and int(offs[nr]) > cap:
                cap = int(offs[nr])
                continue
            N.check(rc)
            break
        return tuple(tuple(int(x) for x in buf[int(offs[r]):int(offs[r + 1])]) for r in range(nr))

    def locate(self, row):
        """fm_index.cpp:168-172."""
        out = C.c_uint64()
        N.check(N.lib().fmidx_locate(self._h, _u64(row, "row"), C.byref(out)))
        return int(out.value)

    def extract_text(self, begin, end):
        """fm_index.cpp:174-194 (as written: walks back from isa[end-1])."""
        begin, end = _u64(begin, "begin"), _u64(end, "end")
        n = self.size()
        cap = max(0, min(end, n) - begin)
        buf = np.empty(max(cap, 1), dtype=np.uint64)
        k = C.c_size_t()
        N.check(N.lib().fmidx_extract_text(self._h, begin, end, buf.ctypes.data_as(C.c_void_p), cap, C.byref(k)))
        return tuple(int(x) for x in buf[:k.value])

    # ---- extras (not in the SWIG surface) ------------------------------------------------------------
    def tune(self, **kw):
        """Run-time tunables (fmidx_tune, see include/fmidx.h): scan_max, bytes_occ, unigram_table, scalar_fast,
        scalar_speculate, lazy_clear (2 = also the caller's mask buffers: nothing else may write them between the calls of
        their slot).  Results never depend on them."""
        for k, v in kw.items():
            N.check(N.lib().fmidx_tune(self._h, k.encode(), int(v)))

    def info(self):
        out = (C.c_uint64 * 8)()
        N.check(N.lib().fmidx_info(self._h, out))
        keys = ["size", "levels", "max_sym", "sa_dens", "isa_dens", "device_bytes", "cell_bytes_per_level", "keep_text"]
        return {k: int(v) for k, v in zip(keys, out)}


def load_FMIndex(path, device=None):
    """fm_index.cpp:201-209.  Returns an FMIndex whose __class__ may be re-assigned by the
    caller exactly like the reference does (index.py:203).  Accepts both this library's own
    container and an sdsl `csa_wt_int<>` dump written by the reference's FMIndex::save (converted
    on the fly, see sdsl_format.py)."""
    from . import sdsl_format
    dev = N.default_device() if device is None else int(device)
    if not sdsl_format.is_own_container(path):
        import torch
        tdev = torch.device("cuda", dev)
        text, _ = sdsl_format.text_from_sdsl_file(path, tdev)
        sym = text.to(torch.int32).contiguous()
        obj = FMIndex(device=dev)
        with torch.cuda.device(tdev):
            N.check(N.lib().fmidx_initialize_device(obj._h, C.c_void_p(sym.data_ptr()), sym.numel(),
                                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            torch.cuda.current_stream().synchronize()
        return obj
    h = C.c_void_p()
    N.check(N.lib().fmidx_load(str(path).encode(), dev, C.byref(h)))
    obj = FMIndex.__new__(FMIndex)
    obj._h = h
    return obj
