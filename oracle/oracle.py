"""oracle.py -- ctypes front-end of the C oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

`NativeOracle` mirrors the SWIG-exported class of /root/reference/cpp_modules_linux/fm_index.i:7-21
(method names, tuple returns, (0,0) / 2**64-1 conventions); `OracleFMIndex` mirrors
/root/reference/index.py:20-207 on top of it; `hard_mask_reference` mirrors the hard-mask
arithmetic of /root/reference/document_constrained_generation_vllm.py:46-137.
PARITY UNPINNED against sdsl-lite (see oracle/fm_oracle.c header).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import bisect
import ctypes as C
import os
import subprocess

import numpy as np

SHIFT = 10  # index.py:16
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "_build", "libfm_oracle.so")
    src = os.path.join(_HERE, "fm_oracle.c")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u64, p = C.c_uint64, C.c_void_p
        L.fmo_build.restype = p
        L.fmo_build.argtypes = [p, u64]
        L.fmo_build_from_sa.restype = p
        L.fmo_build_from_sa.argtypes = [p, p, u64]
        L.fmo_suffix_array.restype = C.c_int
        L.fmo_suffix_array.argtypes = [p, u64, p]
        L.fmo_verify_sa.restype = C.c_int
        L.fmo_verify_sa.argtypes = [p, p, u64, C.c_int]
        L.fmo_free.argtypes = [p]
        L.fmo_size.restype = u64
        L.fmo_size.argtypes = [p]
        L.fmo_sigma.restype = u64
        L.fmo_sigma.argtypes = [p]
        L.fmo_backward_search_step.argtypes = [p, u64, u64, u64, p]
        L.fmo_backward_search_multi.argtypes = [p, p, u64, p]
        L.fmo_distinct_count.restype = u64
        L.fmo_distinct_count.argtypes = [p, u64, u64, p, C.c_int]
        L.fmo_locate.restype = u64
        L.fmo_locate.argtypes = [p, u64]
        L.fmo_extract_text.restype = u64
        L.fmo_extract_text.argtypes = [p, u64, u64, p]
        L.fmo_get_range.argtypes = [p, p, u64, C.c_int64, p]
        L.fmo_query_batch.argtypes = [p, p, p, u64, C.c_int64, p, p, p, u64, u64, p, C.c_int]
        L.fmo_locate_batch.argtypes = [p, p, u64, p, C.c_int]
        L.fmo_store_sdsl.restype = C.c_int
        L.fmo_store_sdsl.argtypes = [p, C.c_char_p]
        _LIB = L
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class NativeOracle:
    """SWIG-surface look-alike (fm_index.i:7-21 / fm_index.hpp:20-41)."""

    def __init__(self):
        self._h = None

    def __del__(self):
        if getattr(self, "_h", None):
            lib().fmo_free(self._h)
            self._h = None

    def initialize(self, data):
        a = np.ascontiguousarray(np.asarray(list(data) if not isinstance(data, np.ndarray) else data, dtype=np.uint64))
        h = lib().fmo_build(_ptr(a), a.size)
        if not h:
            raise ValueError("oracle: text contains a zero symbol / symbol too large")
        self._h = h

    def initialize_from_sa(self, text_with_sentinel_u32, sa_u32, verify=True, nthreads=8):
        """Build from a suffix array computed elsewhere (bench.py: the GPU builder's).  verify=True first PROVES, with
        oracle-side code only, that it is the suffix array of the text (fmo_verify_sa), so nothing downstream rests on
        the producer of the array."""
        t = np.ascontiguousarray(text_with_sentinel_u32, dtype=np.uint32)
        s = np.ascontiguousarray(sa_u32, dtype=np.uint32)
        if verify:
            rc = lib().fmo_verify_sa(_ptr(t), _ptr(s), t.size, nthreads)
            if rc != 0:
                raise ValueError("the supplied suffix array is wrong (fmo_verify_sa code %d)" % rc)
        self._h = lib().fmo_build_from_sa(_ptr(t), _ptr(s), t.size)

    def size(self):
        return int(lib().fmo_size(self._h))

    def save(self, path):
        """FMIndex::save (fm_index.cpp:196-199) in sdsl's layout as recalled (see fm_oracle.c)."""
        if lib().fmo_store_sdsl(self._h, str(path).encode()) != 0:
            raise IOError("cannot write %s" % path)

    def sigma(self):
        return int(lib().fmo_sigma(self._h))

    def backward_search_step(self, symbol, low, high):
        out = (C.c_uint64 * 2)()
        lib().fmo_backward_search_step(self._h, symbol, low, high, out)
        return (int(out[0]), int(out[1]))

    def backward_search_multi(self, query):
        q = np.ascontiguousarray(np.asarray(list(query), dtype=np.uint64))
        out = (C.c_uint64 * 2)()
        lib().fmo_backward_search_multi(self._h, _ptr(q), q.size, out)
        return (int(out[0]), int(out[1]))

    def _dc(self, low, high, with_counts):
        buf = np.empty(2 * self.sigma() + 2, dtype=np.uint64)
        w = lib().fmo_distinct_count(self._h, low, high, _ptr(buf), with_counts)
        return tuple(int(x) for x in buf[:w])

    def distinct(self, low, high):
        return self._dc(low, high, 0)

    def distinct_count(self, low, high):
        return self._dc(low, high, 1)

    def distinct_count_multi(self, lows, highs):
        return tuple(self.distinct_count(l, h) for l, h in zip(lows, highs))

    def locate(self, row):
        return int(lib().fmo_locate(self._h, row))

    def extract_text(self, begin, end):
        n = self.size()
        cap = max(0, min(end, n) - begin) + 1
        buf = np.empty(cap, dtype=np.uint64)
        w = lib().fmo_extract_text(self._h, begin, end, _ptr(buf))
        return tuple(int(x) for x in buf[:w])

    # ---- batch drivers (CPU baseline) -----------------------------------------------------
    def query_batch(self, toks, offs, shift=SHIFT, want_mask=False, vocab=128256, nthreads=1):
        toks = np.ascontiguousarray(toks, dtype=np.int64)
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        B = offs.size - 1
        sp = np.empty(B, dtype=np.uint64)
        ep = np.empty(B, dtype=np.uint64)
        words = (vocab + 31) // 32
        mask = np.empty((B, words), dtype=np.uint32) if want_mask else None
        nd = np.empty(B, dtype=np.uint64) if want_mask else None
        lib().fmo_query_batch(self._h, _ptr(toks), _ptr(offs), B, shift, _ptr(sp), _ptr(ep),
                              _ptr(mask) if want_mask else None, words, vocab,
                              _ptr(nd) if want_mask else None, nthreads)
        return sp, ep, mask, nd

    def locate_batch(self, rows, nthreads=1):
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        pos = np.empty(rows.size, dtype=np.uint64)
        lib().fmo_locate_batch(self._h, _ptr(rows), rows.size, _ptr(pos), nthreads)
        return pos


class OracleFMIndex(NativeOracle):
    """index.py:20-207 restated over the oracle (no pickling / files)."""

    def __init__(self):
        super().__init__()
        self.beginnings = [0]
        self.occurring = set()
        self.occurring_distinct = []
        self.occurring_counts = []
        self.labels = None

    def initialize(self, sequences, in_memory=True):
        occurring = set()
        data = []
        for seq in sequences:
            seq = list(seq)
            self.beginnings.append(self.beginnings[-1] + len(seq))
            occurring |= set(seq)
            data.extend(x + SHIFT for x in seq[::-1])
        self.occurring = list(occurring)
        super().initialize(data)
        self.occurring_distinct, self.occurring_counts = self.get_distinct_count(0, len(self))

    def __len__(self):
        return self.beginnings[-1]

    @property
    def n_docs(self):
        return len(self.beginnings) - 1

    def get_range(self, sequence):
        start_row, end_row = 0, self.size()
        for token in sequence:
            start_row, end_row = self.backward_search_step(token + SHIFT, start_row, end_row - 1)
            if start_row == end_row == 0:
                return start_row, end_row
        return start_row, end_row

    def get_count(self, sequence):
        s, e = self.get_range(sequence)
        return e - s

    def get_distinct(self, low, high):
        return [c - SHIFT for c in self.distinct(low, high) if c > 0]

    def get_distinct_count(self, low, high):
        data = self.distinct_count(low, high)
        distinct, counts = [], []
        for d, c in zip(data[0::2], data[1::2]):
            if d > 0:
                distinct.append(d - SHIFT)
                counts.append(c)
        return distinct, counts

    def get_distinct_count_multi(self, lows, highs):
        return [self.get_distinct_count(l, h) for l, h in zip(lows, highs)]

    def get_continuations(self, sequence):
        s, e = self.get_range(sequence)
        return self.get_distinct(s, e)

    def get_doc(self, doc_index):
        doc = self.extract_text(self.beginnings[doc_index], self.beginnings[doc_index + 1])
        return [x - SHIFT for x in doc]

    def get_doc_index(self, token_index):
        return bisect.bisect_right(self.beginnings, token_index) - 1

    def get_doc_length(self, doc_index):
        return self.beginnings[doc_index + 1] - self.beginnings[doc_index]

    def get_token_index_from_row(self, row):
        return self.locate(row)

    def get_doc_index_from_row(self, row):
        return self.get_doc_index(self.locate(row))

    def get_doc_indices(self, sequence):
        s, e = self.get_range(sequence)
        for row in range(s, e):
            yield self.get_doc_index_from_row(row)


def hard_mask_reference(index, input_ids_rows, scores, num_beams, eos_token_id, pad_token_id,
                        stop_at_count=0, always_allow_eos=False, force_decoding_from=None):
    """document_constrained_generation_vllm.py:46-137 (and _seq.py:35-123 without the leading
    decoder-start token), generalised from one row to a [B, V] batch exactly as the _seq
    variant does.  `index` needs get_range / get_count / get_distinct_count_multi /
    occurring_distinct.  Returns scores + mask (torch tensor, same dtype/device as scores)."""
    import torch

    BOOST = 0.0
    mask = torch.full_like(scores, float("-inf"))
    L = len(input_ids_rows[0]) if len(input_ids_rows) else 0
    if L == 0:
        distinct = torch.LongTensor(index.occurring_distinct).to(scores.device)
        mask[:, distinct] = BOOST
    else:
        lows, highs, counts = [], [], []
        for sent in input_ids_rows:
            sent = list(sent)
            if sent[-1] in (eos_token_id, pad_token_id):
                low, high, count = 0, 0, 0
            elif force_decoding_from is not None:
                low, high = index.get_range(force_decoding_from + sent)
                count = index.get_count(force_decoding_from + sent[:-1])
            else:
                low, high = index.get_range(sent)
                count = index.get_count(sent[:-1])
            lows.append(low)
            highs.append(high)
            counts.append(count)
        res = index.get_distinct_count_multi(lows, highs)
        for row, sent in enumerate(input_ids_rows):
            if stop_at_count > 0 and counts[row] <= stop_at_count:
                distinct = [eos_token_id]
            elif sent[-1] == eos_token_id:
                distinct = [pad_token_id]
            elif sent[-1] == pad_token_id:
                distinct = [pad_token_id]
            else:
                distinct, _ = res[row]
            distinct = torch.LongTensor(distinct).to(scores.device)
            mask[row, distinct] = BOOST
    if always_allow_eos:
        mask[:, eos_token_id] = BOOST
    return scores + mask


# ---- soft-boost ("causal") processors -------------------------------------------------------------------------
END_MARKERS = {"qwen": [151645, 198, 151644, 77091, 198], "llama": [78191, 128007, 271]}   # causal.py:40-41
THINK_PREFIX = [151667, 271, 151668, 271]                                                   # causal_qwen.py:60
THINK_DONE = (151668, 271)                                                                  # causal_qwen.py:76


def _strip_end_marker(rows, marker):
    """causal.py:48-60 / causal_qwen.py:45-57: everything after the chat-template end marker; the reference appends one
    result per OCCURRENCE and stacks them, which only lines up with the score rows when each row holds the marker
    exactly once -- that is the supported case here."""
    out = []
    m = len(marker)
    for row in rows:
        hits = [i for i in range(len(row) - m + 1) if list(row[i:i + m]) == list(marker)]
        if len(hits) != 1:
            raise ValueError("every row must contain the end marker exactly once")
        out.append(list(row[hits[0] + m:]))
    return out


def _trailing_ngram(index, sent):
    """causal.py:62-73 as written: the loop returns on its first iteration, so only the LAST token is looked up --
    absent from the corpus -> [], otherwise the whole sequence.  Empty input falls off the loop -> None."""
    for ind in range(len(sent) - 1, -1, -1):
        if index.get_count(sent[ind:]) == 0:
            return sent[ind + 1:]
        return sent
    return None


def soft_boost_reference(index, input_ids_rows, scores, num_beams, variant, eos_token_id, pad_token_id,
                         end_marker, stop_at_count=0, always_allow_eos=False, force_decoding_from=None,
                         forced_bos_token_id=None, length_reward_factor=2.0, min_new_tokens=5, boost=10.0):
    """document_constrained_generation_causal.py:75-180 (variant "causal") and
    document_constrained_generation_causal_qwen.py:100-214 (variant "qwen") restated row by row.
    input_ids_rows: B*beams equal-length token lists INCLUDING the prompt and its end marker."""
    import torch

    rows = _strip_end_marker(input_ids_rows, end_marker)
    mask = torch.full_like(scores, 0.0)
    if variant == "causal" and forced_bos_token_id is not None:          # causal.py:80-85
        if len(rows[0]) == 0:
            mask[:, forced_bos_token_id] = boost
            return scores + mask
        rows = [r[1:] for r in rows]
    width = len(rows[0])
    unig = list(index.occurring_distinct)
    if width == 0:
        if variant == "causal":                                          # causal.py:87-91
            mask[:, torch.LongTensor(unig)] = boost
        # causal_qwen.py:105: nothing happens for an empty continuation (not even the eos rule)
    else:
        if variant == "qwen":                                            # causal_qwen.py:107-112
            def think_done(s):
                return len(s) >= 2 and any((s[i], s[i + 1]) == THINK_DONE for i in range(len(s) - 1))

            def strip_prefix(s):
                k = len(THINK_PREFIX)
                if len(s) < k:
                    return [] if s == THINK_PREFIX[:len(s)] else s
                return s[k:] if s[:k] == THINK_PREFIX else s

            sents = []
            for r in rows:
                if not think_done(r):
                    sents.append(None)
                    continue
                s = strip_prefix(r)
                sents.append([] if s == [] else _trailing_ngram(index, s))
        else:
            sents = [_trailing_ngram(index, r) for r in rows]            # causal.py:100
        lows, highs, counts, which = [], [], [], []
        for i, s in enumerate(sents):
            if s:                                                        # causal.py:104-123
                if s[-1] in (eos_token_id, pad_token_id):
                    lo, hi, cnt = 0, 0, 0
                else:
                    q = (force_decoding_from or []) + s if force_decoding_from is not None else s
                    lo, hi = index.get_range(q)
                    cnt = index.get_count(q)
                lows.append(lo)
                highs.append(hi)
                counts.append(cnt)
                which.append(i)
        res = dict(zip(which, index.get_distinct_count_multi(lows, highs)))
        cnt_of = dict(zip(which, counts))
        for i, s in enumerate(sents):
            if variant == "qwen" and s is None:
                continue
            if not s:                                                    # causal.py:137-140
                mask[i, torch.LongTensor(unig)] = boost
            else:
                if stop_at_count > 0 and cnt_of[i] <= stop_at_count:     # causal.py:144-160
                    distinct = [eos_token_id]
                elif s[-1] == eos_token_id or s[-1] == pad_token_id:
                    distinct = [pad_token_id]
                else:
                    distinct = list(res[i][0])
                dset = set(distinct)
                extra = [u for u in unig if u not in dset]
                mask[i, torch.LongTensor(distinct)] = boost * (length_reward_factor ** len(s))
                if variant == "qwen":                                    # causal_qwen.py:200-205
                    mask[i, torch.LongTensor(extra)] = boost if len(distinct) == 0 else boost / 2
                else:                                                    # causal.py:171
                    mask[i, torch.LongTensor(extra)] = boost / 2
            if variant == "qwen" and always_allow_eos and width >= min_new_tokens:   # causal_qwen.py:207-212
                mask[i, eos_token_id] = boost * (length_reward_factor ** width)
    if variant == "causal" and always_allow_eos:                         # causal.py:173-178
        mask[:, eos_token_id] = boost * (length_reward_factor ** width)
    return scores + mask
