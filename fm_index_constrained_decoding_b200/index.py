"""Python facade with the surface of the reference's `index.py` (class FMIndex, index.py:20-207),
re-implemented over the device-resident index, plus the batched tensor entry points the GPU
path exists for.

Reference-compatible methods (same names, argument meaning, return shapes and quirks):
    initialize, get_doc, get_doc_index, get_doc_length, get_token_index_from_row,
    get_doc_index_from_row, get_range, get_count, get_doc_indices, get_continuations,
    get_distinct, get_distinct_count, get_distinct_count_multi, __len__, n_docs, save, load
Batched additions (torch CUDA tensors in, torch CUDA tensors out, current stream):
    get_range_batch, step_batch, distinct_mask_batch, mask_logits, locate_batch, extract_batch,
    doc_index_batch
"""
import bisect
import ctypes as C
import pickle
import tempfile

import numpy as np

from . import _native as N
from .fm_index import FMIndex as _FMIndex
from .fm_index import load_FMIndex

SHIFT = 10      # index.py:16 -- symbol = token + SHIFT (0 is the sentinel)
FORMAT = "<i4"  # index.py:18 ('<l'): little-endian int32 symbols in the temporary text file


def _torch():
    import torch
    return torch


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(torch):
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ragged(sequences):
    """list of token lists -> (int32 tokens, int64 offsets) numpy arrays."""
    lens = np.fromiter((len(s) for s in sequences), dtype=np.int64, count=len(sequences))
    offs = np.zeros(len(sequences) + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    toks = np.empty(int(offs[-1]), dtype=np.int32)
    for i, s in enumerate(sequences):
        toks[offs[i]:offs[i + 1]] = s
    return toks, offs


class FMIndex(_FMIndex):
    def __init__(self, device=None, sa_dens=None, isa_dens=None, keep_text=None, layout=None):
        super().__init__(device=device, sa_dens=sa_dens, isa_dens=isa_dens, keep_text=keep_text, layout=layout)
        self._devidx = N.default_device() if device is None else int(device)
        self._init_side_data()

    def _init_side_data(self):
        self.beginnings = [0]
        self.occurring = set()
        self.occurring_distinct = []
        self.occurring_counts = []
        self.labels = None
        self._beg_dev = None

    # ---- build (index.py:39-67) ----------------------------------------------------------------------
    def initialize(self, sequences, in_memory=False):
        occurring = set()
        chunks = []
        for seq in sequences:
            seq = list(seq)
            self.beginnings.append(self.beginnings[-1] + len(seq))
            occurring.update(seq)
            chunks.append(np.asarray(seq[::-1], dtype=np.int64) + SHIFT)  # reversed, shifted
        self.occurring = list(occurring)
        flat = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.int64)
        if flat.size and flat.min() < 0:
            raise OverflowError("negative symbol (token + SHIFT < 0)")
        if in_memory:
            _FMIndex.initialize(self, flat.astype(np.uint64))
        else:
            # same hand-off as the reference: a temporary file of little-endian int32 symbols
            with tempfile.NamedTemporaryFile() as tmp:
                flat.astype(FORMAT).tofile(tmp)
                tmp.flush()
                _FMIndex.initialize_from_file(self, tmp.name, 4)
        self._beg_dev = None
        self.occurring_distinct, self.occurring_counts = self.get_distinct_count(0, len(self))

    def initialize_from_device_tokens(self, tokens, beginnings):
        """Build from a CUDA int32/int64 tensor of token ids that is ALREADY the concatenation of
        the documents (forward order); `beginnings` are the cumulative doc offsets (len m+1).
        Used for the synthetic 100M..1B-token corpora, which never exist on the host."""
        torch = _torch()
        beg = torch.as_tensor(beginnings, dtype=torch.int64, device=tokens.device)
        n = int(beg[-1])
        # position p of the reversed-doc concatenation holds token[begin_d + end_d - 1 - p]
        pos = torch.arange(n, device=tokens.device, dtype=torch.int64)
        doc = torch.searchsorted(beg, pos, right=True) - 1
        src = beg[doc] + beg[doc + 1] - 1 - pos
        sym = (tokens.to(torch.int64)[src] + SHIFT).to(torch.int32).contiguous()
        del pos, doc, src
        N.check(N.lib().fmidx_initialize_device(self._h, _ptr(sym), n, _stream(torch)))
        torch.cuda.current_stream().synchronize()
        self.beginnings = [int(x) for x in beg.tolist()]
        self._beg_dev = None
        self.occurring_distinct, self.occurring_counts = self.get_distinct_count(0, len(self))
        self.occurring = list(self.occurring_distinct)

    # ---- documents (index.py:69-101) -------------------------------------------------------------------
    def get_doc(self, doc_index):
        raw = self.extract_text(self.beginnings[doc_index], self.beginnings[doc_index + 1])
        return [s - SHIFT for s in raw]

    def get_doc_index(self, token_index):
        return bisect.bisect_right(self.beginnings, token_index) - 1

    def get_doc_length(self, doc_index):
        return self.beginnings[doc_index + 1] - self.beginnings[doc_index]

    def get_token_index_from_row(self, row):
        return self.locate(row)

    def get_doc_index_from_row(self, row):
        return self.get_doc_index(self.locate(row))

    # ---- search (index.py:103-137) -----------------------------------------------------------------------
    def get_range(self, sequence):
        """One kernel launch for the whole prefix instead of one SWIG call per token.  Prefixes of up to 96 tokens go
        through preallocated ctypes scratch (the numpy round trip of the general path costs more than the GPU call)."""
        n = len(sequence)
        if n <= 96:
            sc = getattr(self, "_rq", None)
            if sc is None:
                sc = self._rq = ((C.c_int32 * 96)(), (C.c_int64 * 2)(), C.c_uint64(), C.c_uint64(), N.lib().fmidx_search_host)
            toks, offs, sp, ep, fn = sc
            try:
                if n and (max(sequence) > 0x7FFFFFFF or min(sequence) < -0x80000000):
                    raise OverflowError  # ctypes would wrap silently
                toks[:n] = sequence if isinstance(sequence, list) else list(sequence)
            except (TypeError, ValueError, OverflowError):
                n = -1  # not int32 material: the general path raises what it always raised
            if n >= 0:
                offs[1] = n
                if fn(self._h, toks, offs, 1, SHIFT, C.byref(sp), C.byref(ep)) != 0:
                    N.check(-1)
                return sp.value, ep.value
        toks, offs = ragged([list(sequence)])
        sp = np.empty(1, dtype=np.uint64)
        ep = np.empty(1, dtype=np.uint64)
        N.check(N.lib().fmidx_search_host(self._h, toks.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p), 1,
                                          SHIFT, sp.ctypes.data_as(C.c_void_p), ep.ctypes.data_as(C.c_void_p)))
        return int(sp[0]), int(ep[0])

    def get_range_many(self, sequences):
        """get_range for a list of prefixes in ONE device batch (host lists in, list of (sp, ep) out)."""
        toks, offs = ragged([list(s) for s in sequences])
        B = len(sequences)
        sp = np.empty(max(B, 1), dtype=np.uint64)
        ep = np.empty(max(B, 1), dtype=np.uint64)
        if B:
            N.check(N.lib().fmidx_search_host(self._h, toks.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p), B,
                                              SHIFT, sp.ctypes.data_as(C.c_void_p), ep.ctypes.data_as(C.c_void_p)))
        return [(int(sp[i]), int(ep[i])) for i in range(B)]

    def doc_indices_of_rows(self, lo, hi):
        """get_doc_index_from_row for rows [lo, hi) in one locate batch."""
        if hi <= lo:
            return []
        rows = np.arange(lo, hi, dtype=np.uint64)
        pos = np.empty(rows.size, dtype=np.uint64)
        N.check(N.lib().fmidx_locate_host(self._h, rows.ctypes.data_as(C.c_void_p), rows.size, pos.ctypes.data_as(C.c_void_p)))
        beg = np.asarray(self.beginnings, dtype=np.uint64)
        return [int(d) for d in np.searchsorted(beg, pos, side="right") - 1]

    def get_count(self, sequence):
        lo, hi = self.get_range(sequence)
        return hi - lo

    def get_doc_indices(self, sequence):
        lo, hi = self.get_range(sequence)
        if hi <= lo:
            return
        rows = np.arange(lo, hi, dtype=np.uint64)
        pos = np.empty(rows.size, dtype=np.uint64)
        N.check(N.lib().fmidx_locate_host(self._h, rows.ctypes.data_as(C.c_void_p), rows.size, pos.ctypes.data_as(C.c_void_p)))
        beg = np.asarray(self.beginnings, dtype=np.uint64)
        for d in np.searchsorted(beg, pos, side="right") - 1:
            yield int(d)

    def get_continuations(self, sequence):
        lo, hi = self.get_range(sequence)
        return self.get_distinct(lo, hi)

    # ---- distinct next tokens (index.py:139-174) -----------------------------------------------------------
    def get_distinct(self, low, high):
        return [s - SHIFT for s in self.distinct(low, high) if s > 0]

    @staticmethod
    def _split(data):
        toks, cnts = [], []
        for s, c in zip(data[0::2], data[1::2]):
            if s > 0:  # drop the sentinel, index.py:156
                toks.append(s - SHIFT)
                cnts.append(c)
        return toks, cnts

    def get_distinct_count(self, low, high):
        return self._split(self.distinct_count(low, high))

    def get_distinct_count_multi(self, lows, highs):
        return [self._split(d) for d in self.distinct_count_multi(lows, highs)]

    # ---- sizes / persistence (index.py:176-207) ----------------------------------------------------------------
    def __len__(self):
        return self.beginnings[-1]

    @property
    def n_docs(self):
        return len(self.beginnings) - 1

    def save(self, path):
        with open(path + ".oth", "wb") as f:
            pickle.dump((self.beginnings, self.occurring, self.labels), f)
        return _FMIndex.save(self, path + ".fmi")

    @classmethod
    def load(cls, path, device=None):
        index = load_FMIndex(path + ".fmi", device=device)
        index.__class__ = cls
        index._devidx = N.default_device() if device is None else int(device)
        index._init_side_data()
        with open(path + ".oth", "rb") as f:
            index.beginnings, index.occurring, index.labels = pickle.load(f)
        index.occurring_distinct, index.occurring_counts = index.get_distinct_count(0, len(index))
        return index

    # =============================================================================================
    # batched tensor API (CUDA tensors on this index's device; work is enqueued on torch's current
    # stream; nothing here synchronises)
    # =============================================================================================
    def _dev(self, torch):
        return torch.device("cuda", self.device_index)

    @property
    def device_index(self):
        if not hasattr(self, "_devidx"):
            self._devidx = N.default_device()
        return self._devidx

    def to_device_ragged(self, sequences):
        torch = _torch()
        toks, offs = ragged(sequences)
        dev = self._dev(torch)
        return torch.from_numpy(toks).to(dev), torch.from_numpy(offs).to(dev)

    def get_range_batch(self, toks, offs, shift=SHIFT):
        """get_range for B ragged prefixes.  toks int32 [total], offs int64 [B+1] (CUDA).
        Returns (sp, ep) int64 [B]."""
        torch = _torch()
        B = offs.numel() - 1
        out = torch.empty((2, max(B, 1)), dtype=torch.int64, device=toks.device)
        assert toks.dtype == torch.int32 and offs.dtype == torch.int64 and toks.is_cuda and offs.is_cuda
        N.check(N.lib().fmidx_search_batch(self._h, _ptr(toks), _ptr(offs), B, shift, _ptr(out[0]), _ptr(out[1]), _stream(torch)))
        return out[0, :B], out[1, :B]

    def step_batch(self, symbols, lows, highs):
        """B independent backward_search_step calls (inclusive highs). int64 CUDA tensors."""
        torch = _torch()
        B = symbols.numel()
        out = torch.empty((2, max(B, 1)), dtype=torch.int64, device=symbols.device)
        N.check(N.lib().fmidx_step_batch(self._h, _ptr(symbols), _ptr(lows), _ptr(highs), B, _ptr(out[0]), _ptr(out[1]), _stream(torch)))
        return out[0, :B], out[1, :B]

    def distinct_mask_batch(self, sp, ep, vocab, shift=SHIFT, want_counts=False, out=None):
        """Allowed-token bitmaps [B, ceil(vocab/32)] int32 (+ optional counts [B, vocab] int32,
        defined at set bits only).  out: optional preallocated contiguous [B, words] int32."""
        torch = _torch()
        B = sp.numel()
        words = (vocab + 31) // 32
        mask = torch.empty((B, words), dtype=torch.int32, device=sp.device) if out is None else out
        assert mask.is_contiguous() and mask.shape == (B, words) and mask.dtype == torch.int32
        counts = torch.empty((B, vocab), dtype=torch.int32, device=sp.device) if want_counts else None
        if B:
            N.check(N.lib().fmidx_distinct_mask_batch(self._h, _ptr(sp), _ptr(ep), B, shift, vocab, _ptr(mask), _ptr(counts), _stream(torch)))
        return (mask, counts) if want_counts else mask

    def query_batch(self, toks, offs, vocab, shift=SHIFT, ws_slot=0, out=None):
        """Ranges AND allowed-token bitmaps of B ragged prefixes in one call (fmidx_query_batch): returns
        (sp, ep, mask).  Batches issued on different CUDA streams must use different ws_slot values (0..31)."""
        torch = _torch()
        B = offs.numel() - 1
        words = (vocab + 31) // 32
        if out is None:
            out = (torch.empty(B, dtype=torch.int64, device=toks.device), torch.empty(B, dtype=torch.int64, device=toks.device),
                   torch.empty((B, words), dtype=torch.int32, device=toks.device))
        sp, ep, mask = out
        assert mask.is_contiguous() and mask.shape == (B, words) and mask.dtype == torch.int32
        if B:
            N.check(N.lib().fmidx_query_batch(self._h, ws_slot, _ptr(toks), _ptr(offs), B, shift, vocab, _ptr(sp), _ptr(ep),
                                              _ptr(mask), _stream(torch)))
        return sp, ep, mask

    def mask_logits(self, sp, ep, logits, out=None, override_tok=None, always_allow=-1, extra_masks=None,
                    mask_out=None, shift=SHIFT):
        """Fused constrained-decoding step: out = logits + (-inf outside the allowed set).
        logits [B, V] float32/bfloat16/float16, row-contiguous. out=None -> in place."""
        torch = _torch()
        B, V = logits.shape
        assert logits.stride(1) == 1 and logits.is_cuda
        if out is None:
            out = logits
        assert out.shape == logits.shape and out.stride() == logits.stride() and out.dtype == logits.dtype
        code = N.DTYPE_CODES[str(logits.dtype).replace("torch.", "")]
        n_extra = 0 if extra_masks is None else int(extra_masks.shape[0])
        N.check(N.lib().fmidx_mask_logits(self._h, _ptr(sp), _ptr(ep), B, shift, V, _ptr(logits), _ptr(out), code,
                                          logits.stride(0), _ptr(override_tok), int(always_allow), _ptr(extra_masks),
                                          n_extra, _ptr(mask_out), _stream(torch)))
        return out

    def locate_batch(self, rows):
        torch = _torch()
        pos = torch.empty_like(rows)
        N.check(N.lib().fmidx_locate_batch(self._h, _ptr(rows), rows.numel(), _ptr(pos), _stream(torch)))
        return pos

    def extract_batch(self, begins, ends, max_len):
        torch = _torch()
        n = begins.numel()
        out = torch.zeros((n, max_len), dtype=torch.int64, device=begins.device)
        lens = torch.empty(n, dtype=torch.int64, device=begins.device)
        N.check(N.lib().fmidx_extract_batch(self._h, _ptr(begins), _ptr(ends), n, max_len, _ptr(out), _ptr(lens), _stream(torch)))
        return out, lens

    def doc_index_batch(self, positions):
        """get_doc_index (index.py:78-83) for a CUDA tensor of text positions."""
        torch = _torch()
        if self._beg_dev is None or self._beg_dev.device != positions.device:
            self._beg_dev = torch.tensor(self.beginnings, dtype=torch.int64, device=positions.device)
        return torch.searchsorted(self._beg_dev, positions, right=True) - 1


class GraphQuerySlots:
    """The host-buffer query pipeline as ONE CUDA-graph launch per batch (fmidx_query_graph_*): `n_slots` slots, each with
    pinned staging buffers that the caller fills / reads in place:
        slots.offs[i] int64 [B+1], slots.toks[i] int32 [max_tokens]      (inputs; offs[0] = 0)
        slots.sp[i], slots.ep[i] uint64 [B], slots.n_distinct[i] uint32 [B]   (results, valid after wait(i))
    launch(i) enqueues H2D -> search -> enumeration -> popcount -> D2H on the slot's stream and returns."""

    def __init__(self, index, n_slots, B, max_tokens, vocab, shift=SHIFT, first_slot=0):
        """first_slot: the handle's slots first_slot .. first_slot + n_slots - 1 are taken (32 per handle; a slot is either
        captured or used by fmidx_query_submit / as the ws_slot of fmidx_query_batch, never both)."""
        self.index, self.n, self.B, self.first = index, n_slots, B, first_slot
        self.offs, self.toks, self.sp, self.ep, self.n_distinct = [], [], [], [], []
        L = N.lib()
        for i in range(n_slots):
            hin, hout = C.c_void_p(), C.c_void_p()
            N.check(L.fmidx_query_graph_prepare(index._h, first_slot + i, B, max_tokens, shift, vocab, C.byref(hin), C.byref(hout)))
            self.offs.append(np.ctypeslib.as_array((C.c_int64 * (B + 1)).from_address(hin.value)))
            self.toks.append(np.ctypeslib.as_array((C.c_int32 * max_tokens).from_address(hin.value + (B + 1) * 8)))
            self.sp.append(np.ctypeslib.as_array((C.c_uint64 * B).from_address(hout.value)))
            self.ep.append(np.ctypeslib.as_array((C.c_uint64 * B).from_address(hout.value + B * 8)))
            self.n_distinct.append(np.ctypeslib.as_array((C.c_uint32 * B).from_address(hout.value + B * 16)))
        self._launch, self._wait = L.fmidx_query_graph_launch, L.fmidx_query_wait

    def launch(self, i):
        if self._launch(self.index._h, self.first + i) != 0:
            N.check(-1)

    def wait(self, i):
        if self._wait(self.index._h, self.first + i) != 0:
            N.check(-1)


class DeviceGraphSteps:
    """The device-resident query step (search + enumeration into a mask buffer) captured once per slot for FIXED device
    buffers and replayed with one CUDA-graph launch (fmidx_query_graph_prepare_device / fmidx_query_graph_launch_on): the
    serving-loop form of FMIndex.query_batch.  Slot i owns toks[i] (int32 [max_tokens]), offs[i] (int64 [B+1], offs[0] = 0),
    sp[i], ep[i] (int64 [B]) and mask[i] (int32 [B, words]); the caller rewrites toks / offs in place between replays."""

    def __init__(self, index, n_slots, B, max_tokens, vocab, shift=SHIFT, first_slot=0, buffers=None):
        import torch
        dev = torch.device("cuda", index.device_index)
        words = (vocab + 31) // 32
        self.index, self.n, self.B, self.first = index, n_slots, B, first_slot
        if buffers is None:
            buffers = [(torch.zeros(max_tokens, dtype=torch.int32, device=dev), torch.zeros(B + 1, dtype=torch.int64, device=dev))
                       for _ in range(n_slots)]
        self.toks = [b[0] for b in buffers]
        self.offs = [b[1] for b in buffers]
        self.sp = [torch.empty(B, dtype=torch.int64, device=dev) for _ in range(n_slots)]
        self.ep = [torch.empty(B, dtype=torch.int64, device=dev) for _ in range(n_slots)]
        self.mask = [torch.empty((B, words), dtype=torch.int32, device=dev) for _ in range(n_slots)]
        L = N.lib()
        for i in range(n_slots):
            N.check(L.fmidx_query_graph_prepare_device(index._h, first_slot + i, C.c_void_p(self.toks[i].data_ptr()),
                                                       C.c_void_p(self.offs[i].data_ptr()), B, shift, vocab,
                                                       C.c_void_p(self.sp[i].data_ptr()), C.c_void_p(self.ep[i].data_ptr()),
                                                       C.c_void_p(self.mask[i].data_ptr())))
        self._launch, self._wait = L.fmidx_query_graph_launch_on, L.fmidx_query_wait

    def launch(self, i, stream=None):
        """Replay slot i's step on `stream` (a torch stream; None: the slot's own stream, see wait)."""
        if self._launch(self.index._h, self.first + i, C.c_void_p(stream.cuda_stream) if stream is not None else None) != 0:
            N.check(-1)

    def wait(self, i):
        if self._wait(self.index._h, self.first + i) != 0:
            N.check(-1)
