"""brute.py -- brute-force model of the reference FM-index semantics (TEST INFRASTRUCTURE).

Pure Python, small inputs only.  Independent of oracle/fm_oracle.c: no wavelet tree, no rank
structure, no sampling -- just sorted suffixes and list counting.  It pins the C oracle (and,
through it, the CUDA path) on random small corpora and produced the worked example of
SURVEY.md section 9 / tests/golden/worked_example.json.

Follows, by definition rather than by algorithm:
  /root/reference/index.py:39-67   (reversal, SHIFT, beginnings)
  /root/reference/index.py:103-207 (get_range, get_count, get_distinct*, get_doc*, __len__)
  /root/reference/cpp_modules_linux/fm_index.cpp:50-194 (size, backward_search_step, distinct*,
  locate, extract_text -- including its isa[end-1] off-by-one, reproduced as written)
PARITY UNPINNED against a real sdsl-lite build (none available here); see fm_oracle.c header.
"""
import bisect

SHIFT = 10  # index.py:16
U64_MAX = (1 << 64) - 1


class BruteFMIndex:
    def __init__(self, docs=None, raw_symbols=None):
        """docs: list of token-id lists (index.py:39-67 path) or raw_symbols: the already
        shifted/reversed symbol list handed to FMIndex::initialize (fm_index.cpp:33-41)."""
        self.beginnings = [0]
        if raw_symbols is None:
            raw_symbols = []
            for d in docs:
                self.beginnings.append(self.beginnings[-1] + len(d))
                raw_symbols.extend(x + SHIFT for x in d[::-1])
        if any(s == 0 for s in raw_symbols):
            raise ValueError("text contains zero symbol")  # sdsl construct() rejects it
        self.T = list(raw_symbols) + [0]  # sdsl construct appends the sentinel
        self.N = len(self.T)
        T = self.T
        self.SA = sorted(range(self.N), key=lambda i: T[i:])
        self.ISA = [0] * self.N
        for r, p in enumerate(self.SA):
            self.ISA[p] = r
        self.BWT = [T[p - 1] for p in self.SA]  # p == 0 -> T[-1] == sentinel
        self._sorted = sorted(T)

    # ---- fm_index.cpp surface -------------------------------------------------------------
    def size(self):
        return self.N

    def _C(self, c):
        return bisect.bisect_left(self._sorted, c)

    def backward_search_step(self, symbol, low, high):
        """fm_index.cpp:67-97 (inclusive in, half-open out)."""
        if high >= self.N:
            high = self.N - 1
        if low > high:
            return (0, 0)
        if symbol > 0 and symbol not in self._sorted_set():
            return (0, 0)  # sdsl: char2comp == 0 and c > 0 -> l=1, r=0 -> wrapper returns (0,0)
        l = self._C(symbol) + self.BWT[:low].count(symbol)
        r = self._C(symbol) + self.BWT[:high + 1].count(symbol) - 1
        if l > r:
            return (0, 0)
        return (l, r + 1)

    def _sorted_set(self):
        if not hasattr(self, "_sset"):
            self._sset = set(self.T)
        return self._sset

    def distinct_count(self, low, high):
        """fm_index.cpp:119-144: interleaved [sym,count,...], ascending symbols."""
        if low == high:
            return ()
        if high > self.N:
            high = self.N
        if low >= high:
            return ()
        seg = self.BWT[low:high]
        out = []
        for s in sorted(set(seg)):
            out += [s, seg.count(s)]
        return tuple(out)

    def distinct(self, low, high):
        return tuple(self.distinct_count(low, high)[0::2])

    def distinct_count_multi(self, lows, highs):
        return tuple(self.distinct_count(l, h) for l, h in zip(lows, highs))

    def locate(self, row):
        """fm_index.cpp:168-172."""
        if row >= self.N:
            return U64_MAX
        return self.SA[row]

    def extract_text(self, begin, end):
        """fm_index.cpp:174-194 as written: starts from isa[end-1]."""
        if end <= begin:
            return ()
        if end > self.N:
            end = self.N
        if end <= begin:
            return ()
        start = self.ISA[end - 1]
        out = [self.BWT[start]]
        for _ in range(end - begin - 1):
            # LF step: backward_search_step(symbol, start, start+1)[0]
            sym = self.BWT[start]
            start = self._C(sym) + self.BWT[:start].count(sym)
            out.append(self.BWT[start])
        return tuple(out)

    # ---- index.py surface -----------------------------------------------------------------
    def __len__(self):
        return self.beginnings[-1]

    def get_range(self, sequence):
        """index.py:103-114."""
        start_row, end_row = 0, self.size()
        for token in sequence:
            start_row, end_row = self.backward_search_step(token + SHIFT, start_row, end_row - 1)
            if start_row == end_row == 0:
                return start_row, end_row
        return start_row, end_row

    def get_range_by_definition(self, sequence):
        """Rows whose suffix starts with reverse(sequence)+SHIFT -- no LF mapping at all."""
        pat = [t + SHIFT for t in sequence[::-1]]
        rows = [r for r, p in enumerate(self.SA) if self.T[p:p + len(pat)] == pat]
        if not rows:
            return (0, 0)
        assert rows == list(range(rows[0], rows[-1] + 1))
        return (rows[0], rows[-1] + 1)

    def get_count(self, sequence):
        s, e = self.get_range(sequence)
        return e - s

    def get_distinct_count(self, low, high):
        """index.py:147-159."""
        data = self.distinct_count(low, high)
        distinct, counts = [], []
        for d, c in zip(data[0::2], data[1::2]):
            if d > 0:
                distinct.append(d - SHIFT)
                counts.append(c)
        return distinct, counts

    def get_distinct(self, low, high):
        return [c - SHIFT for c in self.distinct(low, high) if c > 0]

    def get_doc(self, i):
        """index.py:69-76."""
        return [x - SHIFT for x in self.extract_text(self.beginnings[i], self.beginnings[i + 1])]

    def get_doc_index(self, token_index):
        return bisect.bisect_right(self.beginnings, token_index) - 1

    def occurring_distinct_counts(self):
        """index.py:67,206: over [0, len(self)) -- excludes the last BWT row."""
        return self.get_distinct_count(0, len(self))
