"""Reader/converter for indexes that the REFERENCE built: `FMIndex::save` = sdsl `store_to_file`
of a `csa_wt_int<>` (reference fm_index.cpp:196-199), loaded there by `load_FMIndex` (:201-209).

sdsl-lite is not available in this environment and the reference ships no `.fmi` file, so the
serialisation layout below is RESTATED FROM MEMORY of sdsl-lite v2.1.1 and is UNVERIFIED against a
real file: csa_wt::serialize = wavelet_tree, sa_sample, isa_sample, alphabet;
  wt_int        u64 size, u64 sigma, bit_vector tree, rank_support_v, select_1, select_0, u32 max_level
  bit_vector    u64 bit count, ceil(bits/64) u64 words
  int_vector<0> u64 bit count, u8 width, ceil(bits/64) u64 words
  rank_support_v            = int_vector<64> (u64 bit count + words)
  select_support_mcl        = u64 arg_cnt [, int_vector<0> superblock, bit_vector mini_or_long,
                              one int_vector<0> per superblock]
  int_alphabet  bit_vector m_char, rank, select, int_vector<0> m_C, u64 sigma
The parser is strict: every section must be self-consistent and the file must be consumed to the
last byte, so a layout mismatch fails loudly instead of converting garbage.

Conversion is bit-exact by construction: the wavelet tree is decoded back to the BWT, the BWT is
inverted to the text, and the text is handed to the normal on-device builder -- every query
result is a function of the text alone.  Only the heavy steps use torch (on the index's GPU).
"""
import struct

import numpy as np


class SdslFormatError(ValueError):
    pass


class _Reader:
    def __init__(self, path):
        self.buf = np.fromfile(path, dtype=np.uint8)
        self.off = 0

    def _take(self, n):
        if self.off + n > self.buf.size:
            raise SdslFormatError("unexpected end of file at byte %d (+%d)" % (self.off, n))
        out = self.buf[self.off:self.off + n]
        self.off += n
        return out

    def u64(self):
        return int(struct.unpack("<Q", self._take(8).tobytes())[0])

    def u32(self):
        return int(struct.unpack("<I", self._take(4).tobytes())[0])

    def u8(self):
        return int(self._take(1)[0])

    def words(self, nbits):
        return np.frombuffer(self._take(((nbits + 63) // 64) * 8).tobytes(), dtype=np.uint64)

    def bit_vector(self):
        nbits = self.u64()
        return nbits, self.words(nbits)

    def int_vector0(self):
        nbits = self.u64()
        width = self.u8()
        if width == 0 or width > 64 or nbits % width:
            raise SdslFormatError("int_vector<0>: bad width %d for %d bits at byte %d" % (width, nbits, self.off))
        return nbits // width, width, self.words(nbits)

    def int_vector64(self):
        nbits = self.u64()
        if nbits % 64:
            raise SdslFormatError("int_vector<64>: bit count %d not a multiple of 64" % nbits)
        return self.words(nbits)

    def select_mcl(self, expect_args=None):
        cnt = self.u64()
        if expect_args is not None and cnt != expect_args:
            raise SdslFormatError("select support: %d args, expected %d" % (cnt, expect_args))
        if cnt == 0:
            return
        sb = (cnt + 4095) >> 12
        n, _, _ = self.int_vector0()
        if n != sb:
            raise SdslFormatError("select support: %d superblocks, expected %d" % (n, sb))
        nb, _ = self.bit_vector()
        if nb not in (0, sb):
            raise SdslFormatError("select support: mini_or_long has %d bits" % nb)
        for _ in range(sb):
            self.int_vector0()


def unpack_int_vector(words, count, width):
    """int_vector<0> payload -> np.uint64[count]."""
    if count == 0:
        return np.zeros(0, dtype=np.uint64)
    idx = np.arange(count, dtype=np.uint64) * np.uint64(width)
    w, sh = (idx >> np.uint64(6)).astype(np.int64), idx & np.uint64(63)
    words = np.concatenate([words, np.zeros(1, dtype=np.uint64)])
    lo = words[w] >> sh
    hi = np.where(sh > 0, words[w + 1] << ((np.uint64(64) - sh) & np.uint64(63)), np.uint64(0))
    v = lo | hi
    if width < 64:
        v &= np.uint64((1 << width) - 1)
    return v


def read_csa_wt_int(path):
    """Parse an sdsl `csa_wt_int<>` dump.  Returns a dict of its parts."""
    r = _Reader(path)
    size, sigma = r.u64(), r.u64()
    tree_bits, tree = r.bit_vector()
    rank_blocks = r.int_vector64()
    # sdsl sizes the block array from the bit vector's CAPACITY (bits rounded up to whole 64-bit words) [sdsl, recalled]
    cap_bits = ((tree_bits + 63) >> 6) << 6
    if rank_blocks.size != ((cap_bits >> 9) + 1) * 2:
        raise SdslFormatError("rank_support_v has %d words for %d bits" % (rank_blocks.size, tree_bits))
    ones = int(np.unpackbits(tree.view(np.uint8)).sum()) if tree_bits < (1 << 33) else None
    r.select_mcl(ones)
    r.select_mcl(None if ones is None else tree_bits - ones)
    max_level = r.u32()
    if size == 0 or max_level == 0 or max_level > 64 or tree_bits != size * max_level:
        raise SdslFormatError("wt_int: size %d x max_level %d != %d tree bits" % (size, max_level, tree_bits))
    n_sa, w_sa, sa_words = r.int_vector0()
    n_isa, w_isa, isa_words = r.int_vector0()
    char_bits, m_char = r.bit_vector()
    r.int_vector64()
    r.select_mcl()
    n_c, w_c, c_words = r.int_vector0()
    sigma2 = r.u64()
    if r.off != r.buf.size:
        raise SdslFormatError("%d trailing bytes: not a csa_wt_int<> dump in the recalled layout" % (r.buf.size - r.off))
    if sigma2 != sigma or n_c != sigma + 1:
        raise SdslFormatError("alphabet: sigma %d / %d, C has %d entries" % (sigma, sigma2, n_c))
    return {"size": size, "sigma": sigma, "max_level": max_level, "tree": tree,
            "sa_samples": unpack_int_vector(sa_words, n_sa, w_sa), "isa_samples": unpack_int_vector(isa_words, n_isa, w_isa),
            "C": unpack_int_vector(c_words, n_c, w_c), "char_bits": char_bits, "m_char": m_char}


def decode_bwt(tree_words, n, max_level, device="cpu"):
    """wt_int's concatenated level bit-vectors (nodes of a level ordered by prefix, stable inside
    nodes) -> the sequence it represents (the BWT), as an int64 torch tensor."""
    import torch
    bits_all = torch.from_numpy(np.unpackbits(tree_words.view(np.uint8), bitorder="little")[:n * max_level].copy()).to(device)
    orig = torch.arange(n, device=device)            # original BWT position of the element at each slot
    prefix = torch.zeros(n, dtype=torch.int64, device=device)
    for k in range(max_level):
        bit = bits_all[k * n:(k + 1) * n].to(torch.int64)
        prefix = prefix * 2 + bit
        if k + 1 < max_level:
            order = torch.sort(prefix, stable=True).indices  # slots are grouped by prefix already: stable partition
            prefix, orig = prefix[order], orig[order]
    bwt = torch.empty(n, dtype=torch.int64, device=device)
    bwt[orig] = prefix
    return bwt


def invert_bwt(bwt):
    """BWT of a text that ends with a unique smallest sentinel -> the text (sentinel included).
    LF = rank of each row in the stable sort of the BWT; the text position of every row comes from
    list ranking of the single LF cycle by pointer doubling (log2 n gather rounds)."""
    import torch
    n = bwt.numel()
    order = torch.sort(bwt, stable=True).indices
    lf = torch.empty(n, dtype=torch.int64, device=bwt.device)
    lf[order] = torch.arange(n, device=bwt.device)
    # steps(r) = number of LF steps from row r to row 0 (the sentinel suffix, text position n-1)
    nxt = lf.clone()
    w = torch.ones(n, dtype=torch.int64, device=bwt.device)
    nxt[0] = 0
    w[0] = 0
    for _ in range(max(1, int(n - 1).bit_length())):
        w = w + w[nxt]
        nxt = nxt[nxt]
    if int(nxt.max()) != 0:
        raise SdslFormatError("BWT is not a single LF cycle: not the BWT of a sentinel-terminated text")
    # row r holds suffix p = (steps-1) mod n ; BWT[r] = T[p-1]
    at = (w - 2) % n
    text = torch.empty(n, dtype=torch.int64, device=bwt.device)
    text[at] = bwt
    if int(text[n - 1]) != 0 or int((text[:n - 1] == 0).sum()) != 0:
        raise SdslFormatError("decoded text does not end in a unique 0 sentinel")
    return text


def text_from_sdsl_file(path, device="cpu"):
    """Reference-built `.fmi` -> the symbol text (WITHOUT the sentinel) it indexes, plus the parsed parts
    (whose C array and samples are cross-checked against the decoded text)."""
    import torch
    parts = read_csa_wt_int(path)
    n = parts["size"]
    bwt = decode_bwt(parts["tree"], n, parts["max_level"], device)
    text = invert_bwt(bwt)
    # cross-checks: the alphabet's C array and the SA samples must agree with the decoded text
    syms, counts = torch.unique(text, return_counts=True)
    if syms.numel() != parts["sigma"]:
        raise SdslFormatError("decoded text has %d distinct symbols, header says %d" % (syms.numel(), parts["sigma"]))
    c = torch.zeros(syms.numel() + 1, dtype=torch.int64, device=text.device)
    c[1:] = torch.cumsum(counts, 0)
    if not np.array_equal(c.cpu().numpy().astype(np.uint64), parts["C"]):
        raise SdslFormatError("C array of the file does not match the decoded text")
    return text[:n - 1], parts


def is_own_container(path):
    with open(path, "rb") as f:
        return f.read(8) == b"FMIDXB2\0"
