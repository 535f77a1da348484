// fmidx_layout.cuh -- device-resident FM-index layout for B200 (sm_100a) and the rank primitives.
//
// Replaces sdsl::csa_wt_int<> (reference cpp_modules_linux/fm_index.hpp:12,36): the level-wise
// binary wt_int (17 levels, 3 rank probes per level) becomes a multi-ary WAVELET MATRIX over the
// token BWT whose digits are, most significant first, `levels` NIBBLES and one final BYTE:
// 3 nibble levels + 1 byte level = 4 dependent probes per rank for the Llama-3 alphabet
// (128,256 + SHIFT symbols < 2^17 -> bits 16..8 as three nibbles, bits 7..0 as the byte).
//
// Nibble level = an array of 128-byte RANK CELLS, one per 128 BWT positions, i.e. exactly one
// L2 line / four 32-B DRAM sectors, fetched by eight lanes with one 128-bit load each:
//
//   u32 hdr[16]      hdr[d] = (#symbols of this level whose digit < d)              "Z_d"
//                           + (#positions before this cell whose digit == d)
//                    = the position, in the NEXT level's order, that the first digit-d
//                      element at or after this cell maps to.  (interleaved superblock counts;
//                      the per-level Z table is folded in so a probe needs no side table)
//   u32 plane[4][4]  plane[g][k] bit t = bit k of the 4-bit digit at position 128*cell+32*g+t
//                    (bit-sliced nibbles: "digit == d" for 32 positions is 3 logic ops + POPC)
//
//   next_pos(pos, d) = hdr[d] + popc(match_d(planes) & below(pos & 127))
//
// Byte level (the last one) = 1,536-byte cells, one per 512 positions: u32 hdr[256] with the same
// meaning, then the 512 digits as plain bytes.  A rank probe reads one header word and the bytes
// below pos (SIMD byte compare, 16 per 128-bit load); "access" is a single byte load; and -- the
// reason this level exists -- enumerating the distinct symbols of a range needs ONE histogram
// over at most two cells for 256 children instead of 16 sixteen-way node evaluations.
//
// After the last level,  C[c] + rank_c(i) = D[c] + descend_c(i)  with  D[c] = C[c] - descend_c(0)
// (SURVEY.md section 7 identity (i), generalised from bits to digits), so one backward-search
// boundary update costs `levels + 1` dependent cell probes and one L2-resident table read.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define FMIDX_CELL_U4 8          // uint4 per cell (128 B)
#define FMIDX_BLOCK 128          // BWT positions per cell
#define FMIDX_MAX_LEVELS 6       // symbols < 2^24 (4 nibble levels + the byte level)
#define FMIDX_BBLOCK 512         // BWT positions per byte-level cell
#define FMIDX_BCELL_U4 96        // uint4 per byte-level cell: 64 header + 32 payload
#define FMIDX_BHDR_U4 64

struct FmDev {
    const uint4 *cells;          // nibble levels [levels][nblocks][8]
    const uint4 *bcells;         // byte level [nbblocks][96]
    const uint32_t *D;           // [max_sym+1]  C[c] - descend_c(0)  (mod 2^32)
    const uint32_t *C;           // [max_sym+2]  C[c] = #symbols < c ; C[max_sym+1] = n
    const uint2 *sa_marks;       // [ceil(n/32)] {ones before this word, bits}: row has SA[row] % sa_dens == 0
    const uint32_t *sa_samples;  // sa_dens==1: SA[row] ; else SA value of the k-th marked row
    const uint32_t *isa_samples; // ISA[k*isa_dens]
    const uint32_t *text;        // optional resident text T (n entries) or nullptr
    const uint32_t *bwt;         // optional resident BWT as plain u32 symbols (n entries) or nullptr: the rows of a range are
                                 // CONTIGUOUS here, so the distinct symbols of a narrow / mid-sized range are one coalesced read
    // unigram continuation table (resident layout): for the symbols c with the most occurrences, the distinct symbols of
    // BWT[C[c], C[c+1]) as a bitmap over the SYMBOL axis -- the continuations of a one-token prefix are a property of the
    // index (the reference precomputes the same thing for the EMPTY prefix: occurring_distinct, index.py:206), and the
    // one-token prefixes of frequent tokens are exactly the ranges whose enumeration costs the most
    const uint32_t *uni_sp;      // [uni_n] C[c], ascending
    const uint32_t *uni_ep;      // [uni_n] C[c+1]
    const uint32_t *uni_masks;   // [uni_n][uni_words] bit s of a row: symbol s occurs in the block
    uint32_t uni_n, uni_words;
    uint64_t level_stride;       // uint4 per level = nblocks*8
    uint32_t n;                  // csa.size(): text length incl. the 0 sentinel
    uint32_t max_sym;
    uint32_t levels;             // number of NIBBLE levels above the byte level (0..4)
    uint32_t nbblocks;
    uint32_t sa_dens;
    uint32_t isa_dens;
    uint32_t nblocks;
};

__device__ __forceinline__ uint4 fm_ldg(const uint4 *p) { return __ldg(p); }

__device__ __forceinline__ uint32_t fm_pick(const uint4 &h, uint32_t i) {
    return i == 0 ? h.x : (i == 1 ? h.y : (i == 2 ? h.z : h.w));
}

// positions (of 32) whose digit == d, from the four bit-planes of one group
__device__ __forceinline__ uint32_t fm_match(const uint4 &p, uint32_t d) {
    uint32_t m0 = (d & 1u) ? 0u : ~0u, m1 = (d & 2u) ? 0u : ~0u, m2 = (d & 4u) ? 0u : ~0u, m3 = (d & 8u) ? 0u : ~0u;
    return (p.x ^ m0) & (p.y ^ m1) & (p.z ^ m2) & (p.w ^ m3);
}

__device__ __forceinline__ uint32_t fm_below(int nbits) {  // nbits in [1,32]
    return nbits >= 32 ? ~0u : ((1u << nbits) - 1u);
}

// One level of descend for digit d, cooperatively by the 8 lanes sub=0..7 of an aligned lane
// octet.  Lanes 0-3 own the header quarter that holds hdr[d]; lanes 4-7 own one 32-position
// group each and only load it when it lies below pos.  Returns the next-level position in
// all 8 lanes.  `act` predicates the loads; the shuffles are executed unconditionally.
__device__ __forceinline__ uint32_t fm_step8(const uint4 *__restrict__ level_cells, uint32_t pos, uint32_t d,
                                             int sub, bool act) {
    const uint4 *cell = level_cells + (size_t)(pos >> 7) * FMIDX_CELL_U4;
    const int off = (int)(pos & 127u);
    uint32_t v = 0;
    if (act) {
        if (sub < 4) {
            if ((uint32_t)sub == (d >> 2)) v = fm_pick(fm_ldg(cell + sub), d & 3u);
        } else {
            const int nb = off - 32 * (sub - 4);
            if (nb > 0) v = __popc(fm_match(fm_ldg(cell + sub), d) & fm_below(nb));
        }
    }
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

// Same, additionally returning the digit stored AT pos (access + rank in one probe: LF step).
__device__ __forceinline__ uint32_t fm_access_step8(const uint4 *__restrict__ level_cells, uint32_t pos, int sub,
                                                   bool act, uint32_t &digit_out) {
    const uint4 *cell = level_cells + (size_t)(pos >> 7) * FMIDX_CELL_U4;
    const int off = (int)(pos & 127u);
    const int g_at = off >> 5;
    uint4 h = make_uint4(0, 0, 0, 0), p = make_uint4(0, 0, 0, 0);
    if (act) {
        if (sub < 4) h = fm_ldg(cell + sub);                      // digit unknown yet: all four header quarters
        else if (sub - 4 <= g_at) p = fm_ldg(cell + sub);
    }
    // the lane owning group g_at extracts the digit and broadcasts it inside the octet
    uint32_t dig = 0;
    if (sub - 4 == g_at) {
        const int t = off & 31;
        dig = ((p.x >> t) & 1u) | (((p.y >> t) & 1u) << 1) | (((p.z >> t) & 1u) << 2) | (((p.w >> t) & 1u) << 3);
    }
    const int lane = threadIdx.x & 31;
    dig = __shfl_sync(0xffffffffu, dig, (lane & ~7) + 4 + g_at);
    uint32_t v = 0;
    if (act) {
        if (sub < 4) {
            if ((uint32_t)sub == (dig >> 2)) v = fm_pick(h, dig & 3u);
        } else {
            const int nb = off - 32 * (sub - 4);
            if (nb > 0) v = __popc(fm_match(p, dig) & fm_below(nb));
        }
    }
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    digit_out = dig;
    return v;
}

// Single-thread variants (build-time D table, tiny scalar paths).
__device__ __forceinline__ uint32_t fm_step1(const uint4 *__restrict__ level_cells, uint32_t pos, uint32_t d) {
    const uint4 *cell = level_cells + (size_t)(pos >> 7) * FMIDX_CELL_U4;
    const int off = (int)(pos & 127u);
    uint32_t v = fm_pick(fm_ldg(cell + (d >> 2)), d & 3u);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const int nb = off - 32 * g;
        if (nb > 0) v += __popc(fm_match(fm_ldg(cell + 4 + g), d) & fm_below(nb));
    }
    return v;
}

// nibble digit of symbol c examined at nibble level l (levels = number of nibble levels)
__device__ __forceinline__ uint32_t fm_digit(uint32_t c, uint32_t levels, uint32_t l) {
    return (c >> (8u + 4u * (levels - 1u - l))) & 15u;
}

// ---- byte level ----------------------------------------------------------------------------------
// bytes equal to the replicated pattern among the first nb (0..16) bytes of a 16-byte group
__device__ __forceinline__ uint32_t fm_count16(const uint4 &w, uint32_t pat, int nb) {
    uint32_t m0 = __vcmpeq4(w.x, pat), m1 = __vcmpeq4(w.y, pat), m2 = __vcmpeq4(w.z, pat), m3 = __vcmpeq4(w.w, pat);
    if (nb < 16) {
        const int r = nb & 3, q = nb >> 2;  // q full words, then r bytes of word q
        const uint32_t part = r ? ((1u << (8 * r)) - 1u) : 0u;
        m0 = q > 0 ? m0 : (q == 0 ? (m0 & part) : 0u);
        m1 = q > 1 ? m1 : (q == 1 ? (m1 & part) : 0u);
        m2 = q > 2 ? m2 : (q == 2 ? (m2 & part) : 0u);
        m3 = q > 3 ? m3 : (q == 3 ? (m3 & part) : 0u);
    }
    return (uint32_t)(__popc(m0) + __popc(m1) + __popc(m2) + __popc(m3)) >> 3;
}

// rank step at the byte level, cooperatively by an aligned lane octet: lane 0 fetches hdr[v], every
// lane scans the 16-byte groups sub, sub+8, sub+16, sub+24 that lie below pos (all loads issued first).
__device__ __forceinline__ uint32_t fm_step8_byte(const uint4 *__restrict__ bcells, uint32_t pos, uint32_t v, int sub,
                                                  bool act) {
    const uint4 *cell = bcells + (size_t)(pos >> 9) * FMIDX_BCELL_U4;
    const int off = (int)(pos & 511u);
    uint32_t cnt = 0;
    if (act) {
        uint4 w[4];
        int nb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int j = sub + 8 * i;
            nb[i] = min(16, off - 16 * j);
            if (nb[i] > 0) w[i] = fm_ldg(cell + FMIDX_BHDR_U4 + j);
        }
        if (sub == 0) cnt = __ldg(reinterpret_cast<const uint32_t *>(cell) + v);
        const uint32_t pat = v * 0x01010101u;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (nb[i] > 0) cnt += fm_count16(w[i], pat, nb[i]);
    }
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 4);
    return cnt;
}

__device__ __forceinline__ uint32_t fm_byte_at(const uint4 *__restrict__ bcells, uint32_t pos) {
    const uint4 *cell = bcells + (size_t)(pos >> 9) * FMIDX_BCELL_U4;
    return __ldg(reinterpret_cast<const unsigned char *>(cell + FMIDX_BHDR_U4) + (pos & 511u));
}

// single-thread rank step at the byte level (build-time D table)
__device__ __forceinline__ uint32_t fm_step1_byte(const uint4 *__restrict__ bcells, uint32_t pos, uint32_t v) {
    const uint4 *cell = bcells + (size_t)(pos >> 9) * FMIDX_BCELL_U4;
    const int off = (int)(pos & 511u);
    uint32_t cnt = __ldg(reinterpret_cast<const uint32_t *>(cell) + v);
    const uint32_t pat = v * 0x01010101u;
    for (int j = 0; 16 * j < off; ++j) cnt += fm_count16(fm_ldg(cell + FMIDX_BHDR_U4 + j), pat, min(16, off - 16 * j));
    return cnt;
}

// full descend of symbol c from position pos by one thread / by an octet
__device__ __forceinline__ uint32_t fm_descend1(const FmDev &ix, uint32_t pos, uint32_t c) {
    for (uint32_t l = 0; l < ix.levels; ++l) pos = fm_step1(ix.cells + l * ix.level_stride, pos, fm_digit(c, ix.levels, l));
    return fm_step1_byte(ix.bcells, pos, c & 255u);
}
__device__ __forceinline__ uint32_t fm_descend8(const FmDev &ix, uint32_t pos, uint32_t c, int sub, bool act) {
    for (uint32_t l = 0; l < ix.levels; ++l)
        pos = fm_step8(ix.cells + l * ix.level_stride, pos, fm_digit(c, ix.levels, l), sub, act);
    return fm_step8_byte(ix.bcells, pos, c & 255u, sub, act);
}
