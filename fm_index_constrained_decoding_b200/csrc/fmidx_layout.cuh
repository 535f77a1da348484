// fmidx_layout.cuh -- device-resident FM-index layout for B200 (sm_100a) and the rank primitives.
//
// Replaces sdsl::csa_wt_int<> (reference cpp_modules_linux/fm_index.hpp:12,36): the level-wise
// binary wt_int (17 levels, 3 rank probes per level) becomes a 16-ary WAVELET MATRIX over the
// token BWT: ceil(17/4) = 5 levels for the Llama-3 alphabet (128,256 + SHIFT symbols < 2^17).
//
// One level = an array of 128-byte RANK CELLS, one cell per 128 BWT positions, i.e. exactly one
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
// After the last level,  C[c] + rank_c(i) = D[c] + descend_c(i)  with  D[c] = C[c] - descend_c(0)
// (SURVEY.md section 7 identity (i), generalised from bits to nibbles), so one backward-search
// boundary update costs `levels` dependent cell probes and one L2-resident table read.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define FMIDX_CELL_U4 8          // uint4 per cell (128 B)
#define FMIDX_BLOCK 128          // BWT positions per cell
#define FMIDX_MAX_LEVELS 6       // symbols < 2^24

struct FmDev {
    const uint4 *cells;          // [levels][nblocks][8]
    const uint32_t *D;           // [max_sym+1]  C[c] - descend_c(0)  (mod 2^32)
    const uint32_t *C;           // [max_sym+2]  C[c] = #symbols < c ; C[max_sym+1] = n
    const uint2 *sa_marks;       // [ceil(n/32)] {ones before this word, bits}: row has SA[row] % sa_dens == 0
    const uint32_t *sa_samples;  // sa_dens==1: SA[row] ; else SA value of the k-th marked row
    const uint32_t *isa_samples; // ISA[k*isa_dens]
    const uint32_t *text;        // optional resident text T (n entries) or nullptr
    uint64_t level_stride;       // uint4 per level = nblocks*8
    uint32_t n;                  // csa.size(): text length incl. the 0 sentinel
    uint32_t max_sym;
    uint32_t levels;
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

__device__ __forceinline__ uint32_t fm_digit(uint32_t c, uint32_t levels, uint32_t l) {
    return (c >> (4u * (levels - 1u - l))) & 15u;
}
