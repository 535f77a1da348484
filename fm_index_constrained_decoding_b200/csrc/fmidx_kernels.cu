// fmidx_kernels.cu -- the sm_100a query kernels of the FM-index hot path.
//
//   K1  k_search   batched backward search         index.py:103-114 over fm_index.cpp:67-97
//       k_step     batched single LF/backward step  fm_index.cpp:67-97
//   K2  k_mask     distinct next tokens of [sp,ep) -> vocab bitmap (+counts), fused with the
//                  -inf masking of the beam's logits row
//                  fm_index.cpp:119-166 + index.py:147-174 + document_constrained_generation_vllm.py:52,105-137
//       k_popcount / k_compact   bitmap(+counts) -> ascending (token,count) lists
//   K3  k_locate   SA-sample locate                 fm_index.cpp:168-172
//   K4  k_extract  text extraction                  fm_index.cpp:174-194 (as written: isa[end-1])
//
// All of it is random-access integer work bounded by HBM/L2 latency and sector throughput; no
// tensor cores.  Thread mapping: an aligned OCTET of lanes owns one rank probe (one 128-B cell,
// one LDG.128 per lane, 3-step xor-shuffle popcount reduction); a half-warp owns one query
// (sp octet + ep octet) or one enumeration node.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "fmidx_internal.h"

namespace {

constexpr unsigned FULL = 0xffffffffu;

// system-scope flag accesses of the peer-memory exchange (one process per GPU; flags are monotonically increasing step numbers)
__device__ __forceinline__ uint32_t fm_ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fm_st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// =================================================================================================
// K1: backward search.  16 lanes per prefix: octet 0 chases sp, octet 1 chases ep, both through
// the same `levels` cells per token; D[c] and the next token are fetched ahead of the descent.
// RAW=false : index.FMIndex.get_range semantics (early exit with (0,0), index.py:111-112)
// RAW=true  : FMIndex::backward_search_multi semantics (fm_index.cpp:55-65; no guards, an absent
//             symbol gives l=1,r=0 and the arithmetic just continues)
// =================================================================================================
// token sources of the search: the batch's token array in global memory / the tokens of ONE prefix handed over as kernel
// parameters (scalar calls: no H2D copy in front of the launch)
struct FmToksGlobal {
    const int32_t *__restrict__ p;
    __device__ __forceinline__ int32_t at(int64_t i) const { return p[i]; }
};
#define FM_SCALAR_MAX_TOKENS 96
struct FmToksParam {
    int32_t t[FM_SCALAR_MAX_TOKENS];
};
struct FmToksShared {
    const int32_t *p;  // shared memory
    __device__ __forceinline__ int32_t at(int64_t i) const { return p[i]; }
};

// the search of prefix q by its 16 lanes (whole warps call this; q >= B: idle lanes)
template <bool RAW, class Toks>
__device__ __forceinline__ void fm_search_one(const FmDev &ix, const Toks toks, int64_t o0, int len, bool valid, uint32_t q, int64_t shift,
                                              uint64_t *__restrict__ sp, uint64_t *__restrict__ ep,
                                              uint32_t *__restrict__ zero_rows, uint32_t zero_words, uint32_t zero_lo,
                                              uint32_t zero_n) {
    const int lane = threadIdx.x & 31, sub = lane & 7, half = (lane >> 3) & 1;
    // fmidx_query_batch: the mask row of this prefix is zeroed HERE (16 lanes, 16-byte stores, issued before the
    // dependent-load chain starts and never waited for), so the enumeration kernels find pre-zeroed rows and the
    // latency-bound search hides the 16 KB of store traffic per prefix.  zero_words % 4 == 0 and a 16-byte aligned
    // base are checked by the launcher.
    if (zero_rows && valid && q - zero_lo < zero_n) {
        uint4 *row = reinterpret_cast<uint4 *>(zero_rows + (size_t)q * zero_words);
        const uint32_t nv = zero_words >> 2;
        for (uint32_t i = (uint32_t)(lane & 15); i < nv; i += 16) row[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    const int other_len = __shfl_xor_sync(FULL, len, 16);
    const int maxlen = len > other_len ? len : other_len;
    uint32_t lo = 0, hi = ix.n;
    bool alive = valid;
    // When the full SA, the full ISA and the text are resident, a range that has narrowed to <= 16 rows
    // is finished without further rank probes: lane i of the query's half-warp takes row lo+i, whose
    // suffix starts at p_i = SA[lo+i]; the remaining tokens must equal T[p_i-1], T[p_i-2], ...
    // (BWT[row] = T[SA[row]-1], LF(row) = ISA[SA[row]-1]); the surviving occurrences' final rows
    // ISA[p_i - remaining] form the (contiguous) answer range.  ~4 dependent round trips instead of
    // 5 per remaining token.
    const bool accel = !RAW && ix.text != nullptr && ix.sa_dens == 1 && ix.isa_dens == 1;
    int t_stop = -1;  // >= 0: range fell to <= 16 rows after token t_stop-1; tokens t_stop.. are pending
    int32_t tok_next = (valid && len > 0) ? toks.at(o0) : 0;
    for (int t = 0; t < maxlen; ++t) {
        const bool act = alive && t < len;
        if (!__any_sync(FULL, act)) break;
        const int64_t sym = (int64_t)tok_next + shift;
        if (act && t + 1 < len) tok_next = toks.at(o0 + t + 1);
        const bool symok = act && sym >= 0 && sym <= (int64_t)ix.max_sym;
        const uint32_t c = symok ? (uint32_t)sym : 0u;
        // RAW: sdsl::backward_search answers l=1,r=0 for a symbol that is not in the text
        const bool present = RAW ? (symok && ix.C[c + 1] > ix.C[c]) : symok;
        uint32_t nlo, nhi;
        if (t == 0) {
            // whole range: sdsl::backward_search's (l==0 && r+1==size) shortcut -> [C[c], C[c+1])
            nlo = present ? ix.C[c] : 1u;
            nhi = present ? ix.C[c + 1] : 1u;
        } else {
            const uint32_t Dc = symok ? ix.D[c] : 0u;
            uint32_t pos = half ? hi : lo;
            pos = fm_descend8(ix, pos, c, sub, symok);
            const uint32_t res = present ? Dc + pos : 1u;
            const uint32_t oth = __shfl_xor_sync(FULL, res, 8);
            nlo = half ? oth : res;
            nhi = half ? res : oth;
        }
        if (act) {
            if (RAW) {
                lo = nlo;
                hi = nhi;
            } else if (!symok || nlo >= nhi) {
                lo = 0;
                hi = 0;
                alive = false;
            } else {
                lo = nlo;
                hi = nhi;
                if (accel && hi - lo <= 16u && t + 1 < len) {
                    t_stop = t + 1;
                    alive = false;
                }
            }
        }
    }
    if (accel) {
        // Three dependent round trips for the whole tail, however many tokens remain: (1) lane i < w loads p_i = SA[lo+i]
        // while every lane loads "its" remaining tokens (hl, hl+16); (2) for every candidate row the half-warp compares
        // the remaining tokens with T[p_i-1], T[p_i-2], ... (lane hl: positions hl and hl+16, i.e. two coalesced 64-byte
        // reads per row; the loads of four rows are issued before any is consumed); (3) the surviving rows' final rows
        // ISA[p_i - rem].  Tails longer than 32 tokens and texts shorter than 64 symbols (the cyclic index could wrap more
        // than once) take the serial loop below.
        const bool pend = t_stop >= 0;  // uniform inside the query's 16 lanes
        const int hl = lane & 15, hbase = lane & 16;
        const int rem = pend ? len - t_stop : 0;
        const uint32_t n = ix.n;
        const uint32_t w = pend ? hi - lo : 0u;  // 1..16
        const bool fast = n >= 64u && rem <= 32;
        bool ok = pend && (uint32_t)hl < w;
        uint32_t p = 0;
        if (ok) p = ix.sa_samples[lo + (uint32_t)hl];
        if (__any_sync(FULL, pend && fast)) {
            int64_t s0 = -1, s1 = -1;
            if (pend && fast && hl < rem) s0 = (int64_t)toks.at(o0 + t_stop + hl) + shift;
            if (pend && fast && hl + 16 < rem) s1 = (int64_t)toks.at(o0 + t_stop + hl + 16) + shift;
            const bool use0 = pend && fast && hl < rem, use1 = pend && fast && hl + 16 < rem;
            // a symbol outside u32 can never equal a text symbol
            const bool inv0 = use0 && (s0 < 0 || s0 > 0xFFFFFFFFll), inv1 = use1 && (s1 < 0 || s1 > 0xFFFFFFFFll);
            const uint32_t wo = __shfl_xor_sync(FULL, fast ? w : 0u, 16);
            const uint32_t wmax = max(fast ? w : 0u, wo);
            uint32_t okmask = 0;
            for (uint32_t i0 = 0; i0 < wmax; i0 += 4) {
                uint32_t t0v[4], t1v[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const uint32_t pi = __shfl_sync(FULL, p, hbase + (int)((i0 + r) & 15u));
                    const bool row = i0 + r < w;
                    t0v[r] = 0u;
                    t1v[r] = 0u;
                    if (row && use0) t0v[r] = ix.text[pi > (uint32_t)hl ? pi - 1u - (uint32_t)hl : pi + n - 1u - (uint32_t)hl];
                    if (row && use1) t1v[r] = ix.text[pi > (uint32_t)hl + 16u ? pi - 17u - (uint32_t)hl : pi + n - 17u - (uint32_t)hl];
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const bool row = i0 + r < w;
                    const bool bad = row && (inv0 || inv1 || (use0 && t0v[r] != (uint32_t)s0) || (use1 && t1v[r] != (uint32_t)s1));
                    const uint32_t anybad = (__ballot_sync(FULL, bad) >> hbase) & 0xFFFFu;
                    if (row && !anybad) okmask |= 1u << (i0 + r);
                }
            }
            if (fast) ok = ok && ((okmask >> hl) & 1u);
        }
        if (pend && !fast) {
            for (int j = 0; j < rem && ok; ++j) {
                const int64_t sym = (int64_t)toks.at(o0 + t_stop + j) + shift;
                const uint32_t at = (uint32_t)(((uint64_t)p + n - 1u - ((uint32_t)j % n)) % n);  // the text is cyclic under LF
                ok = sym >= 0 && sym <= 0xFFFFFFFFll && (uint32_t)sym == ix.text[at];
            }
        }
        uint32_t mn = 0xFFFFFFFFu, mx = 0u;
        if (ok) {
            const uint32_t back = fast ? (p >= (uint32_t)rem ? p - (uint32_t)rem : p + n - (uint32_t)rem)
                                       : (uint32_t)(((uint64_t)p + n - ((uint32_t)rem % n)) % n);
            mn = mx = ix.isa_samples[back];
        }
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            mn = min(mn, __shfl_xor_sync(FULL, mn, o));
            mx = max(mx, __shfl_xor_sync(FULL, mx, o));
        }
        if (pend) {
            if (mn == 0xFFFFFFFFu) {
                lo = 0;
                hi = 0;
            } else {
                lo = mn;
                hi = mx + 1u;
            }
        }
    }
    if (valid && (lane & 15) == 0) {
        sp[q] = lo;
        ep[q] = hi;
    }
}

template <bool RAW>
__global__ void __launch_bounds__(128) k_search(FmDev ix, const int32_t *__restrict__ toks,
                                                const int64_t *__restrict__ offs, uint32_t B, int64_t shift,
                                                uint64_t *__restrict__ sp, uint64_t *__restrict__ ep,
                                                uint32_t *__restrict__ zero_rows, uint32_t zero_words, uint32_t zero_lo,
                                                uint32_t zero_n) {
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const bool valid = q < B;
    int64_t o0 = 0;
    int len = 0;
    if (valid) {
        o0 = offs[q];
        len = (int)(offs[q + 1] - o0);
    }
    fm_search_one<RAW>(ix, FmToksGlobal{toks}, o0, len, valid, q, shift, sp, ep, zero_rows, zero_words, zero_lo, zero_n);
}

// ---- scalar calls (batch 1): arguments travel as kernel parameters, results go straight into MAPPED PINNED host memory and
// the host polls a sequence flag -- no H2D copy, no D2H copy, no stream synchronisation around a one-prefix query.
struct FmScalarOut {            // lives in pinned host memory (device-visible under UVA)
    uint64_t v[4];              // search / step: sp, ep ; locate: position
    uint32_t k;                 // distinct: number of u64 written to data
    uint32_t status;            // distinct: 0 ok, 1 = range too wide for the scalar kernel (caller falls back)
    uint32_t flag;              // sequence number of the call whose results are complete (written last, system scope)
    uint32_t pad;
    uint64_t data[1];           // distinct: (sym[, count]) ascending, up to the capacity given to the kernel
};

__device__ __forceinline__ void fm_scalar_done(FmScalarOut *out, uint32_t seq) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(&out->flag), "r"(seq) : "memory");
}

__device__ __forceinline__ uint32_t fm_access1(const uint4 *__restrict__ level_cells, uint32_t pos, uint32_t &dig_out);  // below

template <bool RAW>
__global__ void __launch_bounds__(32) k_search1(FmDev ix, FmToksParam tp, int len, int64_t shift, FmScalarOut *out, uint32_t seq,
                                                uint64_t *d_range, FmScalarOut *spec_out, uint32_t spec_seq) {
    __shared__ int32_t st[FM_SCALAR_MAX_TOKENS];
    __shared__ uint64_t res[2];
    for (int i = threadIdx.x; i < len; i += 32) st[i] = tp.t[i];
    __syncwarp();
    fm_search_one<RAW>(ix, FmToksShared{st}, 0, len, threadIdx.x < 16, 0u, shift, &res[0], &res[1], nullptr, 0u, 0u, 0u);
    __syncwarp();
    if (threadIdx.x == 0) {
        out->v[0] = res[0];
        out->v[1] = res[1];
        fm_scalar_done(out, seq);
    }
    if (!d_range) return;
    // Speculative enumeration of the range just found (see fm_scalar_search).  Up to 32 rows -- the usual case once a history
    // has a few tokens -- are finished by this warp right here: one lane per row, the row's symbol, then rank-by-counting over
    // the warp (distinct symbols ascending, with their multiplicities) straight into the second pinned block.  Wider ranges are
    // left to the k_distinct1 that follows on the stream (it sees d_range[2] != spec_seq and runs).
    const uint64_t lo = res[0], hi = res[1];
    const uint32_t w = hi > lo ? (uint32_t)min(hi - lo, (uint64_t)64) : 0u;
    if (threadIdx.x == 0) {
        d_range[0] = lo;
        d_range[1] = hi;
    }
    if (w > 32u) return;
    const int lane = threadIdx.x;
    const bool has = (uint32_t)lane < w;
    uint32_t c = 0xFFFFFFFFu;
    if (has) {
        uint32_t pos = (uint32_t)lo + (uint32_t)lane;
        if (ix.bwt) c = __ldg(ix.bwt + pos);
        else {
            c = 0;
            for (uint32_t l = 0; l < ix.levels; ++l) {
                uint32_t dig;
                pos = fm_access1(ix.cells + l * ix.level_stride, pos, dig);
                c = (c << 4) | dig;
            }
            c = (c << 8) | fm_byte_at(ix.bcells, pos);
        }
    }
    const uint32_t same = __match_any_sync(FULL, c);
    const bool head = has && (uint32_t)(__ffs(same) - 1) == (uint32_t)lane;
    uint32_t rank = 0;
#pragma unroll 8
    for (int j = 0; j < 32; ++j) {
        const uint32_t cj = __shfl_sync(FULL, c, j);
        const bool hj = __shfl_sync(FULL, (int)head, j) != 0;
        rank += (hj && cj < c) ? 1u : 0u;
    }
    const uint32_t nd = (uint32_t)__popc(__ballot_sync(FULL, head));
    if (head) {
        spec_out->data[2u * rank] = c;
        spec_out->data[2u * rank + 1u] = (uint64_t)__popc(same);
    }
    __threadfence_system();
    __syncwarp();
    if (lane == 0) {
        spec_out->status = 0u;
        spec_out->k = 2u * nd;
        d_range[2] = spec_seq;  // tells the k_distinct1 behind us that its work is done
        fm_scalar_done(spec_out, spec_seq);
    }
}

// B independent FMIndex::backward_search_step calls (inclusive [low,high] in, half-open out).
__device__ __forceinline__ void fm_step_one(const FmDev &ix, uint64_t sym, uint64_t low, uint64_t high, bool valid, uint64_t *sp_out,
                                            uint64_t *ep_out) {
    const int lane = threadIdx.x & 31, sub = lane & 7, half = (lane >> 3) & 1;
    if (valid && high >= ix.n) high = (uint64_t)ix.n - 1;  // fm_index.cpp:74-76
    const bool ok = valid && low <= high && sym <= ix.max_sym;  // fm_index.cpp:77-82 ; absent symbol
    const uint32_t c = ok ? (uint32_t)sym : 0u;
    const uint32_t Dc = ok ? ix.D[c] : 0u;
    uint32_t pos = ok ? (half ? (uint32_t)high + 1u : (uint32_t)low) : 0u;
    pos = fm_descend8(ix, pos, c, sub, ok);
    const uint32_t res = Dc + pos;
    const uint32_t oth = __shfl_xor_sync(FULL, res, 8);
    const uint32_t nlo = half ? oth : res, nhi = half ? res : oth;
    if (valid && (lane & 15) == 0) {
        const bool hit = ok && nlo < nhi;  // fm_index.cpp:87-92
        *sp_out = hit ? nlo : 0;
        *ep_out = hit ? nhi : 0;
    }
}

__global__ void __launch_bounds__(128) k_step(FmDev ix, const uint64_t *__restrict__ syms,
                                              const uint64_t *__restrict__ lows, const uint64_t *__restrict__ highs,
                                              uint32_t B, uint64_t *__restrict__ sp, uint64_t *__restrict__ ep) {
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const bool valid = q < B;
    uint64_t sym = 0, low = 1, high = 0;
    if (valid) {
        sym = syms[q];
        low = lows[q];
        high = highs[q];
    }
    fm_step_one(ix, sym, low, high, valid, sp + (valid ? q : 0), ep + (valid ? q : 0));
}

// scalar FMIndex::backward_search_step: arguments by value, result into mapped pinned host memory (see k_search1)
__global__ void __launch_bounds__(32) k_step1(FmDev ix, uint64_t sym, uint64_t low, uint64_t high, FmScalarOut *out, uint32_t seq) {
    fm_step_one(ix, sym, low, high, threadIdx.x < 16, &out->v[0], &out->v[1]);
    __syncwarp();
    if (threadIdx.x == 0) fm_scalar_done(out, seq);
}

// Incremental decoding: one backward-search step per beam per generated token, state kept on the device.
// state (sp, ep) of beam r = get_range(force + history_r) as index.py:103-114 computes it ((0,0) once empty, and then for
// good); the new row r continues the state of row parent[r] (beam reordering) with its newest token, read on the device
// out of input_ids (last column, row stride given).  The kernel also prepares what the fused mask kernel needs,
// exactly as document_constrained_generation_vllm.py:78-118 decides it row by row:
//   last token eos / pad             -> query range (0,0), override {pad}
//   stop_at_count > 0 and count(history[:-1]) = width of the PARENT state <= stop_at_count (checked first) -> override {eos}
//   otherwise                        -> query range = new state, no override
__global__ void __launch_bounds__(128)
k_decode_step(FmDev ix, const int64_t *__restrict__ last_tok, int64_t tok_stride, const int32_t *__restrict__ parent,
              const uint64_t *__restrict__ st_sp_in, const uint64_t *__restrict__ st_ep_in, uint64_t *__restrict__ st_sp_out,
              uint64_t *__restrict__ st_ep_out, uint32_t B, int64_t shift, int64_t eos, int64_t pad, int64_t stop_at_count,
              uint64_t *__restrict__ q_sp, uint64_t *__restrict__ q_ep, int32_t *__restrict__ override_tok) {
    const int lane = threadIdx.x & 31, sub = lane & 7, half = (lane >> 3) & 1;
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const bool valid = q < B;
    int64_t tok = 0;
    uint64_t psp = 0, pep = 0;
    if (valid) {
        tok = last_tok[(int64_t)q * tok_stride];
        const uint32_t pr = parent ? (uint32_t)parent[q] : q;
        psp = st_sp_in[pr];
        pep = st_ep_in[pr];
    }
    const int64_t sym = tok + shift;
    const bool alive = valid && pep > psp;                                  // (0,0): the history left the corpus earlier
    const bool ok = alive && sym >= 0 && sym <= (int64_t)ix.max_sym;        // absent symbol -> empty
    const uint32_t c = ok ? (uint32_t)sym : 0u;
    const uint32_t Dc = ok ? ix.D[c] : 0u;
    uint32_t pos = ok ? (half ? (uint32_t)pep : (uint32_t)psp) : 0u;        // step(sym, low, high = ep - 1): rank at low and at high + 1
    pos = fm_descend8(ix, pos, c, sub, ok);
    const uint32_t res = Dc + pos;
    const uint32_t oth = __shfl_xor_sync(FULL, res, 8);
    const uint32_t nlo = half ? oth : res, nhi = half ? res : oth;
    if (valid && (lane & 15) == 0) {
        const bool hit = ok && nlo < nhi;
        const uint64_t nsp = hit ? nlo : 0, nep = hit ? nhi : 0;
        st_sp_out[q] = nsp;
        st_ep_out[q] = nep;
        const bool finished = tok == eos || tok == pad;
        const uint64_t count_prev = finished ? 0 : pep - psp;               // vllm.py:81-92
        int32_t ovr = -1;
        if (stop_at_count > 0 && (int64_t)count_prev <= stop_at_count) ovr = (int32_t)eos;   // vllm.py:105-108
        else if (finished) ovr = (int32_t)pad;                                               // vllm.py:110-118
        q_sp[q] = finished ? 0 : nsp;
        q_ep[q] = finished ? 0 : nep;
        override_tok[q] = ovr;
    }
}

// =================================================================================================
// K2: distinct-symbol enumeration of BWT[sp,ep) -> vocab bitmap (+counts), fused with the -inf
// masking of the beam's logits row.
//
//  k_mask_top    one warp per range.  Bitmap mode: zeroes the mask row; a range of <= 32 rows is finished here (one
//                lane per row, `levels` access probes + 1 byte load, bit ORed into the row).  Otherwise: descend all
//                nibble levels but the last (LT = levels-1) ONCE and leave the dense table of the depth-LT nodes (each
//                covers 4096 symbols); bitmap mode appends the non-empty nodes (b, node, lo, hi) to a work list.
//  k_mask_expand (bitmap mode) one half-warp per work item: ranks the 16 children at the last nibble level; children
//                of <= 32 rows are finished right there (all of the warp's small children as ONE flat row list), the
//                wider ones become byte-level nodes (b, prefix, lo, hi) of a second list.
//  k_mask_bytes  (bitmap mode) one warp per byte-level node, presence flags in 256 bytes of shared memory, bits ORed
//                into the mask row by L2 atomics.  No block barrier anywhere in the bitmap pipeline: every warp runs
//                on its own, so several batches on different streams interleave freely.
//  k_mask_logits one CTA per (beam, 4096-token slice): issues the slice's logits loads FIRST (registers), then
//                finishes the walk from the depth-LT table inside the CTA (half-warp per node at the last nibble
//                level, warp per byte-level node; its latency overlaps the streaming traffic), ORs the per-shard
//                masks of the other GPUs, applies override / always-allowed tokens, streams out.
// =================================================================================================
constexpr int SLICE = 4096;        // tokens per CTA
constexpr int MASK_THREADS = 256;   // logits kernels (one CTA per slice)
constexpr int MAXF = 288;          // BFS frontier capacity: <= SLICE/16 + 2 nodes at the deepest inner level

template <int DT> struct Elem;
template <> struct Elem<FMIDX_F32> { using T = float; static constexpr int VEC = 4; };
template <> struct Elem<FMIDX_BF16> { using T = __nv_bfloat16; static constexpr int VEC = 8; };
template <> struct Elem<FMIDX_F16> { using T = __half; static constexpr int VEC = 8; };

__device__ __forceinline__ float fm_add_mask(float x, bool allowed) { return x + (allowed ? 0.0f : -INFINITY); }
__device__ __forceinline__ __nv_bfloat16 fm_add_mask(__nv_bfloat16 x, bool allowed) {
    return __float2bfloat16_rn(__bfloat162float(x) + (allowed ? 0.0f : -INFINITY));
}
__device__ __forceinline__ __half fm_add_mask(__half x, bool allowed) {
    return __float2half_rn(__half2float(x) + (allowed ? 0.0f : -INFINITY));
}

// header value hdr[d] + matches of digit d below `off` in a staged cell (8 x uint4 in smem)
__device__ __forceinline__ uint32_t fm_rank_staged(const uint4 *cell, uint32_t d, int off) {
    uint32_t v = reinterpret_cast<const uint32_t *>(cell)[d];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const int nb = off - 32 * g;
        if (nb > 0) v += __popc(fm_match(cell[4 + g], d) & fm_below(nb));
    }
    return v;
}

// same with the lane's plane XOR masks precomputed and branch-free position masks (BFS inner loop)
__device__ __forceinline__ uint32_t fm_rank_fast(const uint4 *cell, uint32_t d, uint32_t off, uint32_t x0, uint32_t x1,
                                                 uint32_t x2, uint32_t x3) {
    const uint64_t mlo = off >= 64u ? ~0ull : ((1ull << off) - 1ull);          // positions 0..63 below off
    const uint64_t mhi = off <= 64u ? 0ull : ((1ull << (off - 64u)) - 1ull);   // positions 64..127 below off
    const uint32_t mg[4] = {(uint32_t)mlo, (uint32_t)(mlo >> 32), (uint32_t)mhi, (uint32_t)(mhi >> 32)};
    uint32_t v = reinterpret_cast<const uint32_t *>(cell)[d];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const uint4 p = cell[4 + g];
        v += __popc((p.x ^ x0) & mg[g] & (p.y ^ x1) & (p.z ^ x2) & (p.w ^ x3));
    }
    return v;
}

// single-thread access+rank at one level: returns the next-level position of the element AT pos
__device__ __forceinline__ uint32_t fm_access1(const uint4 *__restrict__ level_cells, uint32_t pos, uint32_t &dig_out) {
    const uint4 *cell = level_cells + (size_t)(pos >> 7) * FMIDX_CELL_U4;
    const int off = (int)(pos & 127u), g = off >> 5, t = off & 31;
    // the header quarter is only known after the digit: fetch all four speculatively so the level
    // costs one memory round trip (the line is fetched once anyway)
    const uint4 h0 = fm_ldg(cell), h1 = fm_ldg(cell + 1), h2 = fm_ldg(cell + 2), h3 = fm_ldg(cell + 3);
    const uint4 p = fm_ldg(cell + 4 + g);
    const uint32_t dig = ((p.x >> t) & 1u) | (((p.y >> t) & 1u) << 1) | (((p.z >> t) & 1u) << 2) | (((p.w >> t) & 1u) << 3);
    const uint4 hq = (dig >> 2) == 0 ? h0 : ((dig >> 2) == 1 ? h1 : ((dig >> 2) == 2 ? h2 : h3));
    uint32_t v = fm_pick(hq, dig & 3u) + __popc(fm_match(p, dig) & ((1u << t) - 1u));
    for (int g2 = 0; g2 < g; ++g2) v += __popc(fm_match(fm_ldg(cell + 4 + g2), dig));
    dig_out = dig;
    return v;
}

struct EnumSmem {
    uint32_t bm[2 * SLICE / 32];       // 2-byte logits use 8192-token slices
    uint32_t seed[4][3];            // depth-LT nodes of this slice (<= 2 when a top node spans one slice)
    uint32_t bnode[48][3];          // byte-level nodes (each covers 256 symbols): <= 16 per seed (+ overlap)
    uint32_t n_seed, n_bnode, width;
    alignas(1024) int cnt[MASK_THREADS / 32][256];  // per-warp histogram, or 256 presence-flag BYTES, of one byte-level node
    uint4 stage[MASK_THREADS / 16][16];
};

// number of depth-l nodes (l nibble levels consumed) that can hold symbols <= max_sym
__device__ __host__ __forceinline__ uint32_t fm_nodes_at(uint32_t max_sym, uint32_t levels, uint32_t l) {
    return (max_sym >> (8u + 4u * (levels - l))) + 1u;
}

// OR `bits` (<= 32 consecutive symbols starting at base_sym; bits of symbols outside the bitmap are zero) into the bitmap
__device__ __forceinline__ void fm_or_bits(uint32_t *bm, uint32_t bits, uint32_t base_sym, int64_t shift, uint32_t t0) {
    const int64_t r = (int64_t)base_sym - shift - (int64_t)t0;  // bit position of the first symbol
    if (r >= 0) {
        const uint64_t v = (uint64_t)bits << (r & 31);
        atomicOr(&bm[r >> 5], (uint32_t)v);
        if (v >> 32) atomicOr(&bm[(r >> 5) + 1], (uint32_t)(v >> 32));
    } else if (-r < 32) {
        const uint32_t v = bits >> (uint32_t)(-r);
        if (v) atomicOr(&bm[0], v);
    }
}

// ---- presence flags: 256 BYTES per warp (the first 256 bytes of its cnt[] block, 256-aligned) -------------------
// flag[v] = 1 by a plain one-byte store; same-address stores of a warp merge and different bytes of one word do not
// conflict, so a skewed byte distribution costs nothing.  The store address is formed by ONE byte permute (byte e of
// x into the low byte of the 256-aligned shared-space base): two issued instructions per scanned row.
#define FM_FLAG_ST(x, e) asm volatile("st.shared.u8 [%0], %1;" ::"r"(__byte_perm((x), fbase, 0x7650 + (e))), "r"(1) : "memory")
__device__ __forceinline__ void fm_flag4(uint32_t fbase, uint32_t x) {
    FM_FLAG_ST(x, 0);
    FM_FLAG_ST(x, 1);
    FM_FLAG_ST(x, 2);
    FM_FLAG_ST(x, 3);
}
__device__ __forceinline__ void fm_flags_zero(int *cnt, int lane) { reinterpret_cast<uint2 *>(cnt)[lane] = make_uint2(0u, 0u); }
// the lane's 8 flags (symbols 8*lane..8*lane+7) as bits; flag bytes are 0 or 1
__device__ __forceinline__ uint32_t fm_flags_read(const int *cnt, int lane) {
    const uint2 f = reinterpret_cast<const uint2 *>(cnt)[lane];
    return (((f.x * 0x01020408u) >> 24) & 0xFu) | ((((f.y * 0x01020408u) >> 24) & 0xFu) << 4);
}

// Presence flags of a byte-level node wider than 512 rows (so lo and hi lie in different cells A < B): symbol v
// occurs in [lo,hi) iff it occurs in A[a,512), in B[0,b), or in a whole cell between them (header difference
// hdr[B] - hdr[A+1] > 0, read only when such cells exist).  Lane i holds payload bytes 16i..16i+15 of each cell (one
// 16-byte load per cell); bytes outside the scan (in the single boundary lane) are replaced by a byte from inside it
// with a byte permute, so every store is unconditional.  Returns the lane's 8 flags (symbols 8*lane..8*lane+7).
__device__ __forceinline__ uint32_t fm_byte_node_presence_wide(const FmDev &ix, uint32_t lo, uint32_t hi, int *cnt, int lane) {
    const uint32_t cA = lo >> 9, a = lo & 511u, cB = hi >> 9, b = hi & 511u;
    const uint4 *cellA = ix.bcells + (size_t)cA * FMIDX_BCELL_U4, *cellB = ix.bcells + (size_t)cB * FMIDX_BCELL_U4;
    const bool far = cB > cA + 1u;
    const int fa = (int)a - 16 * lane;  // scan A: local bytes [max(fa,0), 16)
    const int tb = (int)b - 16 * lane;  // scan B: local bytes [0, min(tb,16))
    uint4 qa = make_uint4(0, 0, 0, 0), qb = qa, a0 = qa, a1 = qa, b0 = qa, b1 = qa;
    if (fa < 16) qa = fm_ldg(cellA + FMIDX_BHDR_U4 + lane);
    if (tb > 0) qb = fm_ldg(cellB + FMIDX_BHDR_U4 + lane);
    if (far) {
        a0 = fm_ldg(cellA + FMIDX_BCELL_U4 + 2 * lane);
        a1 = fm_ldg(cellA + FMIDX_BCELL_U4 + 2 * lane + 1);
        b0 = fm_ldg(cellB + 2 * lane);
        b1 = fm_ldg(cellB + 2 * lane + 1);
    }
    const uint32_t hb = (b0.x != a0.x ? 1u : 0u) | (b0.y != a0.y ? 2u : 0u) | (b0.z != a0.z ? 4u : 0u) | (b0.w != a0.w ? 8u : 0u) |
                        (b1.x != a1.x ? 16u : 0u) | (b1.y != a1.y ? 32u : 0u) | (b1.z != a1.z ? 64u : 0u) | (b1.w != a1.w ? 128u : 0u);
    if (__all_sync(FULL, hb == 0xFFu)) return 0xFFu;  // every symbol already occurs in the whole cells
    const uint32_t fbase = (uint32_t)__cvta_generic_to_shared(cnt);
    fm_flags_zero(cnt, lane);
    __syncwarp();
    if (fa < 16) {
        // leading max(fa,0) bytes lie before lo: selector nibble i = i (keep) or 7 (= last byte of the lane's 16)
        const uint32_t x[4] = {qa.x, qa.y, qa.z, qa.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int nj = min(max(fa - 4 * j, 0), 4);
            const uint32_t m = (1u << (4 * nj)) - 1u;  // nj <= 4
            fm_flag4(fbase, __byte_perm(x[j], qa.w, (0x3210u & ~m) | (0x7777u & m)));
        }
    }
    if (tb > 0) {
        // bytes from min(tb,16) on lie at or after hi: selector nibble i = i (keep) or 4 (= first byte of the lane's 16)
        const uint32_t x[4] = {qb.x, qb.y, qb.z, qb.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int vj = min(max(tb - 4 * j, 0), 4);
            const uint32_t m = (1u << (4 * vj)) - 1u;
            fm_flag4(fbase, __byte_perm(x[j], qb.x, (0x3210u & m) | (0x4444u & ~m)));
        }
    }
    __syncwarp();
    return hb | fm_flags_read(cnt, lane);
}

// Distinct symbols of one byte-level node (positions [lo,hi) of the last level, all sharing `prefix`), by one warp;
// lane i owns the 8 symbols 8i..8i+7.  Without counts: presence only, three tiers by width (<= 32 rows: one lane per
// row; <= 512 rows: flag bytes from a word-granular scan of the rows; wider: header difference of the whole cells in
// between + scans of the two partial cells).  With counts: a 256-bin signed histogram
//   cnt[v] = hdrB[v] - hdrA[v] - #v in A[0,a) + #v in B[0,b)        (cells A, B of lo and hi; cheaper side of each)
// built with shared-memory atomics.
__device__ __forceinline__ void fm_eval_byte_node(const FmDev &ix, uint32_t lo, uint32_t hi, uint32_t prefix, uint32_t sa,
                                                  uint32_t se, int64_t shift, uint32_t t0, uint32_t *bm, int *cnt,
                                                  uint32_t *__restrict__ counts_row, int lane) {
    const uint32_t base_sym = prefix << 8;
    if (!counts_row) {
        // Presence only (the mask path); three tiers by node width, each with as few issued instructions as possible
        // (the kernel is issue-bound: ~90 K nodes per step, most of them a handful of rows).
        const uint32_t w = hi - lo;
        if (w <= 32u) {  // one lane per row, straight into the bitmap
            if ((uint32_t)lane < w) {
                const uint32_t sy = base_sym | fm_byte_at(ix.bcells, lo + (uint32_t)lane);
                if (sy >= sa && sy < se) {
                    const uint32_t rel = (uint32_t)((int64_t)sy - shift) - t0;
                    atomicOr(&bm[rel >> 5], 1u << (rel & 31u));
                }
            }
            return;
        }
        // lane i owns symbols 8i..8i+7: which of them lie in [sa,se) (the window clips at most the first / last node)
        uint32_t keep = 0xFFu;
        if (base_sym < sa || base_sym + 256u > se) {  // warp-uniform, rare
            const int64_t sb = (int64_t)base_sym + 8 * lane;
            const int64_t c_lo = (int64_t)sa - sb, c_hi = (int64_t)se - sb;
            keep = (c_hi <= 0 || c_lo >= 8) ? 0u : ((c_hi >= 8 ? 0xFFu : ((1u << (int)c_hi) - 1u)) & ~(c_lo <= 0 ? 0u : ((1u << (int)c_lo) - 1u)));
        }
        uint32_t bits = 0;
        if (w <= 512u) {
            // Flags by PLAIN shared-memory stores of 1 (same-address stores of a warp merge; no atomics, so a skewed
            // byte distribution costs nothing extra).  The rows are read as aligned 4-byte words (<= 130 of them, all
            // loads issued first); in the two boundary words the bytes outside [lo,hi) are overwritten with a byte from
            // inside, so the four stores per word are unconditional.
            const uint32_t fbase = (uint32_t)__cvta_generic_to_shared(cnt);
            fm_flags_zero(cnt, lane);
            const uint32_t p0 = lo & ~3u, nwords = (((hi + 3u) & ~3u) - p0) >> 2;
            const char *pay = reinterpret_cast<const char *>(ix.bcells) + FMIDX_BHDR_U4 * 16;
            uint32_t wv[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                wv[j] = 0u;
                if (32u * j < nwords) {  // warp-uniform
                    const uint32_t p = p0 + 4u * (32u * j + (uint32_t)lane);
                    if (32u * j + (uint32_t)lane < nwords)
                        wv[j] = __ldg(reinterpret_cast<const uint32_t *>(pay + (size_t)(p >> 9) * (FMIDX_BCELL_U4 * 16) + (p & 511u)));
                }
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                if (32u * j < nwords && 32u * j + (uint32_t)lane < nwords) {
                    const uint32_t p = p0 + 4u * (32u * j + (uint32_t)lane);
                    uint32_t x = wv[j];
                    if (p < lo || p + 4u > hi) {  // boundary word: valid bytes [e0,e1), never empty; others := byte e0
                        const uint32_t e0 = p < lo ? lo - p : 0u, e1 = min(hi - p, 4u);
                        const uint32_t vm = ((1u << (4u * e1)) - 1u) & ~((1u << (4u * e0)) - 1u);
                        x = __byte_perm(x, x, (0x3210u & vm) | ((e0 * 0x1111u) & ~vm));
                    }
                    fm_flag4(fbase, x);
                }
            }
            __syncwarp();
            bits = fm_flags_read(cnt, lane);
        } else {
            bits = fm_byte_node_presence_wide(ix, lo, hi, cnt, lane);
        }
        bits &= keep;
        // four lanes assemble one 32-symbol word: at most 2 OR operations per 32 symbols reach the bitmap
        uint32_t v = bits << (8 * (lane & 3));
        v |= __shfl_xor_sync(FULL, v, 1);
        v |= __shfl_xor_sync(FULL, v, 2);
        if ((lane & 3) == 0 && v) fm_or_bits(bm, v, base_sym + (uint32_t)(lane >> 2) * 32u, shift, t0);
        __syncwarp();
        return;
    }
    const uint32_t cA = lo >> 9, a = lo & 511u, cB = hi >> 9, b = hi & 511u;
    const uint4 *cellA = ix.bcells + (size_t)cA * FMIDX_BCELL_U4, *cellB = ix.bcells + (size_t)cB * FMIDX_BCELL_U4;
    const bool same = cA == cB;
    // Each end is counted from its CHEAPER side: #v in A[0,a) is either scanned directly (a <= 256) or taken as
    // (hdr of cell A+1 - hdr of cell A) - #v in A[a,512); likewise for B when cell B+1 exists (then cell B has no
    // invalid tail positions).  So
    //   cnt[v] = hB[v] - hA[v] + sA * #v in A[fa,ta) + sB * #v in B[fb,tb)
    // with (hA, fa, ta, sA) = (hdrA, 0, a, -1) or (hdrA+1, a, 512, +1) and (hB, fb, tb, sB) = (hdrB, 0, b, +1) or
    // (hdrB+1, b, 512, -1): at most 256 bytes are scanned per end.  Same cell: one scan of [a,b), no headers.
    const bool a_hi = !same && a > 256u;
    const bool b_hi = !same && b > 256u && cB + 1u < ix.nbblocks;
    const uint32_t fa = same ? a : (a_hi ? a : 0u), ta = same ? b : (a_hi ? 512u : a);
    const uint32_t fb = b_hi ? b : 0u, tb = same ? 0u : (b_hi ? 512u : b);
    const int sign_a = same ? 1 : (a_hi ? 1 : -1), sign_b = b_hi ? -1 : 1;
    const uint4 *hA = a_hi ? cellA + FMIDX_BCELL_U4 : cellA, *hB = b_hi ? cellB + FMIDX_BCELL_U4 : cellB;
    const uint32_t *pwa = reinterpret_cast<const uint32_t *>(cellA + FMIDX_BHDR_U4);
    const uint32_t *pwb = reinterpret_cast<const uint32_t *>(cellB + FMIDX_BHDR_U4);
    // every load of the node is issued before anything is consumed: the two header slices of this lane's
    // 8 symbols and the payload words of both scans (a cell has 128 words -> at most 4 per lane per scan)
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0, b0 = a0, b1 = a0;
    if (!same) {
        a0 = fm_ldg(hA + 2 * lane);
        a1 = fm_ldg(hA + 2 * lane + 1);
        b0 = fm_ldg(hB + 2 * lane);
        b1 = fm_ldg(hB + 2 * lane + 1);
    }
    uint32_t wa[4], wb[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t wi = (fa >> 2) + (uint32_t)lane + 32u * k;
        wa[k] = wi < ((ta + 3u) >> 2) ? __ldg(pwa + wi) : 0u;
        const uint32_t wj = (fb >> 2) + (uint32_t)lane + 32u * k;
        wb[k] = wj < ((tb + 3u) >> 2) ? __ldg(pwb + wj) : 0u;
    }
    {
        int *c = cnt + lane * 8;
        c[0] = (int)(b0.x - a0.x); c[1] = (int)(b0.y - a0.y); c[2] = (int)(b0.z - a0.z); c[3] = (int)(b0.w - a0.w);
        c[4] = (int)(b1.x - a1.x); c[5] = (int)(b1.y - a1.y); c[6] = (int)(b1.z - a1.z); c[7] = (int)(b1.w - a1.w);
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t wi = (fa >> 2) + (uint32_t)lane + 32u * k;
        if (wi < ((ta + 3u) >> 2)) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t p = wi * 4u + (uint32_t)e;
                if (p >= fa && p < ta) atomicAdd(&cnt[(wa[k] >> (8 * e)) & 255u], sign_a);
            }
        }
        const uint32_t wj = (fb >> 2) + (uint32_t)lane + 32u * k;
        if (wj < ((tb + 3u) >> 2)) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t p = wj * 4u + (uint32_t)e;
                if (p >= fb && p < tb) atomicAdd(&cnt[(wb[k] >> (8 * e)) & 255u], sign_b);
            }
        }
    }
    __syncwarp();
    uint32_t bits = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = cnt[lane * 8 + k];
        const uint32_t sy = base_sym | (uint32_t)(lane * 8 + k);
        if (c > 0 && sy >= sa && sy < se) {
            bits |= 1u << k;
            if (counts_row) counts_row[(uint32_t)((int64_t)sy - shift)] = (uint32_t)c;
        }
    }
    if (bits) fm_or_bits(bm, bits, base_sym + (uint32_t)lane * 8u, shift, t0);
    __syncwarp();
}

// ---- resident BWT: the rows of a range are contiguous plain symbols ---------------------------------------------
// Calls f(symbol) for every row of [lo,hi), the work split over `nl` cooperating lanes (this one is lane l): 16-byte
// loads of four consecutive rows per lane, i.e. fully coalesced 512-byte (warp) / 256-byte (half-warp) requests.
// The array is padded so that the last vector may be read past row n-1.
constexpr uint32_t NDIST_PENDING = 0xFFFFFFFFu;  // k_mask_top -> k_popcount_pending: this row's count must come from its bitmap
constexpr uint32_t SCAN_INLINE = 1024;
// logits mode: a whole CTA reads a range of up to this many rows (<= 64 KB, coalesced; the beam's other slice-CTAs hit the L2)
// rather than walking the matrix with ~10 dependent probes, like the bitmap pipeline's scan_max.  (Measured neutral on the
// 256-beam decode step, profiles/r2_decode_probe.log: that kernel's time is set by the logits stream, not by the few wide beams.)
constexpr uint32_t SCAN_CTA = 16384;  // ranges up to this many rows are scanned by the warp / CTA that owns the prefix
template <typename F>
__device__ __forceinline__ void fm_scan_rows(const uint32_t *__restrict__ bwt, uint32_t lo, uint32_t hi, uint32_t l, uint32_t nl, F &&f) {
#pragma unroll 2
    for (uint32_t p = (lo & ~3u) + 4u * l; p < hi; p += 4u * nl) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(bwt + p));
        if (p >= lo) f(v.x);
        if (p + 1u >= lo && p + 1u < hi) f(v.y);
        if (p + 2u >= lo && p + 2u < hi) f(v.z);
        if (p + 3u >= lo && p + 3u < hi) f(v.w);
    }
}

// ---- unigram continuation table ---------------------------------------------------------------------------------
// index of the table entry whose block is exactly [s,e), or -1 (binary search over <= 1024 sorted block starts)
__device__ __forceinline__ int fm_uni_find(const FmDev &ix, uint32_t s, uint32_t e) {
    int lo = 0, hi = (int)ix.uni_n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(ix.uni_sp + mid) < s) lo = mid + 1;
        else hi = mid;
    }
    return (lo < (int)ix.uni_n && __ldg(ix.uni_sp + lo) == s && __ldg(ix.uni_ep + lo) == e) ? lo : -1;
}
// the same by `nl` cooperating lanes (this one is lane l) in ONE round trip: every lane loads whole 16-byte vectors of the
// block starts and compares; returns the lane's own finding (or -1) -- reduce with max over the group
__device__ __forceinline__ int fm_uni_find_part(const FmDev &ix, uint32_t s, uint32_t l, uint32_t nl) {
    int found = -1;
    for (uint32_t i = 4u * l; i < ix.uni_n; i += 4u * nl) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(ix.uni_sp + i));  // the array is padded past uni_n
        if (v.x == s) found = (int)i;
        if (v.y == s && i + 1u < ix.uni_n) found = (int)i + 1;
        if (v.z == s && i + 2u < ix.uni_n) found = (int)i + 2;
        if (v.w == s && i + 3u < ix.uni_n) found = (int)i + 3;
    }
    return found;
}
// word i of the TOKEN-axis bitmap (token = symbol - shift, 0 <= token < vocab) cut out of a symbol-axis table row
__device__ __forceinline__ uint32_t fm_uni_word(const uint32_t *__restrict__ srow, uint32_t src_words, uint32_t i, int64_t shift,
                                                uint32_t vocab) {
    const int64_t b0 = (int64_t)32 * i + shift;  // symbol index of the word's bit 0
    uint32_t v = 0;
    if (b0 >= 0) {
        const uint64_t w = (uint64_t)b0 >> 5;
        const uint32_t r = (uint32_t)b0 & 31u;
        const uint32_t lo = w < src_words ? __ldg(srow + w) : 0u, hi = (r && w + 1 < src_words) ? __ldg(srow + w + 1) : 0u;
        v = r ? (lo >> r) | (hi << (32u - r)) : lo;
    } else if (b0 > -32) {
        v = (src_words ? __ldg(srow) : 0u) << (uint32_t)(-b0);
    }
    const uint32_t t0 = 32u * i;
    if (t0 >= vocab) return 0u;
    if (vocab - t0 < 32u) v &= (1u << (vocab - t0)) - 1u;
    return v;
}

// ---- phase A --------------------------------------------------------------------------------------
// lvl = concatenated dense node tables, level l at offset B * sum_{j<l} P_j, row b = P_l entries.
__global__ void __launch_bounds__(128)
k_mask_top(FmDev ix, const uint64_t *__restrict__ sp, const uint64_t *__restrict__ ep, uint32_t B, uint32_t LT,
           uint2 *__restrict__ lvl, const int32_t *__restrict__ override_tok, int64_t shift, uint32_t vocab,
           uint4 *__restrict__ work_items, uint32_t *__restrict__ work_count, uint32_t *__restrict__ mask_out,
           uint32_t *__restrict__ counts_out, uint32_t *__restrict__ occ_out, uint32_t dense_lo, uint32_t dense_n,
           uint32_t zflags, uint64_t *__restrict__ xcounts, uint32_t scan_max, uint32_t *__restrict__ ndist_out,
           uint32_t *__restrict__ seq_inc, uint32_t *__restrict__ occ_state) {
    // occ_state (nullable; bitmap mode, every row dense, <= 32 slices): LAZY CLEARING.  occ_state[b] says which 4096-token slices
    // of row b may hold bits from the buffer's previous use; only those are zeroed here (a narrow prefix dirties 1-3 of its
    // 32 slices: ~1 KB of stores instead of 16 KB -- the row writes were 3/4 of the step's DRAM traffic), and the row's new
    // occupancy is left for the next use (all ones for the paths that do not know better: table rows, wide scans, walks).
    // sharded step: this kernel -- the first of the step that nobody reads the step number in -- advances the buffer set's
    // step number in device memory; the publication at the tail of k_mask_bytes and the pull kernel read it there, so the
    // whole step has no per-step host argument and can be captured in a CUDA graph
    if (seq_inc && blockIdx.x == 0 && threadIdx.x == 0) *seq_inc += 1u;
    // ndist_out (nullable, bitmap mode): number of distinct tokens of row b when the row is finished HERE (empty range, or
    // a range scanned off the resident BWT / resolved row by row: the bits newly set are counted from the atomics' return
    // values), else NDIST_PENDING -- k_popcount_pending then counts just those rows, instead of re-reading every 16-KB row
    __shared__ uint4 stage[8][16];
    const int lane = threadIdx.x & 31, hw = (threadIdx.x >> 4) & 1, hl = threadIdx.x & 15;
    const int wslot = threadIdx.x >> 4;  // half-warp slot in the CTA
    const uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;  // whole warp
    const uint32_t L = ix.levels;
    if (work_items) {
        // bitmap-only mode, range of <= 32 rows (the common case once a prefix has a few tokens): this
        // warp reads the rows' symbols straight off the matrix (one lane per row, `levels` access probes)
        // and sets the bits in the pre-zeroed mask row; no node table, no work item.
        uint64_t s = sp[b], e = ep[b];
        if (e > ix.n) e = ix.n;
        // Rows [dense_lo, dense_lo+dense_n) must come out as complete bitmaps and are zeroed here (the stores
        // overlap the other warps' walks; no separate memset).  The remaining rows (sharded mode: rows whose final
        // mask another rank assembles) are only ever read slice-wise through their occupancy bits, so only the
        // slices that receive bits are zeroed: a prefix with a few continuations writes ~1 KB instead of 16 KB.
        const uint32_t mw = (vocab + 31u) >> 5;
        uint32_t *row = mask_out + (size_t)b * mw;
        const bool dense = b - dense_lo < dense_n;
        if (xcounts && lane == 0) xcounts[b] = e > s ? e - s : 0ull;  // occurrence count of the prefix on this shard
        if (ndist_out && lane == 0) ndist_out[b] = (s >= e && !override_tok) ? 0u : NDIST_PENDING;
        const bool count_here = ndist_out && !override_tok;
        const bool lazy = occ_state != nullptr;
        const uint32_t nsl_all = (vocab + SLICE - 1) / SLICE;
        const uint32_t all_occ = nsl_all >= 32u ? ~0u : ((1u << nsl_all) - 1u);
        uint32_t old_occ = all_occ;
        if (lazy) {
            old_occ = occ_state[b] & all_occ;
            __syncwarp();
            if (lane == 0) occ_state[b] = s < e ? all_occ : 0u;  // the exact paths below overwrite this
        }
        if (ix.uni_n && !counts_out && e > s && e - s > (uint64_t)SCAN_INLINE) {
            int ui = __reduce_max_sync(FULL, fm_uni_find_part(ix, (uint32_t)s, (uint32_t)lane, 32u));
            if (ui >= 0 && __ldg(ix.uni_ep + ui) != (uint32_t)e) ui = -1;
            const uint32_t nslices = (vocab + SLICE - 1) / SLICE;
            if (ui >= 0 && L > 0 && nslices <= fm_nodes_at(ix.max_sym, L, LT)) {  // the work list has P_top slots per prefix
                // the range is one frequent symbol's whole block: its continuation set is in the table.  One work item per
                // 4096-token slice (a half-warp of k_mask_expand copies 512 B); the row is written once, complete -- no
                // zeroing, no walk.  Every slice is reported as possibly occupied (the words are conservative by contract).
                const uint32_t ow = (nslices + 31u) >> 5;
                uint32_t slot = 0;
                if (lane == 0) slot = atomicAdd(work_count, nslices);
                slot = __shfl_sync(FULL, slot, 0);
                for (uint32_t i = (uint32_t)lane; i < nslices; i += 32) work_items[slot + i] = make_uint4(b, 0xFFFFFFFEu, (uint32_t)ui, i);
                if (occ_out)
                    for (uint32_t i = (uint32_t)lane; i < ow; i += 32)
                        occ_out[(size_t)b * ow + i] = (i + 1u) * 32u <= nslices ? ~0u : ((1u << (nslices & 31u)) - 1u);
                return;
            }
        }
        if (dense && !(zflags & 1u)) {  // bit 0: the rows were zeroed by k_search (fmidx_query_batch)
            if (lazy && old_occ != all_occ) {
                uint32_t todo = old_occ;  // only the slices the previous use may have dirtied
                while (todo) {
                    const uint32_t w0 = ((uint32_t)__ffs(todo) - 1u) * (SLICE / 32), w1 = min(w0 + SLICE / 32, mw);
                    if (zflags & 2u) {
                        const uint32_t i = w0 + 4u * (uint32_t)lane;
                        if (i < w1) *reinterpret_cast<uint4 *>(row + i) = make_uint4(0u, 0u, 0u, 0u);
                    } else {
                        for (uint32_t i = w0 + (uint32_t)lane; i < w1; i += 32) row[i] = 0u;
                    }
                    todo &= todo - 1u;
                }
            } else if (zflags & 2u) {    // bit 1: 16-byte stores are legal
                uint4 *row4 = reinterpret_cast<uint4 *>(row);
                for (uint32_t i = (uint32_t)lane; i < (mw >> 2); i += 32) row4[i] = make_uint4(0u, 0u, 0u, 0u);
            } else {
                for (uint32_t i = (uint32_t)lane; i < mw; i += 32) row[i] = 0u;
            }
        }
        if (occ_out) {  // slice-occupancy words of this row (bit s: slice s may hold set bits)
            const uint32_t ow = ((vocab + SLICE - 1) / SLICE + 31u) >> 5;
            for (uint32_t i = (uint32_t)lane; i < ow; i += 32) occ_out[(size_t)b * ow + i] = 0u;
        }
        __syncwarp();
        if (s >= e) return;
        if (ix.bwt && !counts_out && e - s <= (uint64_t)scan_max) {
            // Resident BWT: the distinct symbols of a range of up to scan_max rows are READ, not searched for -- one
            // coalesced pass over BWT[s,e) instead of a descent per visited node.
            const uint32_t lo = (uint32_t)s, hi = (uint32_t)e, w = hi - lo;
            const uint32_t nslices = (vocab + SLICE - 1) / SLICE, ow = (nslices + 31u) >> 5;
            const int64_t ga64 = shift > 0 ? shift : 0, ge64 = min((int64_t)vocab + shift, (int64_t)ix.max_sym + 1);
            const uint32_t ga = (uint32_t)ga64, ge = ge64 > ga64 ? (uint32_t)ge64 : ga;
            if (w <= SCAN_INLINE) {
                if (!dense || occ_out) {
                    // slice occupancy first (sparse rows: exactly the occupied slices are zeroed before any bit is set)
                    for (uint32_t ws = 0; ws < ow; ++ws) {
                        uint32_t occw = 0;
                        fm_scan_rows(ix.bwt, lo, hi, (uint32_t)lane, 32u, [&](uint32_t c) {
                            if (c >= ga && c < ge) {
                                const uint32_t sl = (uint32_t)((int64_t)c - shift) / SLICE;
                                if ((sl >> 5) == ws) occw |= 1u << (sl & 31u);
                            }
                        });
                        occw = __reduce_or_sync(FULL, occw);
                        if (occ_out && lane == 0) occ_out[(size_t)b * ow + ws] = occw;
                        if (!dense) {
                            uint32_t todo = occw;
                            while (todo) {
                                const uint32_t s_cur = ws * 32u + (uint32_t)__ffs(todo) - 1u;
                                const uint32_t w0 = s_cur * (SLICE / 32), w1 = min(w0 + SLICE / 32, mw);
                                for (uint32_t i = w0 + (uint32_t)lane; i < w1; i += 32) row[i] = 0u;
                                todo &= todo - 1u;
                            }
                        }
                    }
                    __syncwarp();
                }
                uint32_t occ_new = 0;  // slices that receive bits (lazy clearing; tk / SLICE < 32 there)
                if (count_here) {
                    // the atomics' return values tell which bits were new (ATOM: a round trip per row -- only when asked for)
                    uint32_t fresh = 0;
                    fm_scan_rows(ix.bwt, lo, hi, (uint32_t)lane, 32u, [&](uint32_t c) {
                        if (c >= ga && c < ge) {
                            const uint32_t tk = (uint32_t)((int64_t)c - shift);
                            const uint32_t bit = 1u << (tk & 31u);
                            fresh += (atomicOr(&row[tk >> 5], bit) & bit) ? 0u : 1u;
                            occ_new |= 1u << ((tk / SLICE) & 31u);
                        }
                    });
                    fresh = __reduce_add_sync(FULL, fresh);
                    if (lane == 0) ndist_out[b] = fresh;
                } else {
                    fm_scan_rows(ix.bwt, lo, hi, (uint32_t)lane, 32u, [&](uint32_t c) {
                        if (c >= ga && c < ge) {
                            const uint32_t tk = (uint32_t)((int64_t)c - shift);
                            atomicOr(&row[tk >> 5], 1u << (tk & 31u));  // RED: fire and forget
                            occ_new |= 1u << ((tk / SLICE) & 31u);
                        }
                    });
                }
                if (lazy) {
                    occ_new = __reduce_or_sync(FULL, occ_new);
                    if (lane == 0) occ_state[b] = occ_new;
                }
                return;
            }
            // wider: at most P_top (<= 32) chunks of the range go to the work list, one half-warp of k_mask_expand each.
            // The occupancy is not known without reading the rows: every slice is reported as possibly occupied (the
            // words are conservative by contract) and sparse rows are zeroed whole.
            const uint32_t P_top = fm_nodes_at(ix.max_sym, L, LT);
            uint32_t chunk = (w + P_top - 1u) / P_top;
            chunk = max(256u, (chunk + 63u) & ~63u);
            const uint32_t nchunks = (w + chunk - 1u) / chunk;
            uint32_t slot = 0;
            if (lane == 0) slot = atomicAdd(work_count, nchunks);
            slot = __shfl_sync(FULL, slot, 0);
            if ((uint32_t)lane < nchunks)
                work_items[slot + (uint32_t)lane] = make_uint4(b, 0xFFFFFFFFu, lo + (uint32_t)lane * chunk, min(hi, lo + ((uint32_t)lane + 1u) * chunk));
            if (occ_out)
                for (uint32_t i = (uint32_t)lane; i < ow; i += 32)
                    occ_out[(size_t)b * ow + i] = (i + 1u) * 32u <= nslices ? ~0u : ((1u << (nslices & 31u)) - 1u);
            if (!dense)
                for (uint32_t i = (uint32_t)lane; i < mw; i += 32) row[i] = 0u;
            return;
        }
        if (e - s <= 32u) {
            const bool valid = (uint64_t)lane < e - s;
            uint32_t pos = (uint32_t)s + (uint32_t)lane, c = 0;
            if (valid) {
                for (uint32_t l = 0; l < L; ++l) {
                    uint32_t dig;
                    pos = fm_access1(ix.cells + l * ix.level_stride, pos, dig);
                    c = (c << 4) | dig;
                }
                c = (c << 8) | fm_byte_at(ix.bcells, pos);
            }
            const int64_t tk = (int64_t)c - shift;
            const bool inside = valid && tk >= 0 && tk < (int64_t)vocab;
            const uint32_t sl = inside ? (uint32_t)tk / SLICE : 0xFFFFFFFFu;
            if (!dense) {  // zero exactly the slices that are about to receive bits
                uint32_t todo = __ballot_sync(FULL, inside);
                while (todo) {
                    const uint32_t s_cur = __shfl_sync(FULL, sl, __ffs(todo) - 1);
                    const uint32_t w0 = s_cur * (SLICE / 32), w1 = min(w0 + SLICE / 32, mw);
                    for (uint32_t i = w0 + (uint32_t)lane; i < w1; i += 32) row[i] = 0u;
                    todo &= ~__ballot_sync(FULL, sl == s_cur);
                }
                __syncwarp();
            }
            bool fresh1 = false;
            if (inside) {
                const uint32_t bit = 1u << ((uint32_t)tk & 31u);
                if (count_here) fresh1 = !(atomicOr(&row[(uint32_t)tk >> 5], bit) & bit);
                else atomicOr(&row[(uint32_t)tk >> 5], bit);
                if (occ_out) {
                    const uint32_t ow = ((vocab + SLICE - 1) / SLICE + 31u) >> 5;
                    atomicOr(&occ_out[(size_t)b * ow + (sl >> 5)], 1u << (sl & 31u));
                }
            }
            if (counts_out) {
                const uint32_t same = __match_any_sync(FULL, inside ? c : (0xFFFFFF00u + (uint32_t)lane));
                if (inside) counts_out[(size_t)b * vocab + (uint32_t)tk] = (uint32_t)__popc(same);
            }
            if (count_here) {
                const uint32_t nf = (uint32_t)__popc(__ballot_sync(FULL, fresh1));
                if (lane == 0) ndist_out[b] = nf;
            }
            if (lazy) {
                const uint32_t occ_new = __reduce_or_sync(FULL, inside ? 1u << (sl & 31u) : 0u);
                if (lane == 0) occ_state[b] = occ_new;
            }
            return;
        }
    }
    uint2 *cur = lvl + b;
    if (!work_items && ix.bwt && !(override_tok && override_tok[b] >= 0)) {  // logits mode: k_mask_logits reads narrow ranges off the resident BWT itself
        uint64_t s = sp[b], e = ep[b];
        if (e > ix.n) e = ix.n;
        if (s >= e || e - s <= (uint64_t)SCAN_INLINE) return;
        if (ix.uni_n && fm_uni_find(ix, (uint32_t)s, (uint32_t)e) >= 0) return;  // k_mask_logits takes the slice from the table
    }
    if (lane == 0) {
        uint64_t s = sp[b], e = ep[b];
        if (e > ix.n) e = ix.n;                                          // fm_index.cpp:130-132
        const bool go = s < e && !(override_tok && override_tok[b] >= 0);  // fm_index.cpp:127,133
        cur[0] = go ? make_uint2((uint32_t)s, (uint32_t)e) : make_uint2(0u, 0u);
    }
    __syncwarp();
    size_t base = B;
    uint32_t P_cur = 1;
    for (uint32_t l = 0; l < LT; ++l) {
        const uint32_t P_next = fm_nodes_at(ix.max_sym, L, l + 1);
        uint2 *nxt = lvl + base + (size_t)b * P_next;
        const uint4 *lc = ix.cells + l * ix.level_stride;
        for (uint32_t p0 = 0; p0 < P_cur; p0 += 2) {
            const uint32_t p = p0 + hw;
            const bool has = p < P_cur;
            const uint2 node = has ? cur[p] : make_uint2(0u, 0u);
            const bool ne = node.y > node.x;
            const uint32_t pos = hl < 8 ? node.x : node.y;
            if (ne) stage[wslot][hl] = fm_ldg(lc + (size_t)(pos >> 7) * FMIDX_CELL_U4 + (hl & 7));
            __syncwarp();
            uint32_t r_lo = 0, r_hi = 0;
            if (ne) {
                r_lo = fm_rank_staged(&stage[wslot][0], (uint32_t)hl, (int)(node.x & 127u));
                r_hi = fm_rank_staged(&stage[wslot][8], (uint32_t)hl, (int)(node.y & 127u));
            }
            __syncwarp();
            const uint32_t cp = (p << 4) | (uint32_t)hl;
            if (has && cp < P_next) nxt[cp] = (r_hi > r_lo) ? make_uint2(r_lo, r_hi) : make_uint2(0u, 0u);
        }
        __syncwarp();
        cur = nxt;
        P_cur = P_next;
        base += (size_t)B * P_next;
    }
    if (work_items) {
        // bitmap-only mode: one work item per non-empty depth-LT node (b, node, lo, hi) that intersects the token
        // window; k_mask_expand / k_mask_bytes finish the walk with warps that never synchronise with each other
        const uint32_t nslices = (vocab + SLICE - 1) / SLICE;
        const uint32_t span_bits = 8u + 4u * (L - LT);
        const int64_t ga = shift > 0 ? shift : 0;
        const int64_t ge = min((int64_t)vocab + shift, (int64_t)ix.max_sym + 1);
        for (uint32_t p0 = 0; p0 < P_cur; p0 += 32) {
            const uint32_t p = p0 + (uint32_t)lane;
            uint2 nd = make_uint2(0u, 0u);
            bool ne = false;
            if (p < P_cur) {
                nd = cur[p];
                const int64_t na = (int64_t)p << span_bits;
                ne = nd.y > nd.x && na < ge && na + ((int64_t)1 << span_bits) > ga;
            }
            const uint32_t bal = __ballot_sync(FULL, ne);
            uint32_t slot = 0;
            if (lane == 0 && bal) slot = atomicAdd(work_count, (uint32_t)__popc(bal));
            slot = __shfl_sync(FULL, slot, 0);
            if (ne) work_items[slot + __popc(bal & ((1u << lane) - 1u))] = make_uint4(b, p, nd.x, nd.y);
        }
        if (occ_out || !(b - dense_lo < dense_n)) {
            // slice occupancy (bit s: slice s may receive bits) and, for sparse rows, zeroing of exactly those slices
            const uint32_t mw = (vocab + 31u) >> 5;
            uint32_t *row = mask_out + (size_t)b * mw;
            for (uint32_t s0 = 0; s0 < nslices; s0 += 32) {
                const uint32_t s = s0 + lane;
                bool ne = false;
                if (s < nslices) {
                    int64_t a64 = (int64_t)s * SLICE + shift, e64 = (int64_t)min((s + 1) * (uint32_t)SLICE, vocab) + shift;
                    if (a64 < 0) a64 = 0;
                    if (e64 > (int64_t)ix.max_sym + 1) e64 = (int64_t)ix.max_sym + 1;
                    if (a64 < e64)
                        for (uint32_t p = (uint32_t)a64 >> span_bits; p <= ((uint32_t)e64 - 1u) >> span_bits; ++p) {
                            const uint2 nd = cur[p];
                            ne |= nd.y > nd.x;
                        }
                }
                const uint32_t bal = __ballot_sync(FULL, ne);
                if (occ_out && lane == 0) occ_out[(size_t)b * ((nslices + 31u) >> 5) + (s0 >> 5)] = bal;
                if (!(b - dense_lo < dense_n)) {
                    uint32_t todo = bal;
                    while (todo) {
                        const uint32_t s_cur = s0 + (uint32_t)__ffs(todo) - 1u;
                        const uint32_t w0 = s_cur * (SLICE / 32), w1 = min(w0 + SLICE / 32, mw);
                        for (uint32_t i = w0 + (uint32_t)lane; i < w1; i += 32) row[i] = 0u;
                        todo &= todo - 1u;
                    }
                }
            }
        }
    }
}

// ---- the walk inside one slice (whole CTA) ---------------------------------------------------------
// top_b: this range's depth-LT node table (LT = levels-1 nibble levels consumed; a node covers 4096
// symbols; with levels == 0 the single root is itself a byte-level node); [sa,se): the slice's symbol
// interval; bm (zeroed by the caller) receives bit (symbol - shift - t0); counts_row (nullable) is
// indexed by token.
__device__ __forceinline__ void fm_seed_from_table(const FmDev &ix, const uint2 *__restrict__ top_b, uint32_t LT, uint32_t sa,
                                                   uint32_t se, EnumSmem &S) {  // thread 0 only
    const uint32_t span_bits = 8u + 4u * (ix.levels - LT);
    uint32_t n = 0, W = 0;
    if (sa < se)
        for (uint32_t p = sa >> span_bits; p <= (se - 1u) >> span_bits && n < 4; ++p) {
            const uint2 nd = top_b[p];
            if (nd.y > nd.x) {
                S.seed[n][0] = nd.x;
                S.seed[n][1] = nd.y;
                S.seed[n][2] = p;
                ++n;
                W += nd.y - nd.x;
            }
        }
    S.n_seed = n;
    S.n_bnode = 0;
    S.width = W;
}

// walk from the seeds already placed in S (S.seed / n_seed / width / n_bnode = 0, made visible by a barrier)
__device__ __forceinline__ void fm_enum_seeded(const FmDev &ix, uint32_t LT, uint32_t sa, uint32_t se, int64_t shift,
                                               uint32_t t0, EnumSmem &S, uint32_t *__restrict__ counts_row) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t L = ix.levels;
    const uint32_t n_seed = S.n_seed, W = S.width;
    if (n_seed == 0) return;
    if (W <= 32u) {
        // narrow: one lane per BWT position, straight access down the remaining levels
        if (tid < 32) {
            const bool valid = (uint32_t)lane < W;
            uint32_t pos = 0, c = 0, acc = 0;
            bool found = false;
            for (uint32_t i = 0; i < n_seed; ++i) {
                const uint32_t lo = S.seed[i][0], w = S.seed[i][1] - lo;
                if (!found && (uint32_t)lane < acc + w) {
                    pos = lo + ((uint32_t)lane - acc);
                    c = S.seed[i][2];
                    found = true;
                }
                acc += w;
            }
            if (valid) {
                for (uint32_t l = LT; l < L; ++l) {
                    uint32_t dig;
                    pos = fm_access1(ix.cells + l * ix.level_stride, pos, dig);
                    c = (c << 4) | dig;
                }
                c = (c << 8) | fm_byte_at(ix.bcells, pos);
            }
            const bool inside = valid && c >= sa && c < se;
            if (inside) {
                const uint32_t rel = (uint32_t)((int64_t)c - shift) - t0;
                atomicOr(&S.bm[rel >> 5], 1u << (rel & 31u));
            }
            if (counts_row) {
                const uint32_t same = __match_any_sync(FULL, inside ? c : (0xFFFFFF00u + (uint32_t)lane));
                if (inside) counts_row[(uint32_t)((int64_t)c - shift)] = (uint32_t)__popc(same);
            }
        }
        return;
    }
    if (L == 0) {
        // the alphabet fits one byte: the seed IS the byte-level node
        if (wid == 0) fm_eval_byte_node(ix, S.seed[0][0], S.seed[0][1], 0u, sa, se, shift, t0, S.bm, S.cnt[0], counts_row, lane);
        __syncthreads();
        return;
    }
    // stage 1: the last nibble level -- one half-warp per seed, lane d evaluates child d (a byte-level node)
    {
        const int hw = tid >> 4, hl = tid & 15;
        const uint32_t x0 = (hl & 1) ? 0u : ~0u, x1 = (hl & 2) ? 0u : ~0u, x2 = (hl & 4) ? 0u : ~0u, x3 = (hl & 8) ? 0u : ~0u;
        const uint4 *lc = ix.cells + (size_t)LT * ix.level_stride;
        if (hw < 4) {  // warps 0 and 1; n_seed <= 4
            const bool has = (uint32_t)hw < n_seed;
            uint32_t lo = 0, hi = 0, prefix = 0;
            if (has) {
                lo = S.seed[hw][0];
                hi = S.seed[hw][1];
                prefix = S.seed[hw][2];
                S.stage[hw][hl] = fm_ldg(lc + (size_t)((hl < 8 ? lo : hi) >> 7) * FMIDX_CELL_U4 + (hl & 7));
            }
            __syncwarp();
            uint32_t r_lo = 0, r_hi = 0;
            if (has) {
                r_lo = fm_rank_fast(&S.stage[hw][0], (uint32_t)hl, lo & 127u, x0, x1, x2, x3);
                r_hi = fm_rank_fast(&S.stage[hw][8], (uint32_t)hl, hi & 127u, x0, x1, x2, x3);
            }
            const uint32_t cp = (prefix << 4) | (uint32_t)hl;
            const uint64_t c_lo = (uint64_t)cp << 8, c_hi = ((uint64_t)cp + 1) << 8;
            const bool exists = has && r_hi > r_lo && c_lo < se && c_hi > sa;
            const uint32_t bal = (__ballot_sync(FULL, exists) >> (lane & 16)) & 0xffffu;
            uint32_t slot = 0;
            if (hl == 0 && bal) slot = atomicAdd(&S.n_bnode, (uint32_t)__popc(bal));
            slot = __shfl_sync(FULL, slot, lane & 16);
            if (exists) {
                const uint32_t idx = slot + __popc(bal & ((1u << hl) - 1u));
                if (idx < 48) {
                    S.bnode[idx][0] = r_lo;
                    S.bnode[idx][1] = r_hi;
                    S.bnode[idx][2] = cp;
                }
            }
        }
    }
    __syncthreads();
    // stage 2: the byte level -- one warp per node
    const uint32_t nb = min(S.n_bnode, 48u);
    for (uint32_t i = (uint32_t)wid; i < nb; i += (blockDim.x >> 5))
        fm_eval_byte_node(ix, S.bnode[i][0], S.bnode[i][1], S.bnode[i][2], sa, se, shift, t0, S.bm, S.cnt[wid], counts_row, lane);
    __syncthreads();
}

__device__ __forceinline__ void fm_enum_slice(const FmDev &ix, const uint2 *__restrict__ top_b, uint32_t LT, uint32_t sa,
                                              uint32_t se, int64_t shift, uint32_t t0, EnumSmem &S,
                                              uint32_t *__restrict__ counts_row, uint64_t s64, uint64_t e64,
                                              uint32_t slice_tokens = SLICE) {
    if (ix.bwt && !counts_row) {
        // resident BWT: a range of up to SCAN_INLINE rows is read by the CTA itself (every slice-CTA of the beam reads the
        // same few hundred bytes, L2 hits after the first) and filtered to the slice -- no table, no dependent probes
        if (e64 > ix.n) e64 = ix.n;
        if (s64 >= e64) return;
        if (e64 - s64 <= (uint64_t)SCAN_CTA) {
            __syncthreads();  // S.bm was zeroed by all threads
            fm_scan_rows(ix.bwt, (uint32_t)s64, (uint32_t)e64, threadIdx.x, blockDim.x, [&](uint32_t c) {
                if (c >= sa && c < se) {
                    const uint32_t rel = (uint32_t)((int64_t)c - shift) - t0;
                    atomicOr(&S.bm[rel >> 5], 1u << (rel & 31u));
                }
            });
            return;
        }
        if (ix.uni_n) {
            __syncthreads();  // S.bm was zeroed by all threads; S.n_seed is free until the seeds are made
            if (threadIdx.x == 0) S.n_seed = 0xFFFFFFFFu;
            __syncthreads();
            const int mine = fm_uni_find_part(ix, (uint32_t)s64, threadIdx.x, blockDim.x);
            if (mine >= 0) S.n_seed = (uint32_t)mine;  // block starts are distinct: at most one thread writes
            __syncthreads();
            int ui = (int)S.n_seed;
            if (ui >= 0 && __ldg(ix.uni_ep + ui) != (uint32_t)e64) ui = -1;
            __syncthreads();  // everybody has read S.n_seed before the seed code may overwrite it
            if (ui >= 0) {  // one frequent symbol's whole block: the slice comes straight out of the unigram table
                const uint32_t *srow = ix.uni_masks + (size_t)ui * ix.uni_words;
                const uint32_t t1 = (uint32_t)min((int64_t)se - shift, (int64_t)t0 + slice_tokens);  // tokens of this slice: [t0, t1)
                for (uint32_t w = threadIdx.x; w < slice_tokens / 32; w += blockDim.x) S.bm[w] = fm_uni_word(srow, ix.uni_words, (t0 >> 5) + w, shift, t1);
                return;
            }
        }
    }
    if (top_b) {
        if (threadIdx.x == 0) fm_seed_from_table(ix, top_b, LT, sa, se, S);
    } else {
        // No table from k_mask_top (logits mode launches nothing before this kernel): the CTA descends by itself to the
        // depth-LT nodes that intersect its slice -- at most four, one half-warp each (sp octet + ep octet, LT dependent probes;
        // the cells of the first levels are shared by every slice-CTA of the beam, i.e. L2 hits).  Only beams whose range is
        // wider than SCAN_INLINE rows and not a unigram-table entry come here.
        const int tid = threadIdx.x, lane = tid & 31, hw = tid >> 4, hl = tid & 15, sub = tid & 7;
        const uint32_t span_bits = 8u + 4u * (ix.levels - LT);
        if (tid < 64) {
            const uint32_t p = (sa >> span_bits) + (uint32_t)hw;
            const bool need = sa < se && p <= ((se - 1u) >> span_bits);
            uint32_t lo = (uint32_t)s64, hi = (uint32_t)min(e64, (uint64_t)ix.n);
            for (uint32_t l = 0; l < LT; ++l) {
                const uint32_t d = (p >> (4u * (LT - 1u - l))) & 15u;
                const uint32_t r = fm_step8(ix.cells + l * ix.level_stride, hl < 8 ? lo : hi, d, sub, need && hi > lo);
                const uint32_t o = __shfl_xor_sync(FULL, r, 8);
                lo = hl < 8 ? r : o;
                hi = hl < 8 ? o : r;
            }
            if (hl == 0) {
                S.bnode[hw][0] = lo;  // staging (the byte-level node list is not in use yet)
                S.bnode[hw][1] = (need && hi > lo) ? hi : lo;
                S.bnode[hw][2] = p;
            }
            (void)lane;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t n = 0, W = 0;
            for (int k = 0; k < 4; ++k)
                if (S.bnode[k][1] > S.bnode[k][0]) {
                    S.seed[n][0] = S.bnode[k][0];
                    S.seed[n][1] = S.bnode[k][1];
                    S.seed[n][2] = S.bnode[k][2];
                    ++n;
                    W += S.bnode[k][1] - S.bnode[k][0];
                }
            S.n_seed = n;
            S.n_bnode = 0;
            S.width = W;
        }
    }
    __syncthreads();
    fm_enum_seeded(ix, LT, sa, se, shift, t0, S, counts_row);
}

__device__ __forceinline__ void fm_slice_interval(const FmDev &ix, uint32_t t0, uint32_t t1, int64_t shift, uint32_t &sa,
                                                  uint32_t &se) {
    int64_t a64 = (int64_t)t0 + shift, e64 = (int64_t)t1 + shift;
    if (a64 < 0) a64 = 0;
    if (e64 > (int64_t)ix.max_sym + 1) e64 = (int64_t)ix.max_sym + 1;
    sa = (uint32_t)a64;
    se = e64 > a64 ? (uint32_t)e64 : (uint32_t)a64;
}

// ---- phase B, bitmap only ---------------------------------------------------------------------------------
// k_mask_expand: one half-warp per depth-LT work item (LT = levels-1): a node of <= 16 rows is resolved by direct
// access (one lane per row); otherwise lane d ranks child d at the last nibble level and the non-empty children
// -- byte-level nodes (b, prefix, lo, hi) -- are appended to a second list.  If that list is full the warp
// evaluates the child itself (slow, correct).
constexpr int EXPAND_THREADS = 128;
constexpr int BYTES_THREADS = 256;

__global__ void __launch_bounds__(EXPAND_THREADS)
k_mask_expand(FmDev ix, uint32_t LT, int64_t shift, uint32_t vocab, const uint4 *__restrict__ items,
              const uint32_t *__restrict__ n_items_ptr, uint4 *__restrict__ bnodes, uint32_t *__restrict__ n_bnodes,
              uint32_t cap, uint32_t *__restrict__ mask_out, uint32_t *__restrict__ counts_out) {
    __shared__ uint4 stage[EXPAND_THREADS / 16][16];
    __shared__ __align__(1024) int cnt[EXPAND_THREADS / 32][256];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, hw = (tid >> 4) & 1, hl = tid & 15, hslot = tid >> 4;
    const uint32_t n_items = *n_items_ptr;
    const uint32_t mask_words = (vocab + 31u) >> 5;
    const uint32_t ga = (uint32_t)(shift > 0 ? shift : 0);
    const int64_t ge64 = min((int64_t)vocab + shift, (int64_t)ix.max_sym + 1);
    const uint32_t ge = ge64 > (int64_t)ga ? (uint32_t)ge64 : ga;
    const uint32_t x0 = (hl & 1) ? 0u : ~0u, x1 = (hl & 2) ? 0u : ~0u, x2 = (hl & 4) ? 0u : ~0u, x3 = (hl & 8) ? 0u : ~0u;
    const uint4 *lc = ix.cells + (size_t)LT * ix.level_stride;
    const uint32_t gwarp = (blockIdx.x * EXPAND_THREADS + tid) >> 5, nwarps = (gridDim.x * EXPAND_THREADS) >> 5;
    for (uint32_t base = gwarp * 2u; base < n_items; base += nwarps * 2u) {
        const uint32_t i = base + (uint32_t)hw;
        bool has = i < n_items;
        const uint4 it = has ? __ldg(items + i) : make_uint4(0u, 0u, 0u, 0u);
        const uint32_t b = it.x, prefix = it.y, lo = it.z, hi = it.w;
        const bool tab_item = has && prefix == 0xFFFFFFFEu;   // one 4096-token slice of a unigram-table row: lo = entry, hi = slice
        if (__any_sync(FULL, tab_item)) {
            if (tab_item) {
                const uint32_t *srow = ix.uni_masks + (size_t)lo * ix.uni_words;
                uint32_t *drow = mask_out + (size_t)b * mask_words;
                uint32_t v[SLICE / 32 / 16];
#pragma unroll
                for (int k = 0; k < SLICE / 32 / 16; ++k) {
                    const uint32_t i = hi * (SLICE / 32) + 16u * (uint32_t)k + (uint32_t)hl;
                    v[k] = i < mask_words ? fm_uni_word(srow, ix.uni_words, i, shift, vocab) : 0u;
                }
#pragma unroll
                for (int k = 0; k < SLICE / 32 / 16; ++k) {
                    const uint32_t i = hi * (SLICE / 32) + 16u * (uint32_t)k + (uint32_t)hl;
                    if (i < mask_words) drow[i] = v[k];
                }
            }
            __syncwarp();
        }
        has = has && !tab_item;
        const bool scan_item = has && prefix == 0xFFFFFFFFu;  // a chunk of a mid-sized range: read its rows off the resident BWT
        if (__any_sync(FULL, scan_item)) {
            if (scan_item)
                fm_scan_rows(ix.bwt, lo, hi, (uint32_t)hl, 16u, [&](uint32_t c) {
                    if (c >= ga && c < ge) {
                        const uint32_t tk = (uint32_t)((int64_t)c - shift);
                        atomicOr(&mask_out[(size_t)b * mask_words + (tk >> 5)], 1u << (tk & 31u));
                    }
                });
            __syncwarp();
        }
        has = has && !scan_item;
        const bool narrow = has && hi - lo <= 16u;
        if (__any_sync(FULL, narrow)) {
            const bool valid = narrow && (uint32_t)hl < hi - lo;
            uint32_t c = prefix;
            if (valid) {
                uint32_t dig;
                const uint32_t pos = fm_access1(lc, lo + (uint32_t)hl, dig);
                c = (((c << 4) | dig) << 8) | fm_byte_at(ix.bcells, pos);
            }
            const bool inside = valid && c >= ga && c < ge;
            const uint32_t tk = (uint32_t)((int64_t)c - shift);
            if (inside) atomicOr(&mask_out[(size_t)b * mask_words + (tk >> 5)], 1u << (tk & 31u));
            if (counts_out) {
                const unsigned long long key = inside ? (((unsigned long long)hw << 32) | c) : ((1ull << 40) | (unsigned)lane);
                const uint32_t same = __match_any_sync(FULL, key);
                if (inside) counts_out[(size_t)b * vocab + tk] = (uint32_t)__popc(same);
            }
        }
        const bool wide = has && !narrow;
        if (wide) stage[hslot][hl] = fm_ldg(lc + (size_t)((hl < 8 ? lo : hi) >> 7) * FMIDX_CELL_U4 + (hl & 7));
        __syncwarp();
        uint32_t r_lo = 0, r_hi = 0;
        if (wide) {
            r_lo = fm_rank_fast(&stage[hslot][0], (uint32_t)hl, lo & 127u, x0, x1, x2, x3);
            r_hi = fm_rank_fast(&stage[hslot][8], (uint32_t)hl, hi & 127u, x0, x1, x2, x3);
        }
        __syncwarp();
        const uint32_t cp = (prefix << 4) | (uint32_t)hl;
        const uint64_t c_lo = (uint64_t)cp << 8, c_hi = ((uint64_t)cp + 1) << 8;
        const bool exists = wide && r_hi > r_lo && c_lo < ge && c_hi > ga;
        // children of <= 32 rows (the majority) are finished right here, all of the warp's children together: their rows
        // form one flat list (prefix sums over the lanes), 32 rows per pass, owner found by a 5-step search over the
        // sums; only the wider children become byte-level nodes of the second list
        const uint32_t cw = r_hi - r_lo;
        const bool small = exists && !counts_out && cw <= 32u;
        const bool big = exists && !small;
        {
            const uint32_t wsm = small ? cw : 0u;
            uint32_t inc = wsm;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, inc, d);
                if (lane >= d) inc += t;
            }
            const uint32_t total = __shfl_sync(FULL, inc, 31), excl = inc - wsm;
            // four passes (128 rows) per iteration: the owners of all four are found first, then the four byte loads are
            // issued, then the bits are set -- the passes are independent, and a serial load -> atomic chain per pass made this
            // kernel latency-bound (20 us for 1.3 M instructions)
            for (uint32_t f0 = 0; f0 < total; f0 += 128u) {
                uint32_t o_cp[4], o_b[4], addr[4];
                bool live[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t f = f0 + 32u * (uint32_t)u + (uint32_t)lane;
                    uint32_t own = 0;  // first lane whose inclusive sum exceeds f
#pragma unroll
                    for (int step = 16; step >= 1; step >>= 1)
                        if (__shfl_sync(FULL, inc, (int)own + step - 1) <= f) own += (uint32_t)step;
                    own = min(own, 31u);
                    const uint32_t o_excl = __shfl_sync(FULL, excl, (int)own), o_lo = __shfl_sync(FULL, r_lo, (int)own);
                    o_cp[u] = __shfl_sync(FULL, cp, (int)own);
                    o_b[u] = __shfl_sync(FULL, b, (int)own);
                    live[u] = f < total;
                    addr[u] = o_lo + (f - o_excl);
                }
                uint32_t by[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) by[u] = live[u] ? fm_byte_at(ix.bcells, addr[u]) : 0u;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t sy = (o_cp[u] << 8) | by[u];
                    if (live[u] && sy >= ga && sy < ge) {
                        const uint32_t tk = (uint32_t)((int64_t)sy - shift);
                        atomicOr(&mask_out[(size_t)o_b[u] * mask_words + (tk >> 5)], 1u << (tk & 31u));
                    }
                }
            }
        }
        const uint32_t bal = __ballot_sync(FULL, big);
        uint32_t slot = 0;
        if (lane == 0 && bal) slot = atomicAdd(n_bnodes, (uint32_t)__popc(bal));
        slot = __shfl_sync(FULL, slot, 0);
        const uint32_t idx = slot + __popc(bal & ((1u << lane) - 1u));
        if (big && idx < cap) bnodes[idx] = make_uint4(b, cp, r_lo, r_hi);
        uint32_t ovf = __ballot_sync(FULL, big && idx >= cap);
        while (ovf) {  // list full: evaluate the child here
            const int src = __ffs(ovf) - 1;
            const uint32_t ob = __shfl_sync(FULL, b, src), ocp = __shfl_sync(FULL, cp, src);
            const uint32_t olo = __shfl_sync(FULL, r_lo, src), ohi = __shfl_sync(FULL, r_hi, src);
            fm_eval_byte_node(ix, olo, ohi, ocp, ga, ge, shift, 0u, mask_out + (size_t)ob * mask_words, cnt[wid],
                              counts_out ? counts_out + (size_t)ob * vocab : nullptr, lane);
            ovf &= ovf - 1u;
        }
    }
}

// Publication of a shard's side data at the TAIL of its last enumeration kernel (sharded step): the CTA that finishes
// last copies the occupancy words + occurrence counts into slot `rank` of every shard's inbox by peer stores and then
// raises data flag[rank] = seq on every shard (st.release.sys) -- no separate publish launch, and the flag can only be
// seen after every bit of this step's bitmaps has been written (each CTA fences before it takes its ticket).
struct FmPublish {
    uint32_t *const *peer_inbox;   // device array [world]: every shard's inbox base (peer-mapped)
    uint32_t *const *peer_flags;   // device array [world]: every shard's data-flag array
    const uint32_t *xlocal;        // this shard's side data (xwords u32)
    uint32_t *done_ctr;            // ticket counter (zero between launches)
    uint32_t xwords, world, rank, seq;  // world == 0: nothing to publish
    const uint32_t *seq_ptr;       // non-null: the step number lives in device memory (graph-capturable step), seq is ignored
};

__device__ __forceinline__ void fm_publish_tail(const FmPublish &pub) {
    __shared__ uint32_t s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(pub.done_ctr, 1u) == gridDim.x - 1u ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const bool vec = (pub.xwords & 3u) == 0 && (((uintptr_t)pub.xlocal) & 15) == 0;
    for (uint32_t d = 0; d < pub.world; ++d) {
        uint32_t *dst = pub.peer_inbox[d] + (size_t)pub.rank * pub.xwords;
        if (vec && (((uintptr_t)dst) & 15) == 0) {
            for (uint32_t i = threadIdx.x; i < (pub.xwords >> 2); i += blockDim.x)
                reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(pub.xlocal)[i];
        } else {
            for (uint32_t i = threadIdx.x; i < pub.xwords; i += blockDim.x) dst[i] = pub.xlocal[i];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < pub.world) fm_st_release_sys(pub.peer_flags[threadIdx.x] + pub.rank, pub.seq_ptr ? __ldcg(pub.seq_ptr) : pub.seq);
    if (threadIdx.x == 0) *pub.done_ctr = 0u;  // the next launch on this buffer set is stream-ordered after this one
}

// k_mask_bytes: one warp per byte-level node, bits ORed straight into the (pre-zeroed) mask row in global memory.
// Warps are fully independent: no block-level barrier in the node loop; node entries are fetched two ahead.  (Prefetching
// the next node's cells into L1/L2 one node ahead was measured and lost 3%: the kernel is issue-bound, not latency-bound.)
// The grid is persistent with `occ` CTAs per SM chosen by the launcher: the kernel is the only issue-bound one of the
// step, so it must NOT own every register of every SM -- the latency-bound kernels of the other batches in flight
// (k_search, k_mask_top, k_mask_expand) need resident warps next to it to overlap.
template <int MINB>
__global__ void __launch_bounds__(BYTES_THREADS, MINB)
k_mask_bytes(FmDev ix, int64_t shift, uint32_t vocab, const uint4 *__restrict__ bnodes, const uint32_t *__restrict__ n_ptr,
             uint32_t cap, uint32_t *__restrict__ mask_out, uint32_t *__restrict__ counts_out, uint32_t *__restrict__ reset_ctr,
             FmPublish pub) {
    __shared__ __align__(1024) int cnt[BYTES_THREADS / 32][256];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (blockIdx.x == 0 && tid == 0) {  // the counter pair of the NEXT call on this workspace (unused by this one)
        reset_ctr[0] = 0u;
        reset_ctr[1] = 0u;
    }
    const uint32_t n = min(*n_ptr, cap);
    const uint32_t mask_words = (vocab + 31u) >> 5;
    const uint32_t ga = (uint32_t)(shift > 0 ? shift : 0);
    const int64_t ge64 = min((int64_t)vocab + shift, (int64_t)ix.max_sym + 1);
    const uint32_t ge = ge64 > (int64_t)ga ? (uint32_t)ge64 : ga;
    const uint32_t gwarp = (blockIdx.x * BYTES_THREADS + tid) >> 5, nwarps = (gridDim.x * BYTES_THREADS) >> 5;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    uint4 it = gwarp < n ? __ldg(bnodes + gwarp) : zero;
    uint4 n1 = gwarp + nwarps < n ? __ldg(bnodes + gwarp + nwarps) : zero;
    for (uint32_t i = gwarp; i < n; i += nwarps) {
        const uint4 n2 = i + 2u * nwarps < n ? __ldg(bnodes + i + 2u * nwarps) : zero;
        fm_eval_byte_node(ix, it.z, it.w, it.y, ga, ge, shift, 0u, mask_out + (size_t)it.x * mask_words, cnt[wid],
                          counts_out ? counts_out + (size_t)it.x * vocab : nullptr, lane);
        it = n1;
        n1 = n2;
    }
    if (pub.world) fm_publish_tail(pub);
}

// ---- phase B, fused with the logits row: TMA-staged variant ---------------------------------------------
// The slice of the logits row (16 KB fp32 / 8 KB bf16) is brought into shared memory by ONE bulk async
// copy (cp.async.bulk global->shared, completion on an mbarrier; SASS UBLKCP) issued before the tree
// walk, masked in place in shared memory, and written back by ONE bulk async store.  No registers hold
// logits, so ~8 CTAs (128 KB of loads) are in flight per SM while walks are waiting on their probes.
__device__ __forceinline__ uint32_t fm_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int DT>
__global__ void __launch_bounds__(MASK_THREADS)
k_mask_logits_tma(FmDev ix, const uint2 *__restrict__ top, uint32_t P_top, uint32_t LT, int64_t shift, uint32_t vocab,
                  const void *logits_v, void *out_v, int64_t row_stride, const int32_t *__restrict__ override_tok,
                  int32_t always_allow, const uint32_t *__restrict__ extra, uint32_t n_extra, uint32_t B,
                  uint32_t *__restrict__ mask_out, const uint64_t *__restrict__ sp, const uint64_t *__restrict__ ep) {
    using ET = typename Elem<DT>::T;
    constexpr int VEC = Elem<DT>::VEC;
    __shared__ __align__(128) uint4 buf[SLICE / VEC];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ EnumSmem S;
    const int tid = threadIdx.x;
    const uint32_t nslices = (vocab + SLICE - 1) / SLICE;
    const uint32_t b = blockIdx.x / nslices;
    const uint32_t t0 = (blockIdx.x % nslices) * SLICE;
    const uint32_t t1 = min(t0 + (uint32_t)SLICE, vocab);
    const uint32_t mask_words = (vocab + 31u) >> 5;
    const uint32_t bytes = (t1 - t0) * (uint32_t)sizeof(ET);  // multiple of 16 (checked by the launcher)
    const ET *row_in = reinterpret_cast<const ET *>(logits_v) + (size_t)b * row_stride + t0;
    ET *row_out = reinterpret_cast<ET *>(out_v) + (size_t)b * row_stride + t0;
    const uint32_t mb = fm_smem_u32(&mbar), sb = fm_smem_u32(buf);

    // 1. one thread arms the barrier and launches the bulk load of the whole slice
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(sb), "l"(row_in), "r"(bytes), "r"(mb) : "memory");
    }
    // 2. bitmap of the slice (the walk overlaps the copy)
    for (int w = tid; w < SLICE / 32; w += MASK_THREADS) S.bm[w] = 0;
    const int32_t ovr = override_tok ? override_tok[b] : -1;
    uint32_t sa, se;
    fm_slice_interval(ix, t0, t1, shift, sa, se);
    if (ovr < 0) fm_enum_slice(ix, top ? top + (size_t)b * P_top : nullptr, LT, sa, se, shift, t0, S, nullptr, sp[b], ep[b]);
    __syncthreads();
    if (ovr >= 0) {
        if (tid == 0 && (uint32_t)ovr >= t0 && (uint32_t)ovr < t1) S.bm[((uint32_t)ovr - t0) >> 5] = 1u << (((uint32_t)ovr - t0) & 31u);
    } else if (n_extra) {
        const uint32_t nw = (t1 - t0 + 31u) >> 5;
        for (uint32_t w = tid; w < nw; w += MASK_THREADS) {
            uint32_t acc = S.bm[w];
            for (uint32_t x = 0; x < n_extra; ++x) acc |= extra[((size_t)x * B + b) * mask_words + (t0 >> 5) + w];
            S.bm[w] = acc;
        }
    }
    __syncthreads();
    if (tid == 0 && always_allow >= 0 && (uint32_t)always_allow >= t0 && (uint32_t)always_allow < t1)
        S.bm[((uint32_t)always_allow - t0) >> 5] |= 1u << (((uint32_t)always_allow - t0) & 31u);
    __syncthreads();
    if (mask_out) {
        const uint32_t nw = (t1 - t0 + 31u) >> 5;
        for (uint32_t w = tid; w < nw; w += MASK_THREADS) mask_out[(size_t)b * mask_words + (t0 >> 5) + w] = S.bm[w];
    }
    // 3. wait for the slice, mask it in shared memory
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "FM_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
        "@p bra FM_DONE;\n"
        "bra FM_WAIT;\n"
        "FM_DONE:\n"
        "}\n" ::"r"(mb) : "memory");
    const uint32_t nvec = (t1 - t0) / VEC;
    for (uint32_t v = tid; v < nvec; v += MASK_THREADS) {
        const uint32_t rel = v * VEC;
        const uint32_t bits = (S.bm[rel >> 5] >> (rel & 31u)) & ((1u << VEC) - 1u);
        uint4 r = buf[v];
        ET *x = reinterpret_cast<ET *>(&r);
#pragma unroll
        for (int k = 0; k < VEC; ++k) x[k] = fm_add_mask(x[k], (bits >> k) & 1u);
        buf[v] = r;
    }
    // 4. make the generic-proxy writes visible to the async proxy, then one bulk store
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(row_out), "r"(sb), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

// ---- phase B, fused with the logits row: persistent, multi-stage TMA pipeline ---------------------------
// One-shot CTAs keep only ~16 KB of logits in flight per resident CTA for the whole CTA lifetime (issue
// loads -> walk -> stores -> exit -> relaunch), which caps the stream at about half of HBM bandwidth.
// Here PIPE_CTAS_PER_SM persistent CTAs per SM each own a PIPE_STAGES-deep ring of slice buffers in
// shared memory: the bulk loads of the next (beam, slice) items are already in flight while the current
// item is walked, masked in shared memory and bulk-stored.  Items are claimed from a global counter
// (one atomicAdd per item, at load-issue time) so that the few heavy walks do not pile up on one CTA.
constexpr int PIPE_STAGES = 2;
constexpr int PIPE_CTAS_PER_SM = 5;

template <int DT>
__global__ void __launch_bounds__(MASK_THREADS)
k_mask_logits_pipe(FmDev ix, const uint2 *__restrict__ top, uint32_t P_top, uint32_t LT, int64_t shift, uint32_t vocab,
                   const void *logits_v, void *out_v, int64_t row_stride, const int32_t *__restrict__ override_tok,
                   int32_t always_allow, const uint32_t *__restrict__ extra, uint32_t n_extra, uint32_t B,
                   uint32_t *__restrict__ mask_out, uint32_t n_items, uint32_t *__restrict__ next_ptr,
                   const uint64_t *__restrict__ sp, const uint64_t *__restrict__ ep) {
    using ET = typename Elem<DT>::T;
    constexpr int VEC = Elem<DT>::VEC;
    constexpr uint32_t STAGE_BYTES = SLICE * sizeof(ET);
    __shared__ uint32_t q_item[PIPE_STAGES];  // item claimed for each stage (claims are dynamic: heavy walks must not pile up)
    extern __shared__ __align__(128) unsigned char ring[];  // PIPE_STAGES x STAGE_BYTES
    __shared__ __align__(8) uint64_t mbar[PIPE_STAGES];
    __shared__ EnumSmem S;
    const int tid = threadIdx.x;
    const uint32_t nslices = (vocab + SLICE - 1) / SLICE;
    const uint32_t mask_words = (vocab + 31u) >> 5;
    const ET *in_base = reinterpret_cast<const ET *>(logits_v);
    ET *out_base = reinterpret_cast<ET *>(out_v);

    auto issue_load = [&](uint32_t item, int stage) {  // thread 0 only
        const uint32_t b = item / nslices, t0 = (item % nslices) * SLICE, t1 = min(t0 + (uint32_t)SLICE, vocab);
        const uint32_t bytes = (t1 - t0) * (uint32_t)sizeof(ET);
        const uint32_t mb = fm_smem_u32(&mbar[stage]), sb = fm_smem_u32(ring + (size_t)stage * STAGE_BYTES);
        const ET *src = in_base + (size_t)b * row_stride + t0;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(sb), "l"(src), "r"(bytes), "r"(mb) : "memory");
    };

    if (tid == 0) {
        for (int s = 0; s < PIPE_STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fm_smem_u32(&mbar[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < PIPE_STAGES; ++s) {
            const uint32_t item = atomicAdd(next_ptr, 1u);
            q_item[s] = item;
            if (item < n_items) issue_load(item, s);
        }
    }
    for (uint32_t k = 0;; ++k) {
        const int stage = (int)(k % PIPE_STAGES);
        __syncthreads();
        const uint32_t item = q_item[stage];
        if (item >= n_items) break;  // claims are handed out in order: nothing left for this CTA
        const uint32_t parity = (k / PIPE_STAGES) & 1u;
        const uint32_t b = item / nslices, t0 = (item % nslices) * SLICE, t1 = min(t0 + (uint32_t)SLICE, vocab);
        // bitmap of the slice (overlaps the loads in flight)
        for (int w = tid; w < SLICE / 32; w += MASK_THREADS) S.bm[w] = 0;
        const int32_t ovr = override_tok ? override_tok[b] : -1;
        uint32_t sa, se;
        fm_slice_interval(ix, t0, t1, shift, sa, se);
        if (ovr < 0) fm_enum_slice(ix, top ? top + (size_t)b * P_top : nullptr, LT, sa, se, shift, t0, S, nullptr, sp[b], ep[b]);
        __syncthreads();
        if (ovr >= 0) {
            if (tid == 0 && (uint32_t)ovr >= t0 && (uint32_t)ovr < t1) S.bm[((uint32_t)ovr - t0) >> 5] = 1u << (((uint32_t)ovr - t0) & 31u);
        } else if (n_extra) {
            const uint32_t nw = (t1 - t0 + 31u) >> 5;
            for (uint32_t w = tid; w < nw; w += MASK_THREADS) {
                uint32_t acc = S.bm[w];
                for (uint32_t x = 0; x < n_extra; ++x) acc |= extra[((size_t)x * B + b) * mask_words + (t0 >> 5) + w];
                S.bm[w] = acc;
            }
        }
        __syncthreads();
        if (tid == 0 && always_allow >= 0 && (uint32_t)always_allow >= t0 && (uint32_t)always_allow < t1)
            S.bm[((uint32_t)always_allow - t0) >> 5] |= 1u << (((uint32_t)always_allow - t0) & 31u);
        __syncthreads();
        if (mask_out) {
            const uint32_t nw = (t1 - t0 + 31u) >> 5;
            for (uint32_t w = tid; w < nw; w += MASK_THREADS) mask_out[(size_t)b * mask_words + (t0 >> 5) + w] = S.bm[w];
        }
        // wait for this stage's slice, mask it in shared memory
        const uint32_t mb = fm_smem_u32(&mbar[stage]);
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "FM_PWAIT:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra FM_PDONE;\n"
            "bra FM_PWAIT;\n"
            "FM_PDONE:\n"
            "}\n" ::"r"(mb), "r"(parity) : "memory");
        uint4 *buf = reinterpret_cast<uint4 *>(ring + (size_t)stage * STAGE_BYTES);
        const uint32_t nvec = (t1 - t0) / VEC;
        for (uint32_t v = tid; v < nvec; v += MASK_THREADS) {
            const uint32_t rel = v * VEC;
            const uint32_t bits = (S.bm[rel >> 5] >> (rel & 31u)) & ((1u << VEC) - 1u);
            uint4 r = buf[v];
            ET *x = reinterpret_cast<ET *>(&r);
#pragma unroll
            for (int e = 0; e < VEC; ++e) x[e] = fm_add_mask(x[e], (bits >> e) & 1u);
            buf[v] = r;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            const uint32_t bytes = (t1 - t0) * (uint32_t)sizeof(ET);
            ET *dst = out_base + (size_t)b * row_stride + t0;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(fm_smem_u32(buf)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // a stage may be refilled once its store has READ it (not once the write has landed): allow
            // this item's store to stay outstanding and refill the PREVIOUS item's stage
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            if (k > 0) {
                const int ps = (int)((k - 1) % PIPE_STAGES);
                const uint32_t nxt = atomicAdd(next_ptr, 1u);
                q_item[ps] = nxt;
                if (nxt < n_items) issue_load(nxt, ps);
            }
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem must outlive the last store's read
}

// ---- phase B, fused with the logits row: persistent CTAs streaming GROUPS of a row through a bulk-copy ring -----------------
// The one-shot kernels below pay the enumeration's dependent-load chain once per 16 KB of logits and keep only 64 KB of loads in
// flight per SM (4 CTAs x 16 KB, registers), i.e. 0.85 (fp32) / 0.70 (bf16) of the HBM roofline.  Here a work item is a beam and a
// GROUP of 32,768 tokens (128 KB fp32 / 64 KB bf16): the group's bitmap (4 KB of shared memory) is enumerated ONCE, and the
// group's logits stream through a ROW_STAGES-deep ring of 16 KB stages by bulk async copies (cp.async.bulk + mbarrier in,
// cp.async.bulk out): the loads of the next stages -- across item boundaries -- are in flight while a stage is masked in shared
// memory and while the next item's range is enumerated.  Two persistent CTAs per SM, items claimed from a global counter.
constexpr int ROW_STAGES = 5;
constexpr int ROW_CTAS_PER_SM = 2;
constexpr uint32_t ROW_GROUP = 32768;        // tokens per work item
constexpr uint32_t ROW_STAGE_BYTES = 16384;

template <int DT>
__global__ void __launch_bounds__(MASK_THREADS)
k_mask_logits_row(FmDev ix, uint32_t LT, int64_t shift, uint32_t vocab, const void *logits_v, void *out_v, int64_t row_stride,
                  const int32_t *__restrict__ override_tok, int32_t always_allow, const uint32_t *__restrict__ extra, uint32_t n_extra,
                  uint32_t B, uint32_t *__restrict__ mask_out, uint32_t n_items, uint32_t *__restrict__ claim_ctr,
                  const uint64_t *__restrict__ sp, const uint64_t *__restrict__ ep) {
    using ET = typename Elem<DT>::T;
    constexpr int VEC = Elem<DT>::VEC;
    constexpr uint32_t CT = ROW_STAGE_BYTES / sizeof(ET);  // tokens per stage
    extern __shared__ __align__(128) unsigned char ring[];  // ROW_STAGES x ROW_STAGE_BYTES
    __shared__ __align__(8) uint64_t mbar[ROW_STAGES];
    __shared__ uint32_t q_item[ROW_STAGES], q_chunk[ROW_STAGES];
    __shared__ uint32_t gbm[ROW_GROUP / 32];
    __shared__ EnumSmem S;
    const int tid = threadIdx.x;
    const uint32_t ngroups = (vocab + ROW_GROUP - 1) / ROW_GROUP;
    const uint32_t mask_words = (vocab + 31u) >> 5;
    const ET *in_base = reinterpret_cast<const ET *>(logits_v);
    ET *out_base = reinterpret_cast<ET *>(out_v);
    // producer state (thread 0): the item being issued and its next chunk
    uint32_t p_item = 0xFFFFFFFFu, p_chunk = 0, p_nchunks = 0;

    auto item_tokens = [&](uint32_t item, uint32_t &b, uint32_t &t0, uint32_t &t1) {
        b = item / ngroups;
        t0 = (item % ngroups) * ROW_GROUP;
        t1 = min(t0 + ROW_GROUP, vocab);
    };
    auto claim = [&]() {  // thread 0
        p_item = atomicAdd(claim_ctr, 1u);
        p_chunk = 0;
        if (p_item < n_items) {
            uint32_t b, t0, t1;
            item_tokens(p_item, b, t0, t1);
            p_nchunks = (t1 - t0 + CT - 1) / CT;
        } else {
            p_item = 0xFFFFFFFFu;
        }
    };
    auto produce = [&](int stage) {  // thread 0: next chunk of the stream into `stage`
        q_item[stage] = p_item;
        q_chunk[stage] = p_chunk;
        if (p_item == 0xFFFFFFFFu) return;
        uint32_t b, t0, t1;
        item_tokens(p_item, b, t0, t1);
        const uint32_t c0 = t0 + p_chunk * CT, c1 = min(c0 + CT, t1);
        const uint32_t bytes = (c1 - c0) * (uint32_t)sizeof(ET);
        const uint32_t mb = fm_smem_u32(&mbar[stage]), sb = fm_smem_u32(ring + (size_t)stage * ROW_STAGE_BYTES);
        const ET *src = in_base + (size_t)b * row_stride + c0;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(sb), "l"(src), "r"(bytes), "r"(mb) : "memory");
        if (++p_chunk == p_nchunks) claim();
    };

    if (tid == 0) {
        for (int s = 0; s < ROW_STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fm_smem_u32(&mbar[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        claim();
        for (int s = 0; s < ROW_STAGES; ++s) produce(s);
    }
    for (uint32_t k = 0;; ++k) {
        const int stage = (int)(k % ROW_STAGES);
        __syncthreads();
        const uint32_t item = q_item[stage], chunk = q_chunk[stage];
        if (item == 0xFFFFFFFFu) break;  // the stream of this CTA has ended (claims are handed out in order)
        uint32_t b, t0, t1;
        item_tokens(item, b, t0, t1);
        if (chunk == 0) {
            // ---- the group's bitmap, once per item (the ring's loads stay in flight meanwhile) ----
            const uint32_t nw = (t1 - t0 + 31u) >> 5;
            for (uint32_t w = tid; w < ROW_GROUP / 32; w += MASK_THREADS) gbm[w] = 0u;
            const int32_t ovr = override_tok ? override_tok[b] : -1;
            uint64_t s64 = sp[b], e64 = ep[b];
            if (e64 > ix.n) e64 = ix.n;
            uint32_t sa, se;
            fm_slice_interval(ix, t0, t1, shift, sa, se);
            __syncthreads();
            if (ovr >= 0) {
                if (tid == 0 && (uint32_t)ovr >= t0 && (uint32_t)ovr < t1) gbm[((uint32_t)ovr - t0) >> 5] = 1u << (((uint32_t)ovr - t0) & 31u);
            } else if (s64 < e64) {
                bool done = false;
                if (ix.bwt && e64 - s64 <= (uint64_t)SCAN_CTA) {
                    // narrow range: read the rows' symbols off the resident BWT, keep the group's tokens
                    fm_scan_rows(ix.bwt, (uint32_t)s64, (uint32_t)e64, (uint32_t)tid, (uint32_t)MASK_THREADS, [&](uint32_t c) {
                        if (c >= sa && c < se) {
                            const uint32_t rel = (uint32_t)((int64_t)c - shift) - t0;
                            atomicOr(&gbm[rel >> 5], 1u << (rel & 31u));
                        }
                    });
                    done = true;
                } else if (ix.bwt && ix.uni_n) {
                    // one frequent symbol's whole block: the group comes straight out of the unigram table
                    if (tid == 0) S.n_seed = 0xFFFFFFFFu;
                    __syncthreads();
                    const int mine = fm_uni_find_part(ix, (uint32_t)s64, (uint32_t)tid, (uint32_t)MASK_THREADS);
                    if (mine >= 0) S.n_seed = (uint32_t)mine;
                    __syncthreads();
                    int ui = (int)S.n_seed;
                    if (ui >= 0 && __ldg(ix.uni_ep + ui) != (uint32_t)e64) ui = -1;
                    __syncthreads();
                    if (ui >= 0) {
                        const uint32_t *srow = ix.uni_masks + (size_t)ui * ix.uni_words;
                        const uint32_t tl = (uint32_t)min((int64_t)se - shift, (int64_t)t1);
                        for (uint32_t w = tid; w < nw; w += MASK_THREADS) gbm[w] = fm_uni_word(srow, ix.uni_words, (t0 >> 5) + w, shift, tl);
                        done = true;
                    }
                }
                if (!done) {
                    // any other range (wide and not a table entry; sampled layout): the slice walk of the one-shot kernels,
                    // 4096 tokens at a time
                    for (uint32_t st0 = t0; st0 < t1; st0 += SLICE) {
                        const uint32_t st1 = min(st0 + (uint32_t)SLICE, t1);
                        for (int w = tid; w < 2 * SLICE / 32; w += MASK_THREADS) S.bm[w] = 0u;
                        uint32_t ssa, sse;
                        fm_slice_interval(ix, st0, st1, shift, ssa, sse);
                        fm_enum_slice(ix, nullptr, LT, ssa, sse, shift, st0, S, nullptr, s64, e64);
                        __syncthreads();
                        for (uint32_t w = tid; w < ((st1 - st0 + 31u) >> 5); w += MASK_THREADS) gbm[((st0 - t0) >> 5) + w] = S.bm[w];
                        __syncthreads();
                    }
                }
            }
            __syncthreads();
            if (ovr < 0 && n_extra) {
                for (uint32_t w = tid; w < nw; w += MASK_THREADS) {
                    uint32_t acc = gbm[w];
                    for (uint32_t x = 0; x < n_extra; ++x) acc |= extra[((size_t)x * B + b) * mask_words + (t0 >> 5) + w];
                    gbm[w] = acc;
                }
                __syncthreads();
            }
            if (tid == 0 && always_allow >= 0 && (uint32_t)always_allow >= t0 && (uint32_t)always_allow < t1)
                gbm[((uint32_t)always_allow - t0) >> 5] |= 1u << (((uint32_t)always_allow - t0) & 31u);
            __syncthreads();
            if (mask_out)
                for (uint32_t w = tid; w < nw; w += MASK_THREADS) mask_out[(size_t)b * mask_words + (t0 >> 5) + w] = gbm[w];
        }
        // ---- wait for the stage, mask it in shared memory, bulk-store it ----
        const uint32_t parity = (k / ROW_STAGES) & 1u;
        const uint32_t mb = fm_smem_u32(&mbar[stage]);
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "FM_RWAIT:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra FM_RDONE;\n"
            "bra FM_RWAIT;\n"
            "FM_RDONE:\n"
            "}\n" ::"r"(mb), "r"(parity) : "memory");
        const uint32_t c0 = t0 + chunk * CT, c1 = min(c0 + CT, t1);
        uint4 *buf = reinterpret_cast<uint4 *>(ring + (size_t)stage * ROW_STAGE_BYTES);
        const uint32_t nvec = (c1 - c0) / VEC, wbase = (c0 - t0) >> 5;
        for (uint32_t v = tid; v < nvec; v += MASK_THREADS) {
            const uint32_t rel = v * VEC;
            const uint32_t bits = (gbm[wbase + (rel >> 5)] >> (rel & 31u)) & ((1u << VEC) - 1u);
            uint4 r = buf[v];
            ET *x = reinterpret_cast<ET *>(&r);
#pragma unroll
            for (int e = 0; e < VEC; ++e) x[e] = fm_add_mask(x[e], (bits >> e) & 1u);
            buf[v] = r;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            const uint32_t bytes = (c1 - c0) * (uint32_t)sizeof(ET);
            ET *dst = out_base + (size_t)b * row_stride + c0;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(fm_smem_u32(buf)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // a stage may be refilled once its store has READ it: this chunk's store stays outstanding, the previous chunk's
            // stage is refilled
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            if (k > 0) produce((int)((k - 1) % ROW_STAGES));
        }
    }
    if (tid == 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem must outlive the last store's read
        // the CTA that leaves last resets the claim counter for the next launch on this workspace (stream / event ordered)
        __threadfence();
        if (atomicAdd(claim_ctr + 1, 1u) == gridDim.x - 1u) {
            claim_ctr[0] = 0u;
            claim_ctr[1] = 0u;
        }
    }
}

// ---- phase B, fused with the logits row -------------------------------------------------------------------
// DT: FMIDX_F32/BF16/F16.   VECTOR: 16-byte accesses are legal.   SL: tokens per CTA -- 4096 for fp32, 8192 for the 2-byte
// dtypes, so that a CTA moves 16 KB each way whatever the dtype (with 4096-token slices the bf16 step reached 0.54 of the HBM
// roofline against 0.78 for fp32: same fixed latency per CTA, half the bytes).  Registers are capped for 4 resident CTAs per
// SM = 64 KB of loads in flight per SM.  grid = ceil(vocab/SL) * B, slice-fastest.
template <int DT, bool VECTOR, int SL>
__global__ void __launch_bounds__(MASK_THREADS, 4)
k_mask_logits(FmDev ix, const uint2 *__restrict__ top, uint32_t P_top, uint32_t LT, int64_t shift, uint32_t vocab,
              const void *logits_v, void *out_v, int64_t row_stride, const int32_t *__restrict__ override_tok,
              int32_t always_allow, const uint32_t *__restrict__ extra, uint32_t n_extra, uint32_t B,
              uint32_t *__restrict__ mask_out, const uint64_t *__restrict__ sp, const uint64_t *__restrict__ ep) {
    __shared__ EnumSmem S;
    const int tid = threadIdx.x;
    const uint32_t nslices = (vocab + SL - 1) / SL;
    const uint32_t b = blockIdx.x / nslices;
    const uint32_t t0 = (blockIdx.x % nslices) * SL;
    const uint32_t t1 = min(t0 + (uint32_t)SL, vocab);
    const uint32_t mask_words = (vocab + 31u) >> 5;

    // 1. issue the logits loads for this slice
    using ET = typename Elem<DT>::T;
    constexpr int VEC = Elem<DT>::VEC;
    constexpr int NV = SL / (MASK_THREADS * VEC);  // vectors per thread
    uint4 regs[VECTOR ? NV : 1];
    const ET *row_in = reinterpret_cast<const ET *>(logits_v) + (size_t)b * row_stride;
    ET *row_out = reinterpret_cast<ET *>(out_v) + (size_t)b * row_stride;
    if (VECTOR) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const uint32_t e = t0 + (uint32_t)(i * MASK_THREADS + tid) * VEC;
            if (e < t1) regs[i] = __ldcs(reinterpret_cast<const uint4 *>(row_in + e));
        }
    }
    // 2. bitmap of the slice
    for (int w = tid; w < SL / 32; w += MASK_THREADS) S.bm[w] = 0;
    const int32_t ovr = override_tok ? override_tok[b] : -1;
    uint32_t sa, se;
    fm_slice_interval(ix, t0, t1, shift, sa, se);
    if (ovr < 0) fm_enum_slice(ix, top ? top + (size_t)b * P_top : nullptr, LT, sa, se, shift, t0, S, nullptr, sp[b], ep[b], SL);  // (k_mask_top left override rows empty)
    __syncthreads();
    // 3. per-shard masks of the other GPUs, override rows, always-allowed token
    if (ovr >= 0) {
        if (tid == 0 && (uint32_t)ovr >= t0 && (uint32_t)ovr < t1) S.bm[((uint32_t)ovr - t0) >> 5] = 1u << (((uint32_t)ovr - t0) & 31u);
    } else if (n_extra) {
        const uint32_t nw = (t1 - t0 + 31u) >> 5;
        for (uint32_t w = tid; w < nw; w += MASK_THREADS) {
            uint32_t acc = S.bm[w];
            for (uint32_t x = 0; x < n_extra; ++x) acc |= extra[((size_t)x * B + b) * mask_words + (t0 >> 5) + w];
            S.bm[w] = acc;
        }
    }
    __syncthreads();
    if (tid == 0 && always_allow >= 0 && (uint32_t)always_allow >= t0 && (uint32_t)always_allow < t1)
        S.bm[((uint32_t)always_allow - t0) >> 5] |= 1u << (((uint32_t)always_allow - t0) & 31u);
    __syncthreads();
    if (mask_out) {
        const uint32_t nw = (t1 - t0 + 31u) >> 5;
        for (uint32_t w = tid; w < nw; w += MASK_THREADS) mask_out[(size_t)b * mask_words + (t0 >> 5) + w] = S.bm[w];
    }
    // 4. stream the slice of the logits row
    if (VECTOR) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const uint32_t rel = (uint32_t)(i * MASK_THREADS + tid) * VEC;
            const uint32_t e = t0 + rel;
            if (e < t1) {
                const uint32_t bits = (S.bm[rel >> 5] >> (rel & 31u)) & ((1u << VEC) - 1u);
                ET *x = reinterpret_cast<ET *>(&regs[i]);
#pragma unroll
                for (int k = 0; k < VEC; ++k) x[k] = fm_add_mask(x[k], (bits >> k) & 1u);
                __stcs(reinterpret_cast<uint4 *>(row_out + e), regs[i]);
            }
        }
    } else {
        for (uint32_t e = t0 + tid; e < t1; e += MASK_THREADS) {
            const uint32_t rel = e - t0;
            row_out[e] = fm_add_mask(row_in[e], (S.bm[rel >> 5] >> (rel & 31u)) & 1u);
        }
    }
}

// ---- soft boost as a streaming epilogue over bitmaps -------------------------------------------------------------
// document_constrained_generation_causal.py:137-178 / _causal_qwen.py:159-212:  out = scores + mask with, per row,
//   kind 0 (qwen: still thinking)        mask = 0
//   kind 1 (no trailing corpus n-gram)    mask = vu_empty on the corpus unigrams
//   kind 2 (trailing n-gram)              mask = vd on the continuations D of the n-gram, and on the other corpus unigrams
//                                         vu_nonempty, or vu_empty when D is empty (qwen boosts a fresh start more)
//   and, last, mask[eos] = ve unless ve is NaN (the eos rule).
// D is the row's bitmap from the enumeration kernels (override rows: a single bit), n_distinct its popcount, U the
// resident unigram bitmap.  The values arrive already rounded to the dtype of the scores (the reference assigns Python
// floats into a tensor of that dtype); the addition is done in fp32 and rounded once, like torch's.  One CTA per
// (row, 4096-token slice): pure streaming, 2 x V x sizeof(dtype) + 16 KB per row.
template <int DT>
__global__ void __launch_bounds__(MASK_THREADS)
k_boost_apply(const void *logits_v, void *out_v, int64_t row_stride, uint32_t vocab, const uint32_t *__restrict__ dmask,
              const uint32_t *__restrict__ n_distinct, const uint32_t *__restrict__ unigram, const int32_t *__restrict__ kind,
              const float *__restrict__ vd, const float *__restrict__ vu_nonempty, const float *__restrict__ vu_empty,
              const float *__restrict__ ve, int32_t eos) {
    using ET = typename Elem<DT>::T;
    const uint32_t nslices = (vocab + SLICE - 1) / SLICE;
    const uint32_t b = blockIdx.x / nslices;
    const uint32_t t0 = (blockIdx.x % nslices) * SLICE, t1 = min(t0 + (uint32_t)SLICE, vocab);
    const uint32_t mask_words = (vocab + 31u) >> 5;
    const ET *row_in = reinterpret_cast<const ET *>(logits_v) + (size_t)b * row_stride;
    ET *row_out = reinterpret_cast<ET *>(out_v) + (size_t)b * row_stride;
    const int k = kind[b];
    const float fd = vd[b], fu = (k == 2 && n_distinct[b] != 0u) ? vu_nonempty[b] : vu_empty[b], fe = ve[b];
    const bool has_e = !(fe != fe) && k != 0;
    for (uint32_t v = t0 + threadIdx.x; v < t1; v += MASK_THREADS) {
        float m = 0.0f;
        if (k != 0) {
            const bool ub = (__ldg(unigram + (v >> 5)) >> (v & 31u)) & 1u;
            const bool db = k == 2 && ((__ldg(dmask + (size_t)b * mask_words + (v >> 5)) >> (v & 31u)) & 1u);
            m = db ? fd : (ub ? fu : 0.0f);
            if (has_e && (int32_t)v == eos) m = fe;
        }
        float x;
        if (DT == FMIDX_F32) x = reinterpret_cast<const float *>(row_in)[v];
        else if (DT == FMIDX_BF16) x = __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(row_in)[v]);
        else x = __half2float(reinterpret_cast<const __half *>(row_in)[v]);
        const float y = x + m;
        if (DT == FMIDX_F32) reinterpret_cast<float *>(row_out)[v] = y;
        else if (DT == FMIDX_BF16) reinterpret_cast<__nv_bfloat16 *>(row_out)[v] = __float2bfloat16_rn(y);
        else reinterpret_cast<__half *>(row_out)[v] = __float2half_rn(y);
    }
}

// ---- bitmap -> sparse lists --------------------------------------------------------------------------
// OR of the per-shard bitmap pieces gathered from the other GPUs (NCCL has no bitwise-OR reduction)
__global__ void __launch_bounds__(256) k_or_pieces(const uint4 *__restrict__ pieces, uint32_t n_pieces, size_t n_vec,
                                                   uint4 *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n_vec; i += stride) {
        uint4 acc = __ldcs(pieces + i);
        for (uint32_t p = 1; p < n_pieces; ++p) {
            const uint4 v = __ldcs(pieces + (size_t)p * n_vec + i);
            acc.x |= v.x; acc.y |= v.y; acc.z |= v.z; acc.w |= v.w;
        }
        out[i] = acc;
    }
}

// ---- fused second half of a sharded step: wait + pull + OR + total counts + ack in ONE launch ------------------------
// Persistent grid of at most sm_count / 2 CTAs of 128 threads: even with all 16 buffer sets of a context waiting at once,
// CTAs spinning on a flag can hold at most half of the GPU's thread slots, so the enumeration kernels whose tails raise
// those very flags (on this GPU for its peers, on the peers for this GPU) always find room to run -- no launch order or
// scheduling policy is relied upon for progress.  A warp owns rows gw, gw + nwarps, ... of the rows
// this rank owns.  Peers are visited in ring order starting after this rank; a warp polls ONLY the data flag of the
// peer it is about to read (ld.acquire.sys on this GPU's own memory; the peer raised it by a remote store at the tail
// of its last enumeration kernel), then reads that peer's occupancy word(s) from the local inbox and pulls just the
// occupied 4096-token slices (512 B each, one LDG.128 per lane, four slices in flight) straight out of the peer's
// bitmap buffer over NVLink, OR-ing them into this rank's own rows in place.  When every peer has been read the warps
// sum the occurrence counts of all B prefixes.  No acknowledgement: the buffer sets ping-pong (fmidx_shard.cu).
// k_xchg_wait: ONE small CTA, warp p spins (system-scope acquire loads, backing off to ~0.5 us) until peer p has published
// this step.  The waiting is separated from the pulling so that nothing that spins ever holds more than one CTA: the pull
// kernels of all buffer sets used to spin with their whole grids, which capped their size (deadlock otherwise) and left the
// pull itself with too few loads in flight for NVLink's loaded latency (4 GPUs: 220 us of pulling per step).
__global__ void __launch_bounds__(1024)
k_xchg_wait(const uint32_t *__restrict__ data_flags, uint32_t world, uint32_t rank, uint32_t seq_arg, const uint32_t *__restrict__ seq_ptr,
            unsigned long long *__restrict__ trace) {
    const uint32_t seq = seq_ptr ? __ldcg(seq_ptr) : seq_arg;
    const uint32_t p = threadIdx.x >> 5;
    if (p >= world || p == rank || (threadIdx.x & 31)) return;
    unsigned long long w0 = 0, w1 = 0;
    if (trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(w0));
    const long long t0 = clock64();
    unsigned ns = 64;
    while ((int32_t)(fm_ld_acquire_sys(data_flags + p) - seq) < 0) {
        __nanosleep(ns);
        if (ns < 512) ns <<= 1;
        if (clock64() - t0 > 20000000000ll) __trap();
    }
    if (trace) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(w1));
        atomicAdd(trace + 0, w1 - w0);
        atomicAdd(trace + 3, 1ull);
    }
}

// k_xchg_pull: one WARP per (owned row, peer): reads the peer's occupancy word(s) of the row from the local inbox and pulls just
// the occupied 4096-token slices (512 B each) straight out of the peer's bitmap buffer over NVLink -- every octet of lanes one
// slice (4 x LDG.128 per lane), two per octet and round, i.e. 8 slices in flight per warp -- and ORs the non-zero words into this
// rank's own row with L2 atomics (the warps of the other peers work on the same row).  The first CTAs also sum the occurrence
// counts of all B prefixes.  Runs after k_xchg_wait on the same stream; no acknowledgement: the buffer sets ping-pong
// (fmidx_shard.cu).
__global__ void __launch_bounds__(128)
k_xchg_pull(uint32_t *const *__restrict__ peer_masks, const uint32_t *__restrict__ inbox,
            uint32_t xwords, uint32_t count_off, uint32_t world, uint32_t rank, uint32_t B, uint32_t vocab,
            uint64_t *__restrict__ counts_out, unsigned long long *__restrict__ trace) {
    const uint32_t mask_words = (vocab + 31u) >> 5;
    const uint32_t nslices = (vocab + SLICE - 1) / SLICE, ow = (nslices + 31u) >> 5;
    const uint32_t own = B / world, row_lo = rank * own;
    const int lane = threadIdx.x & 31, oct = lane >> 3, l8 = lane & 7;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    uint32_t *mine = peer_masks[rank];
    const bool vec_ok = (mask_words & 3u) == 0;  // rows are 16-byte aligned (checked by the launcher for the bases)
    unsigned long long tr0 = 0;
    if (trace && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr0));
    const uint32_t npairs = own * (world - 1u);
    for (uint32_t pr = gw; pr < npairs; pr += nwarps) {
        const uint32_t r = pr / (world - 1u), k = pr % (world - 1u) + 1u;
        const uint32_t p = rank + k < world ? rank + k : rank + k - world;
        const uint32_t row = row_lo + r;
        const uint32_t *theirs = peer_masks[p] + (size_t)row * mask_words;
        uint32_t *dst_row = mine + (size_t)row * mask_words;
        for (uint32_t ws = 0; ws < ow; ++ws) {
            const uint32_t occ = __ldcg(inbox + (size_t)p * xwords + (size_t)row * ow + ws);
            const uint32_t total = (uint32_t)__popc(occ);
            for (uint32_t t0 = 0; t0 < total; t0 += 8) {
                uint4 v[2][4];
                uint32_t wofs[2];
                bool act[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const uint32_t t = t0 + (uint32_t)u * 4u + (uint32_t)oct;
                    act[u] = t < total;
                    const uint32_t sl = act[u] ? ws * 32u + (uint32_t)__fns(occ, 0u, (int)t + 1) : 0u;
                    wofs[u] = sl * (SLICE / 32) + 4u * (uint32_t)l8;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        v[u][q] = make_uint4(0u, 0u, 0u, 0u);
                        const uint32_t w = wofs[u] + 32u * (uint32_t)q;
                        if (act[u]) {
                            const uint32_t *src = theirs + w;
                            if (vec_ok && w + 4u <= mask_words) {
                                v[u][q] = __ldcs(reinterpret_cast<const uint4 *>(src));
                            } else {
                                if (w < mask_words) v[u][q].x = __ldcs(src);
                                if (w + 1 < mask_words) v[u][q].y = __ldcs(src + 1);
                                if (w + 2 < mask_words) v[u][q].z = __ldcs(src + 2);
                                if (w + 3 < mask_words) v[u][q].w = __ldcs(src + 3);
                            }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint4 x = v[u][q];
                        if (act[u] && (x.x | x.y | x.z | x.w)) {
                            uint32_t *dst = dst_row + wofs[u] + 32u * (uint32_t)q;  // words past the row's end were loaded as 0
                            if (x.x) atomicOr(dst, x.x);
                            if (x.y) atomicOr(dst + 1, x.y);
                            if (x.z) atomicOr(dst + 2, x.z);
                            if (x.w) atomicOr(dst + 3, x.w);
                        }
                    }
            }
        }
    }
    if (trace && threadIdx.x == 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        atomicAdd(trace + 1, t1 - tr0);
        atomicAdd(trace + 2, 1ull);
    }
    if (counts_out) {  // every peer's side data has arrived (k_xchg_wait)
        for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
            uint64_t acc = 0;
            for (uint32_t p = 0; p < world; ++p)
                acc += __ldcg(reinterpret_cast<const uint64_t *>(inbox + (size_t)p * xwords + count_off) + b);
            counts_out[b] = acc;
        }
    }
}

// rows whose distinct-token count k_mask_top could not settle (NDIST_PENDING) are counted from their bitmaps; the others keep
// the count they have (one CTA per row; most exit at once)
__global__ void __launch_bounds__(256) k_popcount_pending(const uint32_t *__restrict__ mask, uint32_t mask_words,
                                                          uint32_t *__restrict__ n_out) {
    __shared__ uint32_t part[8];
    const uint32_t b = blockIdx.x;
    if (n_out[b] != NDIST_PENDING) return;
    uint32_t s = 0;
    const uint4 *row4 = reinterpret_cast<const uint4 *>(mask + (size_t)b * mask_words);
    if ((mask_words & 3u) == 0 && (((uintptr_t)mask) & 15) == 0) {
        for (uint32_t w = threadIdx.x; w < (mask_words >> 2); w += 256) {
            const uint4 v = row4[w];
            s += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
        }
    } else {
        for (uint32_t w = threadIdx.x; w < mask_words; w += 256) s += __popc(mask[(size_t)b * mask_words + w]);
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < 8; ++i) t += part[i];
        n_out[b] = t;
    }
}

__global__ void __launch_bounds__(256) k_popcount(const uint32_t *__restrict__ mask, uint32_t mask_words,
                                                  uint32_t *__restrict__ n_out) {
    __shared__ uint32_t part[8];
    const uint32_t b = blockIdx.x;
    uint32_t s = 0;
    for (uint32_t w = threadIdx.x; w < mask_words; w += 256) s += __popc(mask[(size_t)b * mask_words + w]);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < 8; ++i) t += part[i];
        n_out[b] = t;
    }
}

__global__ void __launch_bounds__(256) k_compact(const uint32_t *__restrict__ mask, const uint32_t *__restrict__ counts,
                                                 uint32_t mask_words, uint32_t vocab,
                                                 const uint64_t *__restrict__ offs, int64_t *__restrict__ tokens,
                                                 uint64_t *__restrict__ cnts) {
    __shared__ uint32_t wsum[8];
    __shared__ uint32_t running;
    const uint32_t b = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    const uint64_t out0 = offs[b];
    for (uint32_t w0 = 0; w0 < mask_words; w0 += 256) {
        const uint32_t w = w0 + threadIdx.x;
        uint32_t word = w < mask_words ? mask[(size_t)b * mask_words + w] : 0u;
        uint32_t c = __popc(word), inc = c;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        uint32_t before = running;
        for (int i = 0; i < wid; ++i) before += wsum[i];
        uint32_t idx = before + inc - c;
        while (word) {
            const int bit = __ffs(word) - 1;
            word &= word - 1;
            const uint32_t tk = w * 32u + (uint32_t)bit;
            tokens[out0 + idx] = (int64_t)tk;
            if (cnts) cnts[out0 + idx] = counts ? (uint64_t)counts[(size_t)b * vocab + tk] : 0ull;
            ++idx;
        }
        __syncthreads();
        if (threadIdx.x == 255) running = before + inc;
        __syncthreads();
    }
}

// =================================================================================================
// K3 / K4: LF walks.  One octet per walker; an LF step is `levels` access+rank probes.
// =================================================================================================
__device__ __forceinline__ uint32_t fm_lf8(const FmDev &ix, uint32_t row, int sub, bool act, uint32_t &sym_out) {
    uint32_t pos = row, c = 0;
    for (uint32_t l = 0; l < ix.levels; ++l) {
        uint32_t dig;
        pos = fm_access_step8(ix.cells + l * ix.level_stride, pos, sub, act, dig);
        c = (c << 4) | dig;
    }
    const uint32_t v = act ? fm_byte_at(ix.bcells, pos) : 0u;  // same address in all 8 lanes
    pos = fm_step8_byte(ix.bcells, pos, v, sub, act);
    c = (c << 8) | v;
    sym_out = c;
    return act ? ix.D[c <= ix.max_sym ? c : 0u] + pos : 0u;
}

__global__ void __launch_bounds__(256) k_locate_full(FmDev ix, const uint64_t *__restrict__ rows, size_t n,
                                                     uint64_t *__restrict__ pos) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t r = rows[i];
    pos[i] = r >= ix.n ? ~0ull : (uint64_t)ix.sa_samples[r];  // fm_index.cpp:170
}

__device__ __forceinline__ uint64_t fm_locate_one(const FmDev &ix, uint64_t r64, bool valid, int sub) {
    bool act = valid && r64 < ix.n;
    uint32_t row = act ? (uint32_t)r64 : 0u, steps = 0;
    uint64_t result = ~0ull;  // fm_index.cpp:170 : (size_type)-1
    while (__any_sync(FULL, act)) {
        if (act) {
            const uint2 m = __ldg(ix.sa_marks + (row >> 5));
            if ((m.y >> (row & 31u)) & 1u) {
                result = (uint64_t)ix.sa_samples[m.x + __popc(m.y & ((1u << (row & 31u)) - 1u))] + steps;
                act = false;
            }
        }
        uint32_t sym;
        const uint32_t nrow = fm_lf8(ix, row, sub, act, sym);
        if (act) {
            row = nrow;
            ++steps;
        }
    }
    return result;
}

__global__ void __launch_bounds__(128) k_locate(FmDev ix, const uint64_t *__restrict__ rows, size_t n,
                                                uint64_t *__restrict__ out) {
    const int sub = threadIdx.x & 7;
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const bool valid = i < n;
    const uint64_t result = fm_locate_one(ix, valid ? rows[i] : 0, valid, sub);
    if (valid && sub == 0) out[i] = result;
}

// scalar FMIndex::locate (fm_index.cpp:168-172)
__global__ void __launch_bounds__(32) k_locate1(FmDev ix, uint64_t row, FmScalarOut *out, uint32_t seq) {
    uint64_t result;
    if (ix.sa_dens == 1) result = row >= ix.n ? ~0ull : (uint64_t)ix.sa_samples[row];
    else result = fm_locate_one(ix, row, threadIdx.x < 8, threadIdx.x & 7);
    if (threadIdx.x == 0) {
        out->v[0] = result;
        fm_scalar_done(out, seq);
    }
}

// scalar FMIndex::distinct / distinct_count (fm_index.cpp:99-144) for ranges of up to DISTINCT1_MAX rows: ONE CTA reads the
// rows' symbols (resident BWT: one coalesced read; otherwise `levels` access probes + a byte load per row), sorts them in
// shared memory (bitonic), run-length encodes and writes the ascending (symbol[, count]) list straight into mapped pinned
// host memory.  Wider ranges report status 1 and the caller takes the bitmap path.
constexpr int DISTINCT1_MAX = 4096, DISTINCT1_THREADS = 512;
__global__ void __launch_bounds__(DISTINCT1_THREADS) k_distinct1(FmDev ix, uint64_t low, uint64_t high, int with_counts, uint32_t cap_u64,
                                                                FmScalarOut *out, uint32_t seq, const uint64_t *__restrict__ d_range) {
    if (d_range) {  // speculative form: the range the preceding k_search1 found (narrow ranges: already enumerated there)
        if ((uint32_t)__ldcg(d_range + 2) == seq) return;
        low = __ldcg(d_range);
        high = __ldcg(d_range + 1);
    }
    __shared__ uint32_t key[DISTINCT1_MAX];
    __shared__ uint32_t head[DISTINCT1_MAX];  // sorted index of the k-th run head
    __shared__ uint32_t wsum[DISTINCT1_THREADS / 32];
    __shared__ uint32_t total_s;
    const int tid = threadIdx.x;
    uint64_t e = high > ix.n ? ix.n : high;  // fm_index.cpp:104-106 / 130-132
    const uint32_t w = e > low ? (uint32_t)min(e - low, (uint64_t)DISTINCT1_MAX + 1) : 0u;
    if (w > DISTINCT1_MAX) {
        if (tid == 0) {
            out->status = 1u;
            out->k = 0u;
            fm_scalar_done(out, seq);
        }
        return;
    }
    uint32_t N = 32;
    while (N < w) N <<= 1;
    for (uint32_t i = tid; i < N; i += DISTINCT1_THREADS) {
        uint32_t c = 0xFFFFFFFFu;  // padding sorts last (symbols are <= max_sym < 2^24)
        if (i < w) {
            uint32_t pos = (uint32_t)low + i;
            if (ix.bwt) c = __ldg(ix.bwt + pos);
            else {
                c = 0;
                for (uint32_t l = 0; l < ix.levels; ++l) {
                    uint32_t dig;
                    pos = fm_access1(ix.cells + l * ix.level_stride, pos, dig);
                    c = (c << 4) | dig;
                }
                c = (c << 8) | fm_byte_at(ix.bcells, pos);
            }
        }
        key[i] = c;
    }
    __syncthreads();
    for (uint32_t k2 = 2; k2 <= N; k2 <<= 1)
        for (uint32_t j = k2 >> 1; j > 0; j >>= 1) {
            for (uint32_t i = tid; i < N; i += DISTINCT1_THREADS) {
                const uint32_t ixj = i ^ j;
                if (ixj > i) {
                    const uint32_t a = key[i], b = key[ixj];
                    const bool up = (i & k2) == 0;
                    if ((a > b) == up) { key[i] = b; key[ixj] = a; }
                }
            }
            __syncthreads();
        }
    // run heads -> compact positions (block-wide exclusive scan over N flags, DISTINCT1_MAX / DISTINCT1_THREADS per thread)
    constexpr int PER = DISTINCT1_MAX / DISTINCT1_THREADS;
    uint32_t flags = 0, cnt = 0;
#pragma unroll
    for (int r = 0; r < PER; ++r) {
        const uint32_t i = (uint32_t)tid * PER + r;
        const bool hd = i < w && (i == 0 || key[i] != key[i - 1]);
        flags |= (hd ? 1u : 0u) << r;
        cnt += hd;
    }
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(FULL, incl, o);
        if ((tid & 31) >= o) incl += v;
    }
    if ((tid & 31) == 31) wsum[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        uint32_t v = tid < DISTINCT1_THREADS / 32 ? wsum[tid] : 0u, in2 = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(FULL, in2, o);
            if (tid >= o) in2 += u;
        }
        if (tid < DISTINCT1_THREADS / 32) wsum[tid] = in2 - v;
        if (tid == 31) total_s = in2;
    }
    __syncthreads();
    uint32_t posn = wsum[tid >> 5] + incl - cnt;
#pragma unroll
    for (int r = 0; r < PER; ++r)
        if ((flags >> r) & 1u) head[posn++] = (uint32_t)tid * PER + r;
    __syncthreads();
    const uint32_t nd = total_s;
    const uint32_t per = with_counts ? 2u : 1u;
    if (nd * per > cap_u64) {
        if (tid == 0) {
            out->status = 1u;
            out->k = nd * per;
            fm_scalar_done(out, seq);
        }
        return;
    }
    for (uint32_t k = tid; k < nd; k += DISTINCT1_THREADS) {
        const uint32_t a = head[k], b = k + 1 < nd ? head[k + 1] : w;
        out->data[k * per] = key[a];
        if (with_counts) out->data[k * per + 1] = b - a;
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        out->status = 0u;
        out->k = nd * per;
        fm_scalar_done(out, seq);
    }
}

// isa[p] from text-order samples, then emit BWT[row], row = LF(row), ...
__global__ void __launch_bounds__(128) k_extract(FmDev ix, const uint64_t *__restrict__ begins,
                                                 const uint64_t *__restrict__ ends, size_t n, uint64_t max_len,
                                                 uint64_t *__restrict__ out, uint64_t *__restrict__ lens) {
    const int sub = threadIdx.x & 7;
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const bool valid = i < n;
    uint64_t begin = 0, end = 0;
    if (valid) {
        begin = begins[i];
        end = ends[i];
        if (end > ix.n) end = ix.n;  // fm_index.cpp:178
    }
    uint64_t m = (valid && end > begin) ? end - begin : 0;  // fm_index.cpp:177
    if (m > max_len) m = max_len;
    uint32_t row = 0, walk = 0;
    if (m) {
        const uint32_t p = (uint32_t)(end - 1);
        uint64_t q = ((uint64_t)p + ix.isa_dens - 1) / ix.isa_dens * ix.isa_dens;
        if (q >= ix.n) {
            q = (uint64_t)ix.n - 1;  // ISA[n-1] = 0
            row = 0;
        } else {
            row = ix.isa_samples[q / ix.isa_dens];
        }
        walk = (uint32_t)(q - p);
    }
    uint64_t total = (uint64_t)walk + m, k = 0;
    while (__any_sync(FULL, k < total)) {
        const bool act = k < total;
        uint32_t sym;
        const uint32_t nrow = fm_lf8(ix, row, sub, act, sym);
        if (act) {
            if (k >= walk && sub == 0) out[i * max_len + (k - walk)] = sym;
            row = nrow;
            ++k;
        }
    }
    if (valid && sub == 0) lens[i] = m;
}

// resident-text fast path: extract_text(begin,end)[k] = T[(end-2-k) mod n]
__global__ void __launch_bounds__(256) k_extract_text(FmDev ix, const uint64_t *__restrict__ begins,
                                                      const uint64_t *__restrict__ ends, size_t n, uint64_t max_len,
                                                      uint64_t *__restrict__ out, uint64_t *__restrict__ lens) {
    const size_t i = blockIdx.x;
    uint64_t begin = begins[i], end = ends[i];
    if (end > ix.n) end = ix.n;
    uint64_t m = end > begin ? end - begin : 0;
    if (m > max_len) m = max_len;
    for (uint64_t k = threadIdx.x; k < m; k += blockDim.x) {
        const uint64_t p = (end + (uint64_t)ix.n - 2 - k) % ix.n;
        out[i * max_len + k] = ix.text[p];
    }
    if (threadIdx.x == 0) lens[i] = m;
}

}  // namespace

// =================================================================================================
// launchers
// =================================================================================================
static int check_ready(fmidx *h) {
    if (!h || !h->ready) {
        fm_set_error("index is not initialised");
        return FMIDX_ESTATE;
    }
    return FMIDX_OK;
}

int fm_launch_search(fmidx *h, const int32_t *d_toks, const int64_t *d_offs, size_t B, int64_t shift, uint64_t *d_sp,
                     uint64_t *d_ep, int raw, cudaStream_t s, uint32_t *d_zero_rows, uint32_t zero_words, size_t zero_lo,
                     size_t zero_n) {
    FM_TRY(check_ready(h));
    if (B == 0) return FMIDX_OK;
    if (B > 0x0FFFFFFFull) { fm_set_error("batch too large"); return FMIDX_EINVAL; }
    const unsigned grid = (unsigned)((B * 16 + 127) / 128);
    if (d_zero_rows && ((zero_words & 3u) || ((uintptr_t)d_zero_rows & 15))) d_zero_rows = nullptr;  // caller falls back to k_mask_top
    const uint32_t zlo = (uint32_t)zero_lo, zn = (uint32_t)std::min<size_t>(zero_n, 0xFFFFFFFFu);
    if (raw) FM_LAUNCH(k_search<true>, grid, 128, 0, s, h->dev, d_toks, d_offs, (uint32_t)B, shift, d_sp, d_ep, d_zero_rows, zero_words, zlo, zn);
    else FM_LAUNCH(k_search<false>, grid, 128, 0, s, h->dev, d_toks, d_offs, (uint32_t)B, shift, d_sp, d_ep, d_zero_rows, zero_words, zlo, zn);
    FM_CUDA(cudaGetLastError());
    return FMIDX_OK;
}

// true when fm_launch_search can zero [B][words] mask rows at d_mask on behalf of fm_launch_mask
bool fm_search_can_zero(const uint32_t *d_mask, uint32_t vocab) {
    // Off unless $FMIDX_K1_ZERO=1.  Measured on B200 (profiles/README.md, round 2): the 65 MB of row stores queue up in front
    // of the search's dependent-load chain (lg_throttle) and the kernel cannot retire before they have drained: k_search
    // 30 -> 38 us, while k_mask_top only saves 3 us -- the rows are zeroed by k_mask_top with 16-byte stores instead.
    static const bool on = [] {
        const char *e = getenv("FMIDX_K1_ZERO");
        return e && *e == '1';
    }();
    return on && d_mask && ((((vocab + 31u) >> 5) & 3u) == 0) && (((uintptr_t)d_mask & 15) == 0);
}

int fm_launch_step(fmidx *h, const uint64_t *d_sym, const uint64_t *d_lo, const uint64_t *d_hi, size_t B,
                   uint64_t *d_sp, uint64_t *d_ep, cudaStream_t s) {
    FM_TRY(check_ready(h));
    if (B == 0) return FMIDX_OK;
    if (B > 0x0FFFFFFFull) { fm_set_error("batch too large"); return FMIDX_EINVAL; }
    FM_LAUNCH(k_step, (unsigned)((B * 16 + 127) / 128), 128, 0, s, h->dev, d_sym, d_lo, d_hi, (uint32_t)B, d_sp, d_ep);
    FM_CUDA(cudaGetLastError());
    return FMIDX_OK;
}

int fm_launch_decode_step(fmidx *h, const int64_t *d_last_tok, int64_t tok_stride, const int32_t *d_parent, const uint64_t *d_sp_in,
                          const uint64_t *d_ep_in, uint64_t *d_sp_out, uint64_t *d_ep_out, size_t B, int64_t shift, int64_t eos,
                          int64_t pad, int64_t stop_at_count, uint64_t *d_qsp, uint64_t *d_qep, int32_t *d_override, cudaStream_t s) {
    FM_TRY(check_ready(h));
    if (B == 0) return FMIDX_OK;
    if (B > 0x0FFFFFFFull) { fm_set_error("batch too large"); return FMIDX_EINVAL; }
    FM_LAUNCH(k_decode_step, (unsigned)((B * 16 + 127) / 128), 128, 0, s, h->dev, d_last_tok, tok_stride, d_parent, d_sp_in, d_ep_in,
              d_sp_out, d_ep_out, (uint32_t)B, shift, eos, pad, stop_at_count, d_qsp, d_qep, d_override);
    FM_CUDA(cudaGetLastError());
    return FMIDX_OK;
}

int fm_launch_mask(fmidx *h, const uint64_t *d_sp, const uint64_t *d_ep, size_t B, int64_t shift, uint32_t vocab,
                   const void *d_logits, void *d_out, int dtype, int64_t row_stride, const int32_t *d_override,
                   int32_t always_allow, const uint32_t *d_extra, uint32_t n_extra, uint32_t *d_mask,
                   uint32_t *d_counts, cudaStream_t s, uint32_t *d_occ, size_t dense_lo, size_t dense_n, bool prezeroed,
                   const FmXchgArgs *xa, uint32_t *d_ndist, uint32_t *d_occ_state) {
    FM_TRY(check_ready(h));
    if (B == 0 || vocab == 0) return FMIDX_OK;
    if (!d_occ) { dense_lo = 0; dense_n = B; }  // sparse rows are only meaningful together with occupancy words
    if (d_occ_state && (d_logits || d_occ || d_counts || d_override || prezeroed || (vocab + SLICE - 1) / SLICE > 32u)) {
        // lazy clearing does not cover this call: whoever owns the state must treat the rows as unknown afterwards
        FM_CUDA(cudaMemsetAsync(d_occ_state, 0xFF, B * sizeof(uint32_t), s));
        d_occ_state = nullptr;
    }
    const uint32_t nslices = (vocab + SLICE - 1) / SLICE;
    const size_t nblk = (size_t)nslices * B;
    if (nblk > 0x7FFFFFFFull) { fm_set_error("mask batch too large: split the call"); return FMIDX_EINVAL; }
    const bool has = d_logits != nullptr;
    if (has && !d_out) { fm_set_error("d_out is NULL"); return FMIDX_EINVAL; }
    if (!has && !d_mask) { fm_set_error("d_mask is NULL"); return FMIDX_EINVAL; }
    // workspace: dense node tables of levels 0..LT, then the work list and its two counters
    FmDev ix = h->dev;
    if (!h->tune_unigram) ix.uni_n = 0;
    const uint32_t L = ix.levels, LT = L > 0 ? L - 1 : 0;  // nibble levels consumed by k_mask_top
    size_t n_nodes = 0, top_off = 0;
    for (uint32_t l = 0; l <= LT; ++l) {
        if (l == LT) top_off = n_nodes;
        n_nodes += (size_t)B * fm_nodes_at(ix.max_sym, L, l);
    }
    const uint32_t P_top = fm_nodes_at(ix.max_sym, L, LT);
    const size_t lvl_bytes = n_nodes * sizeof(uint2);
    // work lists (bitmap mode): depth-LT items (<= P_top per range) and byte-level nodes (<= P_leaf per range, capped:
    // k_mask_expand evaluates the overflow itself); FMIDX_BNODE_CAP overrides the cap (tests)
    const uint32_t P_leaf = fm_nodes_at(ix.max_sym, L, L);
    static const size_t cap_max = [] {
        const char *e = getenv("FMIDX_BNODE_CAP");
        return e && *e ? (size_t)strtoull(e, nullptr, 10) : ((size_t)1 << 22);
    }();
    if ((size_t)B * P_top > 0x7FFFFFFFull) { fm_set_error("mask batch too large: split the call"); return FMIDX_EINVAL; }
    const size_t cap = std::min<size_t>((size_t)B * P_leaf, cap_max);
    const size_t items_bytes = has ? 0 : (size_t)B * P_top * 16;
    const size_t bnode_bytes = has || L == 0 ? 0 : cap * 16;
    // Workspace: a pipelined slot's own (the caller guarantees stream order per slot), else the host-style calls'
    // (private stream h->stream) and the device-style calls (caller's stream) each have their own, and a workspace
    // that is used from a different stream than last time first waits for its previous user (event), so neither the
    // node tables nor the alternating counter pairs are ever shared by two calls in flight.
    FmBuf &ws = h->mask_ws_override ? *h->mask_ws_override : (s == h->stream ? h->mask_ws_host : h->mask_ws);
    const bool guarded = !h->mask_ws_override;
    if (guarded && ws.last_use && ws.last_stream != s) FM_CUDA(cudaStreamWaitEvent(s, ws.last_use, 0));
    if (ws.p && 128 + lvl_bytes + items_bytes + bnode_bytes + 128 > ws.cap) {
        // growing frees the old buffer: wait for whoever still runs on it
        if (ws.last_use) FM_CUDA(cudaEventSynchronize(ws.last_use));
        else FM_CUDA(cudaStreamSynchronize(s));
    }
    FM_TRY(fm_scratch(h, ws, 128 + lvl_bytes + items_bytes + bnode_bytes + 128, false));
    struct Stamp {  // records the workspace's last use on every exit path
        FmBuf &ws; cudaStream_t s; bool on;
        ~Stamp() {
            if (!on) return;
            if (!ws.last_use && cudaEventCreateWithFlags(&ws.last_use, cudaEventDisableTiming) != cudaSuccess) { ws.last_use = nullptr; return; }
            cudaEventRecord(ws.last_use, s);
            ws.last_stream = s;
        }
    } stamp{ws, s, guarded};
    // first 128 bytes: two pairs of work-list counters ([0]=#items [1]=#byte nodes / next), used alternately: the last
    // kernel of a call zeroes the pair of the following call, so steady-state calls need no memset
    // (logits mode uses neither pair: it leaves them and their epoch alone)
    if (!has && !ws.ctr_clean) {
        FM_CUDA(cudaMemsetAsync(ws.p, 0, 32, s));
        ws.epoch = 0;
    }
    if (!has) ws.ctr_clean = false;  // set again once every launch of this call went through
    uint32_t *ctr = (uint32_t *)ws.p + 2 * (ws.epoch & 1u), *ctr_next = (uint32_t *)ws.p + 2 * ((ws.epoch & 1u) ^ 1u);
    if (!has) ++ws.epoch;
    uint2 *lvl = (uint2 *)((char *)ws.p + 128);
    const size_t items_off = 128 + ((lvl_bytes + 15) & ~(size_t)15);
    uint4 *items = has ? nullptr : (uint4 *)((char *)ws.p + items_off);
    uint4 *bnodes = (uint4 *)((char *)ws.p + items_off + items_bytes);
    const uint32_t scan_max = std::max<uint32_t>(h->tune_scan_max, SCAN_INLINE);  // rows read off the resident BWT instead of walking the matrix
    const uint32_t mwords = (vocab + 31u) >> 5;
    const uint32_t zflags = (prezeroed ? 1u : 0u) | ((!has && (mwords & 3u) == 0 && ((uintptr_t)d_mask & 15) == 0) ? 2u : 0u);
    // logits mode launches no k_mask_top: the slice-CTAs of k_mask_logits read narrow ranges off the resident BWT, take frequent
    // one-token prefixes from the unigram table and descend to their own depth-LT nodes otherwise (fm_enum_slice)
    if (!has)
    FM_LAUNCH(k_mask_top, (unsigned)((B * 32 + 127) / 128), 128, 0, s, ix, d_sp, d_ep, (uint32_t)B, LT, lvl, d_override, shift,
              vocab, items, ctr, has ? nullptr : d_mask, has ? nullptr : d_counts, has ? nullptr : d_occ, (uint32_t)dense_lo,
              (uint32_t)std::min<size_t>(dense_n, 0xFFFFFFFFu), zflags, xa && !has ? xa->xcounts : nullptr,
              L == 0 ? SCAN_INLINE : scan_max, d_occ ? nullptr : d_ndist, xa ? xa->seq_dev : nullptr, d_occ_state);
    if (!has) {
        // CTAs per SM of the persistent k_mask_bytes grid (fmidx_tune "bytes_occ", 1..6)
        const int occ = h->tune_bytes_occ < 1 ? 1 : (h->tune_bytes_occ > 6 ? 6 : h->tune_bytes_occ);
        const unsigned bgrid = (unsigned)h->sm_count * (unsigned)occ;
        FmPublish pub{};
        if (xa && xa->world) {
            pub.peer_inbox = xa->peer_inbox; pub.peer_flags = xa->peer_data_flags; pub.xlocal = xa->xlocal; pub.done_ctr = xa->done_ctr;
            pub.xwords = (uint32_t)xa->xwords; pub.world = xa->world; pub.rank = xa->rank; pub.seq = xa->seq; pub.seq_ptr = xa->seq_dev;
        }
#define FM_BYTES(LIST, CNT, CAP)                                                                                             \
    do {                                                                                                                     \
        if (occ >= 6) FM_LAUNCH(k_mask_bytes<6>, bgrid, BYTES_THREADS, 0, s, ix, shift, vocab, LIST, CNT, CAP, d_mask, d_counts, ctr_next, pub); \
        else if (occ >= 4) FM_LAUNCH(k_mask_bytes<5>, bgrid, BYTES_THREADS, 0, s, ix, shift, vocab, LIST, CNT, CAP, d_mask, d_counts, ctr_next, pub); \
        else FM_LAUNCH(k_mask_bytes<3>, bgrid, BYTES_THREADS, 0, s, ix, shift, vocab, LIST, CNT, CAP, d_mask, d_counts, ctr_next, pub); \
    } while (0)
        if (L == 0) {  // the root is itself a byte-level node: the item list is the node list
            FM_BYTES(items, ctr, (uint32_t)(B * P_top));
        } else {
            const unsigned egrid = (unsigned)std::min<size_t>(((size_t)B * P_top + 7) / 8, (size_t)h->sm_count * 16);
            FM_LAUNCH(k_mask_expand, egrid, EXPAND_THREADS, 0, s, ix, LT, shift, vocab, items, ctr, bnodes, ctr + 1, (uint32_t)cap,
                      d_mask, d_counts);
            FM_BYTES(bnodes, ctr + 1, (uint32_t)cap);
        }
#undef FM_BYTES
        if (d_ndist) {
            // distinct-token counts: settled by k_mask_top for the rows it finished itself, from the bitmap for the others
            if (d_occ) FM_LAUNCH(k_popcount, (unsigned)B, 256, 0, s, d_mask, mwords, d_ndist);
            else FM_LAUNCH(k_popcount_pending, (unsigned)B, 256, 0, s, d_mask, mwords, d_ndist);
        }
        FM_CUDA(cudaGetLastError());
        ws.ctr_clean = true;
        return FMIDX_OK;
    }
    auto aligned = [&](int esz) {
        const int vec = 16 / esz;
        return ((uintptr_t)d_logits % 16 == 0) && ((uintptr_t)d_out % 16 == 0) && (row_stride % vec == 0) &&
               (vocab % vec == 0);
    };
    const unsigned grid = (unsigned)nblk;
    const unsigned grid2 = (unsigned)((size_t)((vocab + 2 * SLICE - 1) / (2 * SLICE)) * B);  // 8192-token slices (2-byte dtypes)
#define FM_MASK_CASE(DT, VECT)                                                                                     \
    do {                                                                                                           \
        if (VECT && DT != FMIDX_F32)                                                                               \
            FM_LAUNCH((k_mask_logits<DT, VECT, 2 * SLICE>), grid2, MASK_THREADS, 0, s, ix, (const uint2 *)nullptr, P_top, LT, shift, vocab, \
                      d_logits, d_out, row_stride, d_override, always_allow, d_extra, n_extra, (uint32_t)B, d_mask, d_sp, d_ep);  \
        else                                                                                                       \
            FM_LAUNCH((k_mask_logits<DT, VECT, SLICE>), grid, MASK_THREADS, 0, s, ix, (const uint2 *)nullptr, P_top, LT, shift, vocab, \
                      d_logits, d_out, row_stride, d_override, always_allow, d_extra, n_extra, (uint32_t)B, d_mask, d_sp, d_ep);  \
    } while (0)
#define FM_MASK_TMA(DT)                                                                                            \
    FM_LAUNCH((k_mask_logits_tma<DT>), grid, MASK_THREADS, 0, s, ix, (const uint2 *)nullptr, P_top, LT, shift, vocab, d_logits,   \
              d_out, row_stride, d_override, always_allow, d_extra, n_extra, (uint32_t)B, d_mask, d_sp, d_ep)
    // FMIDX_LOGITS_PATH = row (above) | regs (default: one-shot CTAs, logits staged in registers -- the fastest measured,
    // 0.83 of HBM bandwidth on narrow beams) | tma (one-shot CTAs, one bulk async copy each way) | pipe
    // (persistent CTAs, multi-stage bulk-copy ring).  The TMA variants are kept selectable: they match the
    // register variant on narrow beams and lose ~15% when a few beams need long walks (fewer resident CTAs
    // to overlap them); see profiles/README.md.
    const char *path = getenv("FMIDX_LOGITS_PATH");
    const bool tma = path && strcmp(path, "tma") == 0;
    const bool pipe = path && strcmp(path, "pipe") == 0;
    const bool rowp = path && strcmp(path, "row") == 0;
    const int esz = dtype == FMIDX_F32 ? 4 : 2;
    if (rowp && (dtype == FMIDX_F32 || dtype == FMIDX_BF16 || dtype == FMIDX_F16) && aligned(esz)) {
        // persistent CTAs, one work item = a beam's group of 32,768 tokens (k_mask_logits_row).  Measured SLOWER than the one-shot
        // register variant (profiles/README.md, round 2: 48 vs 44 us fp32 on narrow beams -- every chunk pays two block barriers
        // and thread 0's store hand-off -- and 2x slower when some beams need slice walks, which run one after the other inside
        // the item instead of side by side in 32 CTAs): kept selectable, not the default.
        const size_t n_items = (size_t)((vocab + ROW_GROUP - 1) / ROW_GROUP) * B;
        const unsigned rgrid = (unsigned)std::min<size_t>(n_items, (size_t)h->sm_count * ROW_CTAS_PER_SM);
        const int dyn = ROW_STAGES * (int)ROW_STAGE_BYTES;
        uint32_t *claim = (uint32_t *)ws.p + 8;  // words 8, 9 of the workspace header: claim counter, exit counter (self-resetting)
        if (!ws.row_clean) {
            FM_CUDA(cudaMemsetAsync(claim, 0, 8, s));
            ws.row_clean = true;
        }
#define FM_MASK_ROW(DT)                                                                                             \
    do {                                                                                                            \
        static bool attr_set = false;                                                                               \
        if (!attr_set) {                                                                                            \
            FM_CUDA(cudaFuncSetAttribute(k_mask_logits_row<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn)); \
            attr_set = true;                                                                                        \
        }                                                                                                           \
        FM_LAUNCH((k_mask_logits_row<DT>), rgrid, MASK_THREADS, dyn, s, ix, LT, shift, vocab, d_logits, d_out, row_stride, d_override, \
                  always_allow, d_extra, n_extra, (uint32_t)B, d_mask, (uint32_t)n_items, claim, d_sp, d_ep);        \
    } while (0)
        if (dtype == FMIDX_F32) FM_MASK_ROW(FMIDX_F32);
        else if (dtype == FMIDX_BF16) FM_MASK_ROW(FMIDX_BF16);
        else FM_MASK_ROW(FMIDX_F16);
#undef FM_MASK_ROW
        FM_CUDA(cudaGetLastError());
        return FMIDX_OK;
    }
    if (pipe && (dtype == FMIDX_F32 || dtype == FMIDX_BF16 || dtype == FMIDX_F16) && aligned(esz)) {
        const int dyn = PIPE_STAGES * SLICE * esz;
        const unsigned pgrid = (unsigned)std::min<size_t>(nblk, (size_t)h->sm_count * PIPE_CTAS_PER_SM);
        FM_CUDA(cudaMemsetAsync(ctr, 0, 8, s));
        ws.ctr_clean = false;  // the pipe variant borrows (and dirties) a work-list counter pair
#define FM_MASK_PIPE(DT)                                                                                            \
    do {                                                                                                            \
        static bool attr_set = false;                                                                               \
        if (!attr_set) {                                                                                            \
            FM_CUDA(cudaFuncSetAttribute(k_mask_logits_pipe<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn)); \
            attr_set = true;                                                                                        \
        }                                                                                                           \
        FM_LAUNCH((k_mask_logits_pipe<DT>), pgrid, MASK_THREADS, dyn, s, ix, (const uint2 *)nullptr, P_top, LT, shift, vocab,    \
                  d_logits, d_out, row_stride, d_override, always_allow, d_extra, n_extra, (uint32_t)B, d_mask,     \
                  (uint32_t)nblk, ctr + 1, d_sp, d_ep);                                                                    \
    } while (0)
        if (dtype == FMIDX_F32) FM_MASK_PIPE(FMIDX_F32);
        else if (dtype == FMIDX_BF16) FM_MASK_PIPE(FMIDX_BF16);
        else FM_MASK_PIPE(FMIDX_F16);
#undef FM_MASK_PIPE
        FM_CUDA(cudaGetLastError());
        return FMIDX_OK;
    }
    switch (dtype) {
        case FMIDX_F32:
            if (aligned(4) && tma) FM_MASK_TMA(FMIDX_F32);
            else if (aligned(4)) FM_MASK_CASE(FMIDX_F32, true);
            else FM_MASK_CASE(FMIDX_F32, false);
            break;
        case FMIDX_BF16:
            if (aligned(2) && tma) FM_MASK_TMA(FMIDX_BF16);
            else if (aligned(2)) FM_MASK_CASE(FMIDX_BF16, true);
            else FM_MASK_CASE(FMIDX_BF16, false);
            break;
        case FMIDX_F16:
            if (aligned(2) && tma) FM_MASK_TMA(FMIDX_F16);
            else if (aligned(2)) FM_MASK_CASE(FMIDX_F16, true);
            else FM_MASK_CASE(FMIDX_F16, false);
            break;
        default: fm_set_error("unknown logits dtype %d", dtype); return FMIDX_EINVAL;
    }
#undef FM_MASK_TMA
#undef FM_MASK_CASE
    FM_CUDA(cudaGetLastError());
    return FMIDX_OK;
}

int fm_launch_boost_apply(fmidx *h, const void *d_logits, void *d_out, int dtype, int64_t row_stride, size_t B, uint32_t vocab,
                          const uint32_t *d_dmask, const uint32_t *d_n_distinct, const uint32_t *d_unigram, const int32_t *d_kind,
                          const float *d_vd, const float *d_vu_nonempty, const float *d_vu_empty, const float *d_ve, int32_t eos,
                          cudaStream_t s) {
    (void)h;
    if (B == 0 || vocab == 0) return FMIDX_OK;
    const size_t nblk = (size_t)((vocab + SLICE - 1) / SLICE) * B;
    if (nblk > 0x7FFFFFFFull) { fm_set_error("boost batch too large: split the call"); return FMIDX_EINVAL; }
#define FM_BOOST(DT) FM_LAUNCH((k_boost_apply<DT>), (unsigned)nblk, MASK_THREADS, 0, s, d_logits, d_out, row_stride, vocab, d_dmask, d_n_distinct, \
                               d_unigram, d_kind, d_vd, d_vu_nonempty, d_vu_empty, d_ve, eos)
    if (dtype == FMIDX_F32) FM_BOOST(FMIDX_F32);
    else if (dtype == FMIDX_BF16) FM_BOOST(FMIDX_BF16);
    else if (dtype == FMIDX_F16) FM_BOOST(FMIDX_F16);
    else { fm_set_error("unknown logits dtype %d", dtype); return FMIDX_EINVAL; }
#undef FM_BOOST
    FM_CUDA(cudaGetLastError());
    return FMIDX_OK;
}

int fm_launch_or(fmidx *h, const uint32_t *d_pieces, uint32_t n_pieces, size_t n_words, uint32_t *d_out, cudaStream_t s) {
    if (n_words == 0 || n_pieces == 0) return FMIDX_OK;
    if ((n_words & 3) || ((uintptr_t)d_pieces & 15) || ((uintptr_t)d_out & 15)) {
        fm_set_error("mask_or: word count must be a multiple of 4 and the buffers 16-byte aligned");
        return FMIDX_EINVAL;
    }
    const size_t n_vec = n_words / 4;
    const size_t cap = (size_t)h->sm_count * 8;
    const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>((n_vec + 255) / 256, cap));
    FM_LAUNCH(k_or_pieces, grid, 256, 0, s, (const uint4 *)d_pieces, n_pieces, n_vec, (uint4 *)d_out);
    FM_CUDA(cudaGetLastError());
    return FMIDX_OK;
}

int fm_launch_xchg_pull(fmidx *h, uint32_t *const *d_peer_masks, const uint32_t *d_data_flags, const uint32_t *d_inbox, size_t xwords,
                        size_t count_off, uint32_t world, uint32_t rank, size_t B, uint32_t vocab, uint32_t seq, uint64_t *d_counts_out,
                        cudaStream_t s, unsigned long long *d_trace, uint32_t n_sets, const uint32_t *d_seq) {
    (void)h; (void)n_sets;
    if (B == 0 || world == 0) return FMIDX_OK;
    if (world > 32) { fm_set_error("the peer-memory exchange supports up to 32 shards"); return FMIDX_EINVAL; }
    if (world > 1) FM_LAUNCH(k_xchg_wait, 1, 32 * world, 0, s, d_data_flags, world, rank, seq, d_seq, d_trace);
    const size_t own = B / world;
    const size_t warps = std::max<size_t>(own * (world > 1 ? world - 1 : 1), (B + 31) / 32);
    // $FMIDX_PULL_CTAS caps the grid (measurement); default: one warp per (owned row, peer)
    static const long pull_ctas = [] { const char *e = getenv("FMIDX_PULL_CTAS"); return e && *e ? atol(e) : 0l; }();
    size_t ctas = (warps + 3) / 4;
    if (pull_ctas > 0) ctas = std::min<size_t>(ctas, (size_t)pull_ctas);
    FM_LAUNCH(k_xchg_pull, (unsigned)std::max<size_t>(1, ctas), 128, 0, s, d_peer_masks, d_inbox, (uint32_t)xwords, (uint32_t)count_off, world,
              rank, (uint32_t)B, vocab, d_counts_out, d_trace);
    FM_CUDA(cudaGetLastError());
    return FMIDX_OK;
}

// ---- random-gather probe (the roofline the rank probes live under) ---------------------------------
// Independent pseudo-random loads over ALL of the index's large arrays taken as one address space (nibble cells,
// byte cells and -- resident layout -- SA, ISA, text: 18 B per token, >= 14 x L2 at 100 M tokens): each lane reads its
// own random 32-B unit (UNIT = 32), or the 8 lanes of an octet read one random aligned 128-B line (UNIT = 128, the
// shape of a nibble-cell probe).  Addresses never depend on loaded data, so this is the memory system's random-access
// throughput, not its latency.
struct GatherBufs {
    const uint4 *base[5];
    unsigned long long cum[5];  // inclusive prefix sums of the buffers' unit counts
    int n;
};

template <int UNIT>
__global__ void __launch_bounds__(256) k_gather_probe(GatherBufs g, uint32_t loads_per_thread, uint32_t *__restrict__ sink) {
    const uint32_t tid = blockIdx.x * 256u + threadIdx.x;
    unsigned long long x = (unsigned long long)(UNIT == 128 ? (tid >> 3) : tid) * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    const unsigned long long total = g.cum[g.n - 1];
    uint32_t acc = 0;
#pragma unroll 4
    for (uint32_t i = 0; i < loads_per_thread; ++i) {
        x = x * 6364136223846793005ull + 1442695040888963407ull;
        unsigned long long u = __umul64hi(x, total);
        int j = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k + 1 < g.n && u >= g.cum[k]) j = k + 1;
        if (j) u -= g.cum[j - 1];
        const uint4 *b = g.base[j];
        if (UNIT == 128) {
            acc ^= fm_ldg(b + u * 8ull + (tid & 7u)).x;
        } else {
            const uint4 v0 = fm_ldg(b + u * 2ull), v1 = fm_ldg(b + u * 2ull + 1ull);
            acc ^= v0.x ^ v1.w;
        }
    }
    if (acc == 0x5EED5EEDu) *sink = acc;  // keeps the loads alive
}

int fm_launch_gather_probe(fmidx *h, int unit_bytes, size_t n_loads, double *out_gbps, cudaStream_t s) {
    FM_TRY(check_ready(h));
    if ((unit_bytes != 32 && unit_bytes != 128) || !out_gbps || n_loads == 0) { fm_set_error("gather_peak: unit must be 32 or 128"); return FMIDX_EINVAL; }
    const FmDev &ix = h->dev;
    GatherBufs g{};
    auto add = [&](const void *p, unsigned long long bytes) {
        if (!p || ((uintptr_t)p & 127u) || bytes < 128 || g.n >= 5) return;
        g.base[g.n] = (const uint4 *)p;
        g.cum[g.n] = (g.n ? g.cum[g.n - 1] : 0ull) + bytes / (unsigned)unit_bytes;
        ++g.n;
    };
    add(ix.cells, (unsigned long long)ix.levels * ix.level_stride * 16ull);
    add(ix.bcells, (unsigned long long)ix.nbblocks * (FMIDX_BCELL_U4 * 16));
    if (ix.sa_dens == 1) add(ix.sa_samples, (unsigned long long)ix.n * 4ull);
    if (ix.isa_dens == 1) add(ix.isa_samples, (unsigned long long)ix.n * 4ull);
    add(ix.text, (unsigned long long)ix.n * 4ull);
    if (g.n == 0) { fm_set_error("index too small"); return FMIDX_ESTATE; }
    const unsigned grid = (unsigned)h->sm_count * 16u;
    const size_t lanes_per_load = unit_bytes == 128 ? 8 : 1, threads = (size_t)grid * 256;
    const uint32_t per_thread = (uint32_t)std::max<size_t>(1, n_loads * lanes_per_load / threads);
    FM_TRY(fm_scratch(h, h->d_scratch[0], 64, false));
    cudaEvent_t e0, e1;
    FM_CUDA(cudaEventCreate(&e0));
    FM_CUDA(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; ++rep) {  // first launch warms up
        FM_CUDA(cudaEventRecord(e0, s));
        if (unit_bytes == 128) FM_LAUNCH(k_gather_probe<128>, grid, 256, 0, s, g, per_thread, (uint32_t *)h->d_scratch[0].p);
        else FM_LAUNCH(k_gather_probe<32>, grid, 256, 0, s, g, per_thread, (uint32_t *)h->d_scratch[0].p);
        FM_CUDA(cudaEventRecord(e1, s));
    }
    FM_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    FM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double loads = (double)per_thread * (double)threads / (double)lanes_per_load;
    *out_gbps = loads * unit_bytes / (ms * 1e-3) / 1e9;
    return FMIDX_OK;
}

int fm_launch_popcount(fmidx *h, const uint32_t *d_mask, size_t B, uint32_t vocab, uint32_t *d_n, cudaStream_t s) {
    (void)h;
    if (B == 0) return FMIDX_OK;
    FM_LAUNCH(k_popcount, (unsigned)B, 256, 0, s, d_mask, (vocab + 31u) >> 5, d_n);
    FM_CUDA(cudaGetLastError());
    return FMIDX_OK;
}

int fm_launch_compact(fmidx *h, const uint32_t *d_mask, const uint32_t *d_counts, size_t B, uint32_t vocab,
                      const uint64_t *d_offs, int64_t *d_tokens, uint64_t *d_cnts, cudaStream_t s) {
    (void)h;
    if (B == 0) return FMIDX_OK;
    FM_LAUNCH(k_compact, (unsigned)B, 256, 0, s, d_mask, d_counts, (vocab + 31u) >> 5, vocab, d_offs, d_tokens, d_cnts);
    FM_CUDA(cudaGetLastError());
    return FMIDX_OK;
}

// ---- scalar calls: by-value arguments, results in mapped pinned host memory, host polls the sequence flag ------------------
static int fm_scalar_prepare(fmidx *h, FmScalarOut **out, uint32_t *seq) {
    FM_TRY(check_ready(h));
    if (!h->scalar_out) {
        // two result blocks (the call's own + the speculative enumeration's) in one pinned allocation, and 16 bytes of device memory
        // through which k_search1 hands its range to the speculative k_distinct1
        const size_t block = (sizeof(FmScalarOut) + (size_t)2 * DISTINCT1_MAX * 8 + 255) & ~(size_t)255;
        FM_CUDA(cudaHostAlloc(&h->scalar_out, 2 * block, cudaHostAllocMapped | cudaHostAllocPortable));
        memset(h->scalar_out, 0, 2 * block);
        h->scalar_spec_out = (char *)h->scalar_out + block;
        h->scalar_cap_u64 = (size_t)2 * DISTINCT1_MAX;
        FM_CUDA(cudaMalloc(&h->scalar_d_range, 64));
        FM_CUDA(cudaMemset(h->scalar_d_range, 0, 64));
    }
    *out = (FmScalarOut *)h->scalar_out;
    *seq = ++h->scalar_seq;
    if (*seq == 0) *seq = ++h->scalar_seq;  // 0 is the initial flag value
    return FMIDX_OK;
}

static int fm_scalar_wait(fmidx *h, FmScalarOut *out, uint32_t seq) {
    FM_CUDA(cudaGetLastError());
    volatile uint32_t *flag = &out->flag;
    for (uint32_t spin = 1;; ++spin) {
        if (*flag == seq) break;
        if ((spin & 0xFFFu) == 0) {  // a kernel that died never raises the flag
            const cudaError_t e = cudaStreamQuery(h->stream);
            if (e == cudaSuccess) {
                if (*flag == seq) break;
                fm_set_error("scalar kernel finished without a result");
                return FMIDX_ECUDA;
            }
            if (e != cudaErrorNotReady) { fm_set_error("scalar kernel failed: %s", cudaGetErrorString(e)); return FMIDX_ECUDA; }
        }
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    return FMIDX_OK;
}

int fm_scalar_max_tokens() { return FM_SCALAR_MAX_TOKENS; }

// get_range of ONE prefix.  speculate: the decode loop of the reference asks for the continuations of exactly this range next
// (get_range, then get_distinct_count on what it returned: document_constrained_generation_vllm.py:91-92), so the enumeration is
// launched right behind the search on the same stream, reading the range out of device memory; the host returns as soon as the
// SEARCH has answered, and a following fm_scalar_distinct(low, high, with_counts) with the same range finds its answer already
// in (or on its way into) the second pinned block -- one launch latency per pair instead of two.
int fm_scalar_search(fmidx *h, const int64_t *toks, int n, int64_t shift, int raw, uint64_t res[2]) {
    FmScalarOut *out;
    uint32_t seq;
    FM_TRY(fm_scalar_prepare(h, &out, &seq));
    FmToksParam tp;
    for (int i = 0; i < n; ++i) tp.t[i] = (toks[i] < -0x7FFFFFFFll || toks[i] > 0x7FFFFFFFll) ? -0x7FFFFFFF - 1 : (int32_t)toks[i];
    const bool speculate = !raw && h->tune_scalar_speculate;
    uint64_t *d_range = speculate ? (uint64_t *)h->scalar_d_range : nullptr;
    h->scalar_spec_seq = 0;  // whatever was speculated before is about to be overwritten
    uint32_t seq2 = 0;
    if (speculate) {
        seq2 = ++h->scalar_seq;
        if (seq2 == 0) seq2 = ++h->scalar_seq;
    }
    if (raw) FM_LAUNCH(k_search1<true>, 1, 32, 0, h->stream, h->dev, tp, n, shift, out, seq, d_range, (FmScalarOut *)h->scalar_spec_out, seq2);
    else FM_LAUNCH(k_search1<false>, 1, 32, 0, h->stream, h->dev, tp, n, shift, out, seq, d_range, (FmScalarOut *)h->scalar_spec_out, seq2);
    if (speculate) {
        FM_LAUNCH(k_distinct1, 1, DISTINCT1_THREADS, 0, h->stream, h->dev, 0ull, 0ull, 1, (uint32_t)h->scalar_cap_u64,
                  (FmScalarOut *)h->scalar_spec_out, seq2, d_range);
    }
    FM_TRY(fm_scalar_wait(h, out, seq));
    res[0] = out->v[0];
    res[1] = out->v[1];
    if (speculate) {
        h->scalar_spec_seq = seq2;
        h->scalar_spec_lo = res[0];
        h->scalar_spec_hi = res[1];
    }
    return FMIDX_OK;
}

int fm_scalar_step(fmidx *h, uint64_t sym, uint64_t low, uint64_t high, uint64_t res[2]) {
    FmScalarOut *out;
    uint32_t seq;
    FM_TRY(fm_scalar_prepare(h, &out, &seq));
    FM_LAUNCH(k_step1, 1, 32, 0, h->stream, h->dev, sym, low, high, out, seq);
    FM_TRY(fm_scalar_wait(h, out, seq));
    res[0] = out->v[0];
    res[1] = out->v[1];
    return FMIDX_OK;
}

int fm_scalar_locate(fmidx *h, uint64_t row, uint64_t *pos) {
    FmScalarOut *out;
    uint32_t seq;
    FM_TRY(fm_scalar_prepare(h, &out, &seq));
    FM_LAUNCH(k_locate1, 1, 32, 0, h->stream, h->dev, row, out, seq);
    FM_TRY(fm_scalar_wait(h, out, seq));
    *pos = out->v[0];
    return FMIDX_OK;
}

// *fallback = true: the range is wider than the scalar kernel handles (nothing written)
int fm_scalar_distinct(fmidx *h, uint64_t low, uint64_t high, int with_counts, uint64_t *dst, size_t cap, size_t *k, bool *fallback) {
    FmScalarOut *out;
    uint32_t seq;
    *fallback = false;
    if (with_counts && h->scalar_spec_seq && h->scalar_out && low == h->scalar_spec_lo && high == h->scalar_spec_hi) {
        // the enumeration of exactly this range was launched behind the search that returned it
        FmScalarOut *sp = (FmScalarOut *)h->scalar_spec_out;
        FM_TRY(fm_scalar_wait(h, sp, h->scalar_spec_seq));
        if (sp->status == 0) {
            *k = sp->k;
            if (sp->k > cap) { fm_set_error("output buffer too small: need %u", sp->k); return FMIDX_ERANGE; }
            if (sp->k) memcpy(dst, (const void *)sp->data, (size_t)sp->k * 8);
            return FMIDX_OK;
        }
        // too wide for the scalar kernel: the bitmap path below
    }
    FM_TRY(fm_scalar_prepare(h, &out, &seq));
    const uint64_t e = high > h->dev.n ? h->dev.n : high;
    if (e > low && e - low > (uint64_t)DISTINCT1_MAX) { *fallback = true; return FMIDX_OK; }
    FM_LAUNCH(k_distinct1, 1, DISTINCT1_THREADS, 0, h->stream, h->dev, low, high, with_counts, (uint32_t)h->scalar_cap_u64, out, seq,
              (const uint64_t *)nullptr);
    FM_TRY(fm_scalar_wait(h, out, seq));
    if (out->status != 0) { *fallback = true; return FMIDX_OK; }
    *k = out->k;
    if (out->k > cap) { fm_set_error("output buffer too small: need %u", out->k); return FMIDX_ERANGE; }
    if (out->k) memcpy(dst, (const void *)out->data, (size_t)out->k * 8);
    return FMIDX_OK;
}

int fm_launch_locate(fmidx *h, const uint64_t *d_rows, size_t n, uint64_t *d_pos, cudaStream_t s) {
    FM_TRY(check_ready(h));
    if (n == 0) return FMIDX_OK;
    if (h->dev.sa_dens == 1) FM_LAUNCH(k_locate_full, (unsigned)((n + 255) / 256), 256, 0, s, h->dev, d_rows, n, d_pos);
    else FM_LAUNCH(k_locate, (unsigned)((n * 8 + 127) / 128), 128, 0, s, h->dev, d_rows, n, d_pos);
    FM_CUDA(cudaGetLastError());
    return FMIDX_OK;
}

int fm_launch_extract(fmidx *h, const uint64_t *d_begin, const uint64_t *d_end, size_t n, uint64_t max_len,
                      uint64_t *d_out, uint64_t *d_len, cudaStream_t s) {
    FM_TRY(check_ready(h));
    if (n == 0) return FMIDX_OK;
    if (h->dev.text) FM_LAUNCH(k_extract_text, (unsigned)n, 256, 0, s, h->dev, d_begin, d_end, n, max_len, d_out, d_len);
    else FM_LAUNCH(k_extract, (unsigned)((n * 8 + 127) / 128), 128, 0, s, h->dev, d_begin, d_end, n, max_len, d_out, d_len);
    FM_CUDA(cudaGetLastError());
    return FMIDX_OK;
}
