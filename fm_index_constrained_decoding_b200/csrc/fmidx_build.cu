// fmidx_build.cu -- on-device index construction for sm_100a.
//
// Replaces sdsl::construct / construct_im as called by FMIndex::initialize and
// FMIndex::initialize_from_file (reference cpp_modules_linux/fm_index.cpp:33-48):
//   text  = symbols || 0                       (sdsl appends the sentinel; 0 symbols are rejected)
//   SA    : prefix doubling, one 64-bit radix sort (CUB) per round  (sdsl: qsufsort on the CPU)
//   BWT   : BWT[j] = T[SA[j]-1]
//   WM    : 16-ary wavelet matrix, one stable counting pass per level, cells as in
//           fmidx_layout.cuh (sdsl: wt_int + rank_support_v + two select supports)
//   C, D  : int_alphabet's C array on the raw symbol axis; D[c] = C[c] - descend_c(0)
//   SA / ISA samples (sdsl: sa_order_sa_sampling<32>, isa_sampling<64>; here text-order SA
//           samples so a locate walk is bounded by sa_dens-1 LF steps; results are identical)
// The suffix array of a text with a unique smallest sentinel is unique, so the index content
// is the same function of the text as the reference's.
#include <algorithm>

#include <cub/cub.cuh>

#include "fmidx_internal.h"

namespace {

struct MaxOp {
    __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

__global__ void k_copy_text(const uint32_t *__restrict__ src, uint32_t *__restrict__ T, size_t n_data) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i <= n_data; i += stride) T[i] = i < n_data ? src[i] : 0u;
}

__global__ void k_make_keys(const uint32_t *__restrict__ rank, size_t N, size_t h, uint32_t rb,
                            uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < N; i += stride) {
        uint64_t lo = (i + h < N) ? rank[i + h] : 0u;
        keys[i] = ((uint64_t)rank[i] << rb) | lo;
        vals[i] = (uint32_t)i;
    }
}

// head[j] = j when sorted key j starts a new group, else 0; counts the groups.
__global__ void k_group_heads(const uint64_t *__restrict__ keys, size_t N, uint32_t *__restrict__ head,
                              unsigned long long *__restrict__ n_groups) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned int local = 0;
    for (; j < N; j += stride) {
        bool is_head = (j == 0) || (keys[j] != keys[j - 1]);
        head[j] = is_head ? (uint32_t)j : 0u;
        local += is_head ? 1u : 0u;
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(n_groups, (unsigned long long)local);
}

__global__ void k_scatter_rank(const uint32_t *__restrict__ vals, const uint32_t *__restrict__ grp, size_t N,
                               uint32_t *__restrict__ rank) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; j < N; j += stride) rank[vals[j]] = grp[j];
}

__global__ void k_bwt(const uint32_t *__restrict__ T, const uint32_t *__restrict__ SA, size_t N,
                      uint32_t *__restrict__ bwt) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; j < N; j += stride) {
        uint32_t p = SA[j];
        bwt[j] = T[p ? p - 1 : N - 1];
    }
}

// C[c] = first row whose suffix starts with a symbol >= c  (the F column is sorted).
__global__ void k_c_array(const uint32_t *__restrict__ T, const uint32_t *__restrict__ SA, size_t N, uint32_t max_sym,
                          uint32_t *__restrict__ C) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; j < N; j += stride) {
        uint32_t f = T[SA[j]];
        if (j == 0) {
            for (uint32_t c = 0; c <= f; ++c) C[c] = 0;
        } else {
            uint32_t fp = T[SA[j - 1]];
            for (uint32_t c = fp + 1; c <= f; ++c) C[c] = (uint32_t)j;
        }
        if (j == N - 1)
            for (uint32_t c = f + 1; c <= max_sym + 1; ++c) C[c] = (uint32_t)N;
    }
}

// ---- wavelet-matrix level construction: one warp per 128-position cell ------------------------
__global__ void k_level_planes(const uint32_t *__restrict__ cur, size_t N, uint32_t shift, uint32_t *__restrict__ cells,
                               uint32_t *__restrict__ hist, size_t nblocks) {
    const int lane = threadIdx.x & 31;
    size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t b = warp; b < nblocks; b += nwarps) {
        uint32_t cnt = 0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            size_t p = b * 128 + (size_t)g * 32 + lane;
            bool valid = p < N;
            uint32_t dig = valid ? ((cur[p] >> shift) & 15u) : 0u;
            uint4 pl;
            pl.x = __ballot_sync(0xffffffffu, valid && (dig & 1u));
            pl.y = __ballot_sync(0xffffffffu, valid && (dig & 2u));
            pl.z = __ballot_sync(0xffffffffu, valid && (dig & 4u));
            pl.w = __ballot_sync(0xffffffffu, valid && (dig & 8u));
            uint32_t vm = __ballot_sync(0xffffffffu, valid);
            if (lane < 16) cnt += __popc(fm_match(pl, (uint32_t)lane) & vm);
            if (lane == 0) reinterpret_cast<uint4 *>(cells + b * 32)[4 + g] = pl;
        }
        if (lane < 16) hist[(size_t)lane * nblocks + b] = cnt;
    }
}

__global__ void k_level_hdr(const uint32_t *__restrict__ scanned, uint32_t *__restrict__ cells, size_t nblocks) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; t < nblocks * 16; t += stride) {
        size_t d = t / nblocks, b = t - d * nblocks;
        cells[b * 32 + d] = scanned[t];
    }
}

__global__ void k_level_scatter(const uint32_t *__restrict__ cur, uint32_t *__restrict__ nxt, size_t N, uint32_t shift,
                                const uint32_t *__restrict__ cells, size_t nblocks) {
    const int lane = threadIdx.x & 31;
    size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    for (size_t b = warp; b < nblocks; b += nwarps) {
        const uint32_t *cell = cells + b * 32;
        uint32_t x[4], dig[4], within[4];
        bool valid[4];
        uint4 pl[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            size_t p = b * 128 + (size_t)g * 32 + lane;
            valid[g] = p < N;
            x[g] = valid[g] ? cur[p] : 0u;
            dig[g] = (x[g] >> shift) & 15u;
            pl[g] = reinterpret_cast<const uint4 *>(cell)[4 + g];
        }
        // invalid tail positions carry digit 0 planes == 0 bits, i.e. they look like digit 0; they
        // sit after every valid position of the cell, so they never disturb a valid element's rank.
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            uint32_t before = 0;
#pragma unroll
            for (int g2 = 0; g2 < 4; ++g2)
                if (g2 < g) before += __popc(fm_match(pl[g2], dig[g]));
            within[g] = before + __popc(fm_match(pl[g], dig[g]) & lt);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g)
            if (valid[g]) nxt[cell[dig[g]] + within[g]] = x[g];
    }
}

__global__ void k_d_table(FmDev ix, uint32_t *__restrict__ D) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c > ix.max_sym) return;
    D[c] = ix.C[c] - fm_descend1(ix, 0u, (uint32_t)c);
}

// ---- byte level: one warp per 512-position cell -------------------------------------------------------
__global__ void __launch_bounds__(256) k_byte_cells(const uint32_t *__restrict__ cur, size_t N, uint32_t *__restrict__ bcells,
                                                     uint32_t *__restrict__ hist, size_t nbblocks) {
    __shared__ uint32_t cnt[8][256];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t b = warp; b < nbblocks; b += nwarps) {
        for (int v = lane; v < 256; v += 32) cnt[wid][v] = 0;
        __syncwarp();
        unsigned char *payload = reinterpret_cast<unsigned char *>(bcells + b * (FMIDX_BCELL_U4 * 4) + FMIDX_BHDR_U4 * 4);
#pragma unroll 4
        for (int k = 0; k < FMIDX_BBLOCK / 32; ++k) {
            const size_t p = b * FMIDX_BBLOCK + (size_t)k * 32 + lane;
            const bool valid = p < N;
            const uint32_t v = valid ? (cur[p] & 255u) : 0u;
            payload[k * 32 + lane] = (unsigned char)v;
            if (valid) atomicAdd(&cnt[wid][v], 1u);
        }
        __syncwarp();
        for (int v = lane; v < 256; v += 32) hist[(size_t)v * nbblocks + b] = cnt[wid][v];
        __syncwarp();
    }
}

__global__ void k_byte_hdr(const uint32_t *__restrict__ scanned, uint32_t *__restrict__ bcells, size_t nbblocks) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; t < nbblocks * 256; t += stride) {
        size_t v = t / nbblocks, b = t - v * nbblocks;
        bcells[b * (FMIDX_BCELL_U4 * 4) + v] = scanned[t];
    }
}

// ---- SA / ISA samples ---------------------------------------------------------------------------
__global__ void k_mark_words(const uint32_t *__restrict__ SA, size_t N, uint32_t dens, uint32_t *__restrict__ bits,
                             uint32_t *__restrict__ cnt, size_t nwords) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // grid covers nwords*32 rows exactly
    bool f = j < N && (SA[j] % dens) == 0;
    uint32_t w = __ballot_sync(0xffffffffu, f);
    if ((threadIdx.x & 31) == 0 && (j >> 5) < nwords) {
        bits[j >> 5] = w;
        cnt[j >> 5] = __popc(w);
    }
}

__global__ void k_mark_finish(const uint32_t *__restrict__ SA, size_t N, const uint32_t *__restrict__ bits,
                              const uint32_t *__restrict__ before, uint2 *__restrict__ marks,
                              uint32_t *__restrict__ samples, size_t nwords) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t w = j >> 5;
    if (w >= nwords) return;
    const int lane = threadIdx.x & 31;
    uint32_t b = bits[w], pre = before[w];
    if (lane == 0) marks[w] = make_uint2(pre, b);
    if (j < N && ((b >> lane) & 1u)) samples[pre + __popc(b & ((1u << lane) - 1u))] = SA[j];
}

__global__ void k_isa_samples(const uint32_t *__restrict__ SA, size_t N, uint32_t dens, uint32_t *__restrict__ isa) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; j < N; j += stride) {
        uint32_t p = SA[j];
        if (p % dens == 0) isa[p / dens] = (uint32_t)j;
    }
}

inline unsigned grid_for(size_t n, unsigned block, int sms) {
    size_t g = (n + block - 1) / block;
    size_t cap = (size_t)sms * 16;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

inline uint32_t bits_of(uint64_t x) {
    uint32_t b = 0;
    while (x) { ++b; x >>= 1; }
    return b;
}

struct TmpFree {  // frees builder temporaries on every exit path
    std::vector<void *> v;
    ~TmpFree() { for (void *p : v) if (p) cudaFree(p); }
    void drop(void *p) { for (auto &q : v) if (q == p) q = nullptr; }
};

}  // namespace

int fm_build_from_device_text(fmidx *h, const uint32_t *d_symbols, size_t n_data, cudaStream_t s) {
    if (n_data >= 0xFFFFFF00ull) {
        fm_set_error("text too long for one index (%zu symbols >= 2^32-256): shard the corpus", n_data);
        return FMIDX_ERANGE;
    }
    fm_release_index(h);
    const size_t N = n_data + 1;
    const int sms = h->sm_count;
    TmpFree tmp;
    auto talloc = [&](void **p, size_t bytes) -> int {
        cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
        if (e != cudaSuccess) {
            fm_set_error("cudaMalloc(%zu) failed during build: %s", bytes, cudaGetErrorString(e));
            return FMIDX_ECUDA;
        }
        tmp.v.push_back(*p);
        return FMIDX_OK;
    };

    // ---- text, alphabet bounds ------------------------------------------------------------------
    uint32_t *T = nullptr;
    FM_TRY(talloc((void **)&T, N * 4));
    FM_LAUNCH(k_copy_text, grid_for(N, 256, sms), 256, 0, s, d_symbols, T, n_data);
    uint32_t *d_red = nullptr;
    FM_TRY(talloc((void **)&d_red, 16));
    uint32_t max_sym = 0, min_sym = 1;
    void *cub_tmp = nullptr;
    size_t cub_bytes = 0;
    auto cub_reserve = [&](size_t need) -> int {
        if (need <= cub_bytes) return FMIDX_OK;
        void *p = nullptr;
        FM_TRY(talloc(&p, need));
        cub_tmp = p;
        cub_bytes = need;
        return FMIDX_OK;
    };
    if (n_data) {
        size_t need = 0, need2 = 0;
        cub::DeviceReduce::Max(nullptr, need, d_symbols, d_red, (int64_t)n_data, s);
        cub::DeviceReduce::Min(nullptr, need2, d_symbols, d_red + 1, (int64_t)n_data, s);
        FM_TRY(cub_reserve(need > need2 ? need : need2));
        FM_CUDA(cub::DeviceReduce::Max(cub_tmp, cub_bytes, d_symbols, d_red, (int64_t)n_data, s));
        FM_CUDA(cub::DeviceReduce::Min(cub_tmp, cub_bytes, d_symbols, d_red + 1, (int64_t)n_data, s));
        uint32_t mm[2];
        FM_CUDA(cudaMemcpyAsync(mm, d_red, 8, cudaMemcpyDeviceToHost, s));
        FM_CUDA(cudaStreamSynchronize(s));
        max_sym = mm[0];
        min_sym = mm[1];
    }
    if (min_sym == 0) {
        fm_set_error("text contains the reserved symbol 0 (sdsl construct: 'contains zero symbol')");
        return FMIDX_EZERO;
    }
    if (max_sym >= (1u << (4 * FMIDX_MAX_LEVELS))) {
        fm_set_error("symbol %u does not fit %d bits", max_sym, 4 * FMIDX_MAX_LEVELS);
        return FMIDX_ERANGE;
    }
    uint32_t sym_bits = bits_of(max_sym);
    const uint32_t levels = sym_bits > 8 ? (sym_bits - 8 + 3) / 4 : 0;  // nibble levels above the byte level

    // ---- suffix array by prefix doubling -----------------------------------------------------------
    uint32_t *rank = nullptr, *head = nullptr, *valA = nullptr, *valB = nullptr;
    uint64_t *keyA = nullptr, *keyB = nullptr;
    unsigned long long *d_groups = nullptr;
    FM_TRY(talloc((void **)&rank, N * 4));
    FM_TRY(talloc((void **)&head, N * 4));
    FM_TRY(talloc((void **)&valA, N * 4));
    FM_TRY(talloc((void **)&valB, N * 4));
    FM_TRY(talloc((void **)&keyA, N * 8));
    FM_TRY(talloc((void **)&keyB, N * 8));
    FM_TRY(talloc((void **)&d_groups, 8));
    FM_CUDA(cudaMemcpyAsync(rank, T, N * 4, cudaMemcpyDeviceToDevice, s));
    uint32_t *SA = nullptr;
    {
        uint32_t rb = sym_bits ? sym_bits : 1;  // bits of the current rank values
        const uint32_t rb_n = bits_of(N - 1) ? bits_of(N - 1) : 1;
        size_t need = 0, need2 = 0;
        cub::DoubleBuffer<uint64_t> kb0(keyA, keyB);
        cub::DoubleBuffer<uint32_t> vb0(valA, valB);
        cub::DeviceRadixSort::SortPairs(nullptr, need, kb0, vb0, (int64_t)N, 0, 64, s);
        cub::DeviceScan::InclusiveScan(nullptr, need2, head, head, MaxOp(), (int64_t)N, s);
        FM_TRY(cub_reserve(need > need2 ? need : need2));
        for (size_t hh = 1;; hh *= 2) {
            cub::DoubleBuffer<uint64_t> kb(keyA, keyB);
            cub::DoubleBuffer<uint32_t> vb(valA, valB);
            FM_LAUNCH(k_make_keys, grid_for(N, 256, sms), 256, 0, s, rank, N, hh, rb, kb.Current(), vb.Current());
            FM_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, kb, vb, (int64_t)N, 0, (int)(2 * rb), s));
            g_fm_launches.fetch_add(1, std::memory_order_relaxed);
            FM_CUDA(cudaMemsetAsync(d_groups, 0, 8, s));
            FM_LAUNCH(k_group_heads, grid_for(N, 256, sms), 256, 0, s, kb.Current(), N, head, d_groups);
            unsigned long long groups = 0;
            FM_CUDA(cudaMemcpyAsync(&groups, d_groups, 8, cudaMemcpyDeviceToHost, s));
            FM_CUDA(cudaStreamSynchronize(s));
            if (groups == N || hh >= N) {
                SA = vb.Current();
                break;
            }
            FM_CUDA(cub::DeviceScan::InclusiveScan(cub_tmp, cub_bytes, head, head, MaxOp(), (int64_t)N, s));
            FM_LAUNCH(k_scatter_rank, grid_for(N, 256, sms), 256, 0, s, vb.Current(), head, N, rank);
            rb = rb_n;
        }
    }
    // SA now lives in valA or valB.  Release the sort buffers we no longer need.
    FM_CUDA(cudaStreamSynchronize(s));
    cudaFree(keyA); tmp.drop(keyA);
    cudaFree(keyB); tmp.drop(keyB);
    cudaFree(rank); tmp.drop(rank);
    uint32_t *other_val = (SA == valA) ? valB : valA;

    // ---- C array (raw symbol axis) and BWT -----------------------------------------------------------
    uint32_t *C = nullptr, *D = nullptr;
    FM_TRY(fm_dev_alloc(h, (void **)&C, ((size_t)max_sym + 2) * 4, true));
    FM_TRY(fm_dev_alloc(h, (void **)&D, ((size_t)max_sym + 1) * 4, true));
    FM_LAUNCH(k_c_array, grid_for(N, 256, sms), 256, 0, s, T, SA, N, max_sym, C);
    uint32_t *cur = other_val, *nxt = head;  // reuse two N-sized u32 buffers for the level permutation
    FM_LAUNCH(k_bwt, grid_for(N, 256, sms), 256, 0, s, T, SA, N, cur);

    // ---- wavelet matrix: `levels` nibble levels, then the byte level --------------------------------------
    const size_t nblocks = (N >> 7) + 1;
    const size_t nbblocks = (N >> 9) + 1;
    uint32_t *cells = nullptr, *bcells = nullptr, *hist = nullptr;
    FM_TRY(fm_dev_alloc(h, (void **)&cells, (size_t)levels * nblocks * 128, true));
    FM_TRY(fm_dev_alloc(h, (void **)&bcells, nbblocks * (size_t)FMIDX_BCELL_U4 * 16, true));
    const size_t hist_n = std::max(nblocks * 16, nbblocks * 256);
    FM_TRY(talloc((void **)&hist, hist_n * 4));
    {
        size_t need = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, need, hist, hist, (int64_t)hist_n, s);
        FM_TRY(cub_reserve(need));
    }
    for (uint32_t l = 0; l < levels; ++l) {
        uint32_t shift = 8u + 4u * (levels - 1u - l);
        uint32_t *lc = cells + (size_t)l * nblocks * 32;
        FM_LAUNCH(k_level_planes, grid_for(nblocks * 32, 256, sms), 256, 0, s, cur, N, shift, lc, hist, nblocks);
        FM_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, hist, hist, (int64_t)(nblocks * 16), s));
        g_fm_launches.fetch_add(1, std::memory_order_relaxed);
        FM_LAUNCH(k_level_hdr, grid_for(nblocks * 16, 256, sms), 256, 0, s, hist, lc, nblocks);
        FM_LAUNCH(k_level_scatter, grid_for(nblocks * 32, 256, sms), 256, 0, s, cur, nxt, N, shift, lc, nblocks);
        uint32_t *t2 = cur; cur = nxt; nxt = t2;
    }
    FM_LAUNCH(k_byte_cells, grid_for(nbblocks * 32, 256, sms), 256, 0, s, cur, N, bcells, hist, nbblocks);
    FM_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, hist, hist, (int64_t)(nbblocks * 256), s));
    g_fm_launches.fetch_add(1, std::memory_order_relaxed);
    FM_LAUNCH(k_byte_hdr, grid_for(nbblocks * 256, 256, sms), 256, 0, s, hist, bcells, nbblocks);

    FmDev &dv = h->dev;
    dv = FmDev{};
    dv.cells = reinterpret_cast<const uint4 *>(cells);
    dv.bcells = reinterpret_cast<const uint4 *>(bcells);
    dv.nbblocks = (uint32_t)nbblocks;
    dv.C = C;
    dv.D = D;
    dv.level_stride = nblocks * FMIDX_CELL_U4;
    dv.n = (uint32_t)N;
    dv.max_sym = max_sym;
    dv.levels = levels;
    dv.nblocks = (uint32_t)nblocks;
    dv.sa_dens = h->opt_sa_dens ? h->opt_sa_dens : 1;
    dv.isa_dens = h->opt_isa_dens ? h->opt_isa_dens : 1;
    FM_LAUNCH(k_d_table, (unsigned)(((size_t)max_sym + 256) / 256), 256, 0, s, dv, D);

    // ---- SA / ISA samples ----------------------------------------------------------------------------------
    if (dv.sa_dens == 1) {
        uint32_t *sa_keep = nullptr;
        FM_TRY(fm_dev_alloc(h, (void **)&sa_keep, N * 4, true));
        FM_CUDA(cudaMemcpyAsync(sa_keep, SA, N * 4, cudaMemcpyDeviceToDevice, s));
        dv.sa_samples = sa_keep;
        dv.sa_marks = nullptr;
    } else {
        const size_t nwords = (N + 31) / 32;
        const size_t nsamp = (N - 1) / dv.sa_dens + 1;
        uint32_t *bits = nullptr, *cnt = nullptr, *samples = nullptr;
        uint2 *marks = nullptr;
        FM_TRY(talloc((void **)&bits, nwords * 4));
        FM_TRY(talloc((void **)&cnt, nwords * 4));
        FM_TRY(fm_dev_alloc(h, (void **)&marks, nwords * 8, true));
        FM_TRY(fm_dev_alloc(h, (void **)&samples, nsamp * 4, true));
        unsigned g = (unsigned)((nwords * 32 + 255) / 256);
        FM_LAUNCH(k_mark_words, g, 256, 0, s, SA, N, dv.sa_dens, bits, cnt, nwords);
        size_t need = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, need, cnt, cnt, (int64_t)nwords, s);
        FM_TRY(cub_reserve(need));
        FM_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, cnt, cnt, (int64_t)nwords, s));
        g_fm_launches.fetch_add(1, std::memory_order_relaxed);
        FM_LAUNCH(k_mark_finish, g, 256, 0, s, SA, N, bits, cnt, marks, samples, nwords);
        dv.sa_marks = marks;
        dv.sa_samples = samples;
    }
    {
        const size_t nisa = (N - 1) / dv.isa_dens + 1;
        uint32_t *isa = nullptr;
        FM_TRY(fm_dev_alloc(h, (void **)&isa, nisa * 4, true));
        FM_LAUNCH(k_isa_samples, grid_for(N, 256, sms), 256, 0, s, SA, N, dv.isa_dens, isa);
        dv.isa_samples = isa;
    }
    if (h->opt_keep_text) {
        uint32_t *tk = nullptr;
        FM_TRY(fm_dev_alloc(h, (void **)&tk, N * 4, true));
        FM_CUDA(cudaMemcpyAsync(tk, T, N * 4, cudaMemcpyDeviceToDevice, s));
        dv.text = tk;
        // resident layout: the BWT itself as plain symbols (the level permutation consumed the first copy)
        uint32_t *bw = nullptr;
        FM_TRY(fm_dev_alloc(h, (void **)&bw, N * 4, true));
        FM_LAUNCH(k_bwt, grid_for(N, 256, sms), 256, 0, s, T, SA, N, bw);
        dv.bwt = bw;
    }
    FM_CUDA(cudaStreamSynchronize(s));
    FM_CUDA(cudaGetLastError());
    h->ready = true;
    return FMIDX_OK;
}
