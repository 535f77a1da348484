// fmidx_internal.h -- host-side handle and helpers shared by the builder, the kernel
// launchers and the C ABI (include/fmidx.h).
#pragma once
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/fmidx.h"
#include "fmidx_layout.cuh"

#define FMIDX_SLOTS 32  // pipelined slots / scratch workspaces / buffer sets per handle

void fm_set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_fm_launches;

#define FM_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            fm_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return FMIDX_ECUDA;                                                                    \
        }                                                                                          \
    } while (0)

#define FM_TRY(expr)            \
    do {                        \
        int _r = (expr);        \
        if (_r != FMIDX_OK) return _r; \
    } while (0)

// every kernel launch of this library goes through this macro (gpu_launches accounting)
#define FM_LAUNCH(kernel, grid, block, smem, stream, ...)                    \
    do {                                                                     \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);          \
        g_fm_launches.fetch_add(1, std::memory_order_relaxed);               \
    } while (0)

struct FmBuf {  // growable device / pinned-host scratch
    void *p = nullptr;
    size_t cap = 0;
    bool pinned_host = false;
    // mask workspaces only: the two work-list counter pairs at the start of the buffer alternate between calls and
    // each call's last kernel zeroes the pair of the next one (no memset node per call)
    uint32_t epoch = 0;
    bool ctr_clean = false;
    bool row_clean = false;  // words 8, 9 (claim / exit counters of k_mask_logits_row) are zero
    // shared (non-slot) mask workspaces: last use, so that a call on another stream orders itself behind it
    cudaEvent_t last_use = nullptr;
    cudaStream_t last_stream = nullptr;
};

// sharded step: the parts of the cross-shard exchange that are folded into the enumeration kernels
struct FmXchgArgs {
    uint64_t *xcounts = nullptr;            // [B] occurrence counts (ep - sp) of this shard, part of its side data
    // publication at the tail of the last enumeration kernel (world == 0: none)
    uint32_t *const *peer_inbox = nullptr;
    uint32_t *const *peer_data_flags = nullptr;
    const uint32_t *xlocal = nullptr;
    uint32_t *done_ctr = nullptr;
    size_t xwords = 0;
    uint32_t world = 0, rank = 0, seq = 0;
    uint32_t *seq_dev = nullptr;            // the set's step number in device memory: advanced by k_mask_top, read by publish / pull
};

struct fmidx {
    int device = 0;
    bool ready = false;
    FmDev dev{};
    uint32_t opt_sa_dens = 32;   // csa_wt_int<> default t_dens
    uint32_t opt_isa_dens = 64;  // csa_wt_int<> default t_inv_dens
    int opt_keep_text = 0;
    std::vector<void *> owned;   // device allocations that make up the index
    size_t dev_bytes = 0;
    cudaStream_t stream = nullptr;  // private stream of the synchronous host-style calls
    FmBuf d_scratch[8];
    FmBuf mask_ws;               // node tables + work list of the mask kernels: device-style calls (caller's stream)
    FmBuf mask_ws_host;          // same for the host-style calls on the private stream
    FmBuf *mask_ws_override = nullptr;  // set while a pipelined slot launches (its own workspace)
    struct Slot {                // fmidx_query_submit / fmidx_query_wait: independent in-flight host batches
        cudaStream_t stream = nullptr;
        FmBuf in, mask, ws;
        // lazy clearing of the slot's mask rows (k_mask_top): per-row slice occupancy of what the slot's last call left in the
        // buffer `lazy_ptr` ([lazy_B] rows of a lazy_vocab-bit mask); the association is dropped whenever a call cannot use it
        FmBuf occ;
        const uint32_t *lazy_ptr = nullptr;
        size_t lazy_B = 0;
        uint32_t lazy_vocab = 0;
        // graph-captured form (fmidx_query_graph_*): the whole step of the slot as two executable graphs (the mask kernels
        // alternate between two work-list counter pairs, so consecutive steps alternate between the two captures)
        cudaGraphExec_t gexec[2] = {nullptr, nullptr};
        void *g_hin = nullptr, *g_hout = nullptr;   // pinned host staging: [offs | toks] in, [sp | ep | n_distinct] out
        void *g_din = nullptr, *g_dout = nullptr;
        // device-resident captured form (fmidx_query_graph_prepare_device): the caller's buffers, no copies
        const int32_t *g_dev_toks = nullptr; const int64_t *g_dev_offs = nullptr;
        uint64_t *g_dev_sp = nullptr, *g_dev_ep = nullptr; uint32_t *g_dev_mask = nullptr;
        size_t g_B = 0, g_maxtok = 0;
        uint64_t g_launches = 0;
        uint32_t g_step = 0;
    } slots[FMIDX_SLOTS];
    FmBuf h_scratch[4];
    int sm_count = 148;
    // tunables (fmidx_tune; initialised from the environment by fmidx_new)
    uint32_t tune_scan_max = 16384;  // ranges of up to this many rows are read off the resident BWT ($FMIDX_SCAN_MAX)
    int tune_bytes_occ = 5;          // CTAs per SM of the persistent k_mask_bytes grid ($FMIDX_BYTES_OCC)
    int tune_unigram = 1;            // use the unigram continuation table at query time ($FMIDX_UNIGRAM_TABLE)
    // scalar calls: result block in mapped pinned host memory + the sequence number of the last call
    void *scalar_out = nullptr;
    uint32_t scalar_seq = 0;
    size_t scalar_cap_u64 = 0;
    // speculative enumeration behind a scalar search (fm_scalar_search): its result block, its sequence number (0: none) and the
    // range it was launched for
    void *scalar_spec_out = nullptr;
    void *scalar_d_range = nullptr;
    uint32_t scalar_spec_seq = 0;
    uint64_t scalar_spec_lo = 0, scalar_spec_hi = 0;
    int tune_scalar_speculate = 1;   // $FMIDX_SCALAR_SPECULATE
    int tune_scalar_fast = 1;        // scalar calls take the zero-copy single-launch kernels ($FMIDX_SCALAR_FAST)
    int tune_lazy_clear = 1;         // 0 never; 1 mask buffers the library owns (pipelined / captured host forms); 2 also the caller's
                                     // buffers of fmidx_query_batch / fmidx_query_graph_prepare_device ($FMIDX_LAZY_CLEAR)
    int tune_probe_phase = 0;        // measurement only: fmidx_query_batch runs one phase alone (1 search, 2 enumeration)
    uint32_t tune_unigram_max = 1024;  // table entries built at most ($FMIDX_UNIGRAM_MAX; 0: none)
};

int fm_build_unigram_table(fmidx *h, cudaStream_t s);

int fm_dev_alloc(fmidx *h, void **out, size_t bytes, bool owned);
int fm_scratch(fmidx *h, FmBuf &b, size_t bytes, bool pinned_host);
void fm_release_index(fmidx *h);

// builder (fmidx_build.cu)
int fm_build_from_device_text(fmidx *h, const uint32_t *d_symbols, size_t n_data, cudaStream_t s);

// kernel launchers (fmidx_kernels.cu); all enqueue on `s`
int fm_launch_search(fmidx *h, const int32_t *d_toks, const int64_t *d_offs, size_t B, int64_t shift, uint64_t *d_sp,
                     uint64_t *d_ep, int raw, cudaStream_t s, uint32_t *d_zero_rows = nullptr, uint32_t zero_words = 0,
                     size_t zero_lo = 0, size_t zero_n = ~(size_t)0);
bool fm_search_can_zero(const uint32_t *d_mask, uint32_t vocab);
int fm_launch_step(fmidx *h, const uint64_t *d_sym, const uint64_t *d_lo, const uint64_t *d_hi, size_t B,
                   uint64_t *d_sp, uint64_t *d_ep, cudaStream_t s);
int fm_launch_decode_step(fmidx *h, const int64_t *d_last_tok, int64_t tok_stride, const int32_t *d_parent, const uint64_t *d_sp_in,
                          const uint64_t *d_ep_in, uint64_t *d_sp_out, uint64_t *d_ep_out, size_t B, int64_t shift, int64_t eos,
                          int64_t pad, int64_t stop_at_count, uint64_t *d_qsp, uint64_t *d_qep, int32_t *d_override, cudaStream_t s);
int fm_launch_mask(fmidx *h, const uint64_t *d_sp, const uint64_t *d_ep, size_t B, int64_t shift, uint32_t vocab,
                   const void *d_logits, void *d_out, int dtype, int64_t row_stride, const int32_t *d_override,
                   int32_t always_allow, const uint32_t *d_extra, uint32_t n_extra, uint32_t *d_mask,
                   uint32_t *d_counts, cudaStream_t s, uint32_t *d_occ = nullptr, size_t dense_lo = 0, size_t dense_n = ~(size_t)0,
                   bool prezeroed = false, const FmXchgArgs *xa = nullptr, uint32_t *d_ndist = nullptr, uint32_t *d_occ_state = nullptr);
int fm_launch_xchg_pull(fmidx *h, uint32_t *const *d_peer_masks, const uint32_t *d_data_flags, const uint32_t *d_inbox, size_t xwords,
                        size_t count_off, uint32_t world, uint32_t rank, size_t B, uint32_t vocab, uint32_t seq, uint64_t *d_counts_out,
                        cudaStream_t s, unsigned long long *d_trace = nullptr, uint32_t n_sets = 1, const uint32_t *d_seq = nullptr);
int fm_launch_boost_apply(fmidx *h, const void *d_logits, void *d_out, int dtype, int64_t row_stride, size_t B, uint32_t vocab,
                          const uint32_t *d_dmask, const uint32_t *d_n_distinct, const uint32_t *d_unigram, const int32_t *d_kind,
                          const float *d_vd, const float *d_vu_nonempty, const float *d_vu_empty, const float *d_ve, int32_t eos,
                          cudaStream_t s);
int fm_launch_or(fmidx *h, const uint32_t *d_pieces, uint32_t n_pieces, size_t n_words, uint32_t *d_out, cudaStream_t s);
int fm_launch_popcount(fmidx *h, const uint32_t *d_mask, size_t B, uint32_t vocab, uint32_t *d_n, cudaStream_t s);
int fm_launch_gather_probe(fmidx *h, int unit_bytes, size_t n_loads, double *out_gbps, cudaStream_t s);
int fm_launch_compact(fmidx *h, const uint32_t *d_mask, const uint32_t *d_counts, size_t B, uint32_t vocab,
                      const uint64_t *d_offs, int64_t *d_tokens, uint64_t *d_cnts, cudaStream_t s);
int fm_launch_locate(fmidx *h, const uint64_t *d_rows, size_t n, uint64_t *d_pos, cudaStream_t s);
// scalar (batch 1) calls on h->stream: arguments as kernel parameters, results polled from mapped pinned host memory
int fm_scalar_max_tokens();
int fm_scalar_search(fmidx *h, const int64_t *toks, int n, int64_t shift, int raw, uint64_t res[2]);
int fm_scalar_step(fmidx *h, uint64_t sym, uint64_t low, uint64_t high, uint64_t res[2]);
int fm_scalar_locate(fmidx *h, uint64_t row, uint64_t *pos);
int fm_scalar_distinct(fmidx *h, uint64_t low, uint64_t high, int with_counts, uint64_t *dst, size_t cap, size_t *k, bool *fallback);
int fm_launch_extract(fmidx *h, const uint64_t *d_begin, const uint64_t *d_end, size_t n, uint64_t max_len,
                      uint64_t *d_out, uint64_t *d_len, cudaStream_t s);
