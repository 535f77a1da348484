// fmidx_shard.cu -- one document shard's side of a sharded step: ONE host call (fmidx_shard_step) or ONE
// CUDA-graph launch (fmidx_shard_graph_launch) enqueues, on the buffer set's stream,
//   k_search     ranges of all B replicated prefixes on this shard
//   k_mask_top / k_mask_expand / k_mask_bytes   bitmaps + slice-occupancy words + occurrence counts into peer-visible buffers;
//                k_mask_top advances the set's step number IN DEVICE MEMORY (no per-step host argument: the step is capturable);
//                the CTA of k_mask_bytes that finishes last PUBLISHES the side data into every shard's inbox and raises the flags
//   k_xchg_wait  one small CTA: warp p spins until peer p has published this step
//   k_xchg_pull  one warp per (owned row, peer): the peer's occupied 4096-token slices over NVLink, ORed into the own row;
//                total counts
// -- six kernels, no NCCL, no host synchronisation.  Buffer sets are independent (own buffers, flags, step numbers, scratch slot,
// stream), so many steps overlap (bench.py keeps 24 in flight: the ranks' pipelines are coupled set by set).
// No acknowledgements are needed: every set PING-PONGS between two bitmap buffers and two inboxes (step parity).  When this
// rank's wait of step s-1 has returned, every peer has published step s-1, hence (stream order on the peer) finished its
// pull of step s-2 -- the last reader of the buffers that step s is about to overwrite.  History of what did NOT work: ack
// flags awaited at the head of the 1024-CTA k_mask_top (spinning CTAs held every thread slot whenever a peer was a little
// late: 59 M q/s on 4 GPUs against 131 M on one); a fused wait+pull kernel (its grid had to stay small because every CTA spun
// first -- 32 sets x 148 spinning CTAs deadlocked into the trap -- and a small grid has too few loads in flight for NVLink's
// loaded latency: 220 us of pulling per step on 4 GPUs, 79 M q/s; split: 8-10 us, 150 M q/s).  The reference has no
// distributed path (its only parallelism is one std::async per range, fm_index.cpp:146-166); parity for the sharded
// configuration is "one reference index per shard, results OR/sum-reduced on the host".
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>

#include "fmidx_internal.h"

struct fmidx_shard {
    fmidx *h = nullptr;
    uint32_t world = 0, rank = 0, vocab = 0, n_sets = 0;
    size_t B = 0, words = 0, ow = 0, count_off = 0, xwords = 0;
    bool ready = false;
    struct Set {
        uint32_t *mask[2] = {nullptr, nullptr};    // [B][words]    peer-visible (CUDA IPC), by step parity
        uint32_t *inbox[2] = {nullptr, nullptr};   // [world][xwords] peer-visible, by step parity
        uint32_t *xlocal = nullptr;  // [xwords] occupancy words | occurrence counts (u64)
        uint64_t *counts = nullptr;  // [B] totals over all shards
        uint32_t *n_own = nullptr;   // [B/world] distinct-token counts of the owned rows (host-buffer form)
        uint32_t *ctr = nullptr;     // ticket counter of the publish
        uint64_t *tables = nullptr;  // device: [2 parities][3][world] peer pointers (masks, inbox, data flags)
        uint32_t *data_flags = nullptr;  // into `flags`
        uint32_t seq = 0;            // steps issued on this set (host mirror of *seq_dev once they have run)
        uint32_t *seq_dev = nullptr; // the step number in device memory (advanced by the step's k_mask_top)
        cudaGraphExec_t gexec[2] = {nullptr, nullptr};  // captured step per buffer parity (fmidx_shard_graph_prepare)
        uint64_t g_launches = 0;
        cudaStream_t stream = nullptr;
        FmBuf in;                    // host-buffer form: prefixes + ranges on the device
    };
    std::vector<Set> sets;
    uint32_t *flags = nullptr;       // [n_sets][world] peer-visible
    unsigned long long *trace = nullptr;  // $FMIDX_PULL_TRACE: {flag-wait ns, total ns, samples} of the pull kernel's warps 0
    std::vector<std::vector<void *>> peer;  // [3 kinds][2 * n_sets * world] imported pointers (kind 2: one block, buffer 0 only)
    std::vector<void *> opened;
};

static int shard_check(fmidx_shard *x, uint32_t set, uint32_t mult = 1) {
    if (!x || !x->h) { fm_set_error("null shard context"); return FMIDX_EINVAL; }
    if (set >= x->n_sets * mult) { fm_set_error("buffer set %u out of range [0,%u)", set, x->n_sets); return FMIDX_EINVAL; }
    FM_CUDA(cudaSetDevice(x->h->device));
    return FMIDX_OK;
}

extern "C" {

int fmidx_shard_create(fmidx_t *h, uint32_t world, uint32_t rank, size_t B, uint32_t vocab, uint32_t n_sets,
                       fmidx_shard_t **out) {
    if (!h || !out) { fm_set_error("null argument"); return FMIDX_EINVAL; }
    *out = nullptr;
    if (world == 0 || world > 32 || rank >= world || B == 0 || B % world || vocab == 0 || n_sets == 0 || n_sets > FMIDX_SLOTS) {
        fm_set_error("shard_create: need 1 <= world <= 32, rank < world, B a positive multiple of world, 1 <= n_sets <= 32");
        return FMIDX_EINVAL;
    }
    FM_CUDA(cudaSetDevice(h->device));
    std::unique_ptr<fmidx_shard> x(new fmidx_shard());
    x->h = h; x->world = world; x->rank = rank; x->B = B; x->vocab = vocab; x->n_sets = n_sets;
    x->words = ((size_t)vocab + 31) / 32;
    x->ow = (((size_t)vocab + 4095) / 4096 + 31) / 32;
    x->count_off = (B * x->ow + 1) & ~(size_t)1;              // the counts are u64
    x->xwords = (x->count_off + 2 * B + 3) & ~(size_t)3;      // whole 16-byte vectors
    x->sets.resize(n_sets);
    x->peer.assign(3, std::vector<void *>((size_t)2 * n_sets * world, nullptr));
    const size_t flag_words = (size_t)n_sets * world;
    FM_CUDA(cudaMalloc((void **)&x->flags, flag_words * 4));
    FM_CUDA(cudaMemset(x->flags, 0, flag_words * 4));
    for (uint32_t i = 0; i < n_sets; ++i) {
        fmidx_shard::Set &st = x->sets[i];
        for (int par = 0; par < 2; ++par) {
            FM_CUDA(cudaMalloc((void **)&st.mask[par], B * x->words * 4));
            FM_CUDA(cudaMalloc((void **)&st.inbox[par], (size_t)world * x->xwords * 4));
            FM_CUDA(cudaMemset(st.mask[par], 0, B * x->words * 4));
            FM_CUDA(cudaMemset(st.inbox[par], 0, (size_t)world * x->xwords * 4));
            x->peer[0][((size_t)i * 2 + par) * world + rank] = st.mask[par];
            x->peer[1][((size_t)i * 2 + par) * world + rank] = st.inbox[par];
        }
        FM_CUDA(cudaMalloc((void **)&st.xlocal, x->xwords * 4));
        FM_CUDA(cudaMalloc((void **)&st.counts, B * 8));
        FM_CUDA(cudaMalloc((void **)&st.n_own, (B / world) * 4 + 16));
        FM_CUDA(cudaMalloc((void **)&st.ctr, 16));
        FM_CUDA(cudaMalloc((void **)&st.seq_dev, 16));
        FM_CUDA(cudaMemset(st.seq_dev, 0, 16));
        FM_CUDA(cudaMalloc((void **)&st.tables, (size_t)2 * 3 * world * 8));
        FM_CUDA(cudaMemset(st.xlocal, 0, x->xwords * 4));
        FM_CUDA(cudaMemset(st.ctr, 0, 16));
        FM_CUDA(cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking));
        st.data_flags = x->flags + (size_t)i * world;
    }
    x->peer[2][rank] = x->flags;
    if (getenv("FMIDX_PULL_TRACE")) {
        FM_CUDA(cudaMalloc((void **)&x->trace, 64));
        FM_CUDA(cudaMemset(x->trace, 0, 64));
    }
    FM_CUDA(cudaDeviceSynchronize());
    *out = x.release();
    return FMIDX_OK;
}

// kind 0: bitmap buffer `buf`, 1: inbox `buf` (buf = 2 * set + parity, 0 <= buf < 2 * n_sets), 2: the flag block (`buf` ignored)
int fmidx_shard_export(fmidx_shard_t *x, uint32_t buf, int kind, unsigned char handle[64]) {
    FM_TRY(shard_check(x, kind == 2 ? 0 : buf, 2));
    if (!handle || kind < 0 || kind > 2) { fm_set_error("bad argument"); return FMIDX_EINVAL; }
    void *p = kind == 0 ? (void *)x->sets[buf >> 1].mask[buf & 1] : kind == 1 ? (void *)x->sets[buf >> 1].inbox[buf & 1] : (void *)x->flags;
    cudaIpcMemHandle_t hd;
    FM_CUDA(cudaIpcGetMemHandle(&hd, p));
    memcpy(handle, &hd, 64);
    return FMIDX_OK;
}

int fmidx_shard_import(fmidx_shard_t *x, uint32_t buf, int kind, uint32_t peer, const unsigned char handle[64]) {
    const uint32_t set = buf;
    FM_TRY(shard_check(x, kind == 2 ? 0 : buf, 2));
    if (!handle || kind < 0 || kind > 2 || peer >= x->world || peer == x->rank) { fm_set_error("bad argument"); return FMIDX_EINVAL; }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle, 64);
    void *p = nullptr;
    FM_CUDA(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    x->opened.push_back(p);
    x->peer[kind][(kind == 2 ? 0 : (size_t)set * x->world) + peer] = p;
    return FMIDX_OK;
}

// after every peer's buffers were imported: upload the pointer tables
int fmidx_shard_ready(fmidx_shard_t *x) {
    FM_TRY(shard_check(x, 0));
    const uint32_t W = x->world;
    for (uint32_t i = 0; i < x->n_sets; ++i) {
        std::vector<uint64_t> t((size_t)2 * 3 * W);
        for (uint32_t par = 0; par < 2; ++par)
            for (uint32_t p = 0; p < W; ++p) {
                void *m = x->peer[0][((size_t)i * 2 + par) * W + p], *ib = x->peer[1][((size_t)i * 2 + par) * W + p], *fl = x->peer[2][p];
                if (!m || !ib || !fl) { fm_set_error("shard_ready: buffers of peer %u (set %u) were not imported", p, i); return FMIDX_ESTATE; }
                t[(size_t)par * 3 * W + p] = (uint64_t)(uintptr_t)m;
                t[(size_t)par * 3 * W + W + p] = (uint64_t)(uintptr_t)ib;
                t[(size_t)par * 3 * W + 2 * W + p] = (uint64_t)(uintptr_t)((uint32_t *)fl + (size_t)i * W);
            }
        FM_CUDA(cudaMemcpy(x->sets[i].tables, t.data(), t.size() * 8, cudaMemcpyHostToDevice));
    }
    x->ready = true;
    return FMIDX_OK;
}

// enqueues (or, while the stream is being captured, records) the step number `seq` of the set; the kernels take the number
// from device memory, `seq` only selects the buffer parity
static int shard_step_seq(fmidx_shard *x, uint32_t set, uint32_t seq, const int32_t *d_toks, const int64_t *d_offs, int64_t shift,
                          uint64_t *d_sp, uint64_t *d_ep, cudaStream_t s) {
    fmidx *h = x->h;
    fmidx_shard::Set &st = x->sets[set];
    const uint32_t W = x->world;
    const size_t own = x->B / W;
    const uint32_t par = seq & 1u;
    uint32_t *mask = st.mask[par];
    const uint64_t *tab = st.tables + (size_t)par * 3 * W;
    const bool kz = fm_search_can_zero(mask, x->vocab);
    FM_TRY(fm_launch_search(h, d_toks, d_offs, x->B, shift, d_sp, d_ep, 0, s, kz ? mask : nullptr, (uint32_t)x->words,
                            (size_t)x->rank * own, own));
    FmXchgArgs xa;
    xa.xcounts = (uint64_t *)(st.xlocal + x->count_off);
    xa.peer_inbox = (uint32_t *const *)(tab + W);
    xa.peer_data_flags = (uint32_t *const *)(tab + 2 * W);
    xa.xlocal = st.xlocal; xa.done_ctr = st.ctr; xa.xwords = x->xwords; xa.world = W; xa.rank = x->rank; xa.seq = seq;
    xa.seq_dev = st.seq_dev;
    // measurement only ($FMIDX_SHARD_PROBE): 1 = no pull kernel (results incomplete), 2 = neither publish nor pull
    static const int probe = [] { const char *e = getenv("FMIDX_SHARD_PROBE"); return e ? atoi(e) : 0; }();
    if (probe >= 2) xa.world = 0;
    h->mask_ws_override = &h->slots[set].ws;
    struct Reset { fmidx *h; ~Reset() { h->mask_ws_override = nullptr; } } reset{h};
    // only the rows this rank owns come out as complete bitmaps; the others are written slice-wise (occupied slices)
    FM_TRY(fm_launch_mask(h, d_sp, d_ep, x->B, shift, x->vocab, nullptr, nullptr, 3, 0, nullptr, -1, nullptr, 0, mask, nullptr, s,
                          st.xlocal, (size_t)x->rank * own, own, kz, &xa));
    if (probe >= 1) return FMIDX_OK;
    return fm_launch_xchg_pull(h, (uint32_t *const *)tab, st.data_flags, st.inbox[par], x->xwords, x->count_off, W, x->rank, x->B,
                               x->vocab, seq, st.counts, s, x->trace, x->n_sets, st.seq_dev);
}

static int shard_step(fmidx_shard *x, uint32_t set, const int32_t *d_toks, const int64_t *d_offs, int64_t shift, uint64_t *d_sp,
                      uint64_t *d_ep, cudaStream_t s) {
    if (!x->ready) { fm_set_error("shard context is not ready (import the peers' buffers, then fmidx_shard_ready)"); return FMIDX_ESTATE; }
    return shard_step_seq(x, set, ++x->sets[set].seq, d_toks, d_offs, shift, d_sp, d_ep, s);
}

// The sharded step of one buffer set captured for FIXED device buffers (the caller rewrites d_toks / d_offs in place between
// launches): one eager step first (it sizes the workspaces and is a real step: EVERY shard must prepare the same sets in the
// same order), then one capture per buffer parity; fmidx_shard_graph_launch replays them alternately on the set's stream.
int fmidx_shard_graph_prepare(fmidx_shard_t *x, uint32_t set, const int32_t *d_toks, const int64_t *d_offs, int64_t shift, uint64_t *d_sp,
                              uint64_t *d_ep) {
    FM_TRY(shard_check(x, set));
    if (!d_toks || !d_offs || !d_sp || !d_ep) { fm_set_error("null argument"); return FMIDX_EINVAL; }
    if (!x->h->ready) { fm_set_error("index is not initialised"); return FMIDX_ESTATE; }
    fmidx_shard::Set &st = x->sets[set];
    if (st.gexec[0]) { fm_set_error("buffer set %u already holds a captured step", set); return FMIDX_ESTATE; }
    FM_TRY(shard_step(x, set, d_toks, d_offs, shift, d_sp, d_ep, st.stream));
    FM_CUDA(cudaStreamSynchronize(st.stream));
    for (uint32_t e = 1; e <= 2; ++e) {
        const uint32_t seq = st.seq + e;  // the next two steps: parities (seq+1)&1, (seq+2)&1
        const uint64_t l0 = g_fm_launches.load();
        FM_CUDA(cudaStreamBeginCapture(st.stream, cudaStreamCaptureModeThreadLocal));
        int rc = shard_step_seq(x, set, seq, d_toks, d_offs, shift, d_sp, d_ep, st.stream);
        cudaGraph_t g = nullptr;
        cudaError_t ce = cudaStreamEndCapture(st.stream, &g);
        if (rc != FMIDX_OK) { if (g) cudaGraphDestroy(g); return rc; }
        if (ce != cudaSuccess) { fm_set_error("stream capture failed: %s", cudaGetErrorString(ce)); return FMIDX_ECUDA; }
        ce = cudaGraphInstantiate(&st.gexec[seq & 1u], g, 0);
        cudaGraphDestroy(g);
        if (ce != cudaSuccess) { fm_set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ce)); return FMIDX_ECUDA; }
        st.g_launches = g_fm_launches.load() - l0;
        g_fm_launches.store(l0);  // captured, not launched
    }
    return FMIDX_OK;
}

int fmidx_shard_graph_launch(fmidx_shard_t *x, uint32_t set) {
    if (!x || set >= x->n_sets || !x->sets[set].gexec[0] || !x->sets[set].gexec[1]) { fm_set_error("buffer set holds no captured step"); return FMIDX_ESTATE; }
    fmidx_shard::Set &st = x->sets[set];
    const uint32_t seq = ++st.seq;
    FM_CUDA(cudaGraphLaunch(st.gexec[seq & 1u], st.stream));
    g_fm_launches.fetch_add(st.g_launches, std::memory_order_relaxed);
    return FMIDX_OK;
}

int fmidx_shard_step(fmidx_shard_t *x, uint32_t set, const int32_t *d_toks, const int64_t *d_offs, int64_t shift, uint64_t *d_sp,
                     uint64_t *d_ep, void *stream) {
    FM_TRY(shard_check(x, set));
    if (!d_toks || !d_offs || !d_sp || !d_ep) { fm_set_error("null argument"); return FMIDX_EINVAL; }
    if (!x->h->ready) { fm_set_error("index is not initialised"); return FMIDX_ESTATE; }
    return shard_step(x, set, d_toks, d_offs, shift, d_sp, d_ep, stream ? (cudaStream_t)stream : x->sets[set].stream);
}

int fmidx_shard_result(fmidx_shard_t *x, uint32_t set, int parity, uint32_t **d_own_masks, uint64_t **d_counts, void **stream) {
    FM_TRY(shard_check(x, set));
    // the sets ping-pong between two bitmap buffers: step number k (1, 2, ...) of a set writes buffer k & 1.
    // parity 0 / 1: that buffer; -1: the buffer of the step issued last
    const uint32_t par = parity < 0 ? (x->sets[set].seq & 1u) : ((uint32_t)parity & 1u);
    if (d_own_masks) *d_own_masks = x->sets[set].mask[par] + (size_t)x->rank * (x->B / x->world) * x->words;
    if (d_counts) *d_counts = x->sets[set].counts;
    if (stream) *stream = x->sets[set].stream;
    return FMIDX_OK;
}

// host-buffer form on the set's own stream: H2D of the prefixes, the step, D2H of this shard's ranges, the total
// occurrence counts of all B prefixes and the distinct-token counts of the B/world prefixes this rank owns
int fmidx_shard_submit(fmidx_shard_t *x, uint32_t set, const int32_t *toks, const int64_t *offs, int64_t shift, uint64_t *sp,
                       uint64_t *ep, uint64_t *counts, uint32_t *n_distinct_own) {
    FM_TRY(shard_check(x, set));
    if (!toks || !offs || !sp || !ep) { fm_set_error("null argument"); return FMIDX_EINVAL; }
    if (!x->h->ready) { fm_set_error("index is not initialised"); return FMIDX_ESTATE; }
    fmidx_shard::Set &st = x->sets[set];
    const size_t B = x->B, ntok = (size_t)offs[B], own = B / x->world;
    cudaStream_t s = st.stream;
    FM_TRY(fm_scratch(x->h, st.in, (B + 1) * 8 + B * 16 + ntok * 4 + 16, false));
    char *dp = (char *)st.in.p;
    int64_t *d_offs = (int64_t *)dp;
    uint64_t *d_sp = (uint64_t *)(dp + (B + 1) * 8), *d_ep = d_sp + B;
    int32_t *d_toks = (int32_t *)(d_ep + B);
    FM_CUDA(cudaMemcpyAsync(d_offs, offs, (B + 1) * 8, cudaMemcpyHostToDevice, s));
    if (ntok) FM_CUDA(cudaMemcpyAsync(d_toks, toks, ntok * 4, cudaMemcpyHostToDevice, s));
    FM_TRY(shard_step(x, set, d_toks, d_offs, shift, d_sp, d_ep, s));
    FM_CUDA(cudaMemcpyAsync(sp, d_sp, B * 8, cudaMemcpyDeviceToHost, s));
    FM_CUDA(cudaMemcpyAsync(ep, d_ep, B * 8, cudaMemcpyDeviceToHost, s));
    if (counts) FM_CUDA(cudaMemcpyAsync(counts, st.counts, B * 8, cudaMemcpyDeviceToHost, s));
    if (n_distinct_own) {
        FM_TRY(fm_launch_popcount(x->h, st.mask[st.seq & 1u] + (size_t)x->rank * own * x->words, own, x->vocab, st.n_own, s));
        FM_CUDA(cudaMemcpyAsync(n_distinct_own, st.n_own, own * 4, cudaMemcpyDeviceToHost, s));
    }
    return FMIDX_OK;
}

int fmidx_shard_wait(fmidx_shard_t *x, uint32_t set) {
    FM_TRY(shard_check(x, set));
    FM_CUDA(cudaStreamSynchronize(x->sets[set].stream));
    return FMIDX_OK;
}

// every rank calls this after a barrier that follows its last step (peers may still be pulling otherwise)
void fmidx_shard_destroy(fmidx_shard_t *x) {
    if (!x) return;
    cudaSetDevice(x->h->device);
    cudaDeviceSynchronize();
    if (x->trace) {
        unsigned long long t[3] = {0, 0, 0};
        cudaMemcpy(t, x->trace, 24, cudaMemcpyDeviceToHost);
        unsigned long long t3 = 0;
        cudaMemcpy(&t3, x->trace + 3, 8, cudaMemcpyDeviceToHost);
        if (t[2] && t3) fprintf(stderr, "fmidx pull trace (rank %u): %.1f us flag wait per peer (k_xchg_wait), %.1f us per CTA of k_xchg_pull\n",
                                x->rank, t[0] / 1e3 / t3, t[1] / 1e3 / t[2]);
        cudaFree(x->trace);
    }
    for (void *p : x->opened) cudaIpcCloseMemHandle(p);
    for (auto &st : x->sets) {
        for (auto &g : st.gexec) if (g) cudaGraphExecDestroy(g);
        if (st.seq_dev) cudaFree(st.seq_dev);
        if (st.stream) cudaStreamDestroy(st.stream);
        for (void *p : {(void *)st.mask[0], (void *)st.mask[1], (void *)st.inbox[0], (void *)st.inbox[1], (void *)st.xlocal, (void *)st.counts, (void *)st.n_own, (void *)st.ctr, (void *)st.tables})
            if (p) cudaFree(p);
        if (st.in.p) cudaFree(st.in.p);
    }
    if (x->flags) cudaFree(x->flags);
    delete x;
}

}  // extern "C"
