// fmidx_capi.cu -- the extern "C" boundary declared in include/fmidx.h.
// Replaces the SWIG wrapper generated from cpp_modules_linux/fm_index.i:7-21 of the reference.
#include <algorithm>
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>

#include "fmidx_internal.h"

std::atomic<uint64_t> g_fm_launches{0};
static thread_local char g_err[512] = "";

void fm_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int fm_dev_alloc(fmidx *h, void **out, size_t bytes, bool owned) {
    cudaError_t e = cudaMalloc(out, bytes + 16);  // vector loads may read up to 12 bytes past the last element
    if (e != cudaSuccess) {
        fm_set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return FMIDX_ECUDA;
    }
    if (owned) {
        h->owned.push_back(*out);
        h->dev_bytes += bytes;
    }
    return FMIDX_OK;
}

int fm_scratch(fmidx *h, FmBuf &b, size_t bytes, bool pinned_host) {
    (void)h;
    if (bytes <= b.cap && b.p) return FMIDX_OK;
    if (b.p) {
        if (b.pinned_host) cudaFreeHost(b.p); else cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes < 4096 ? 4096 : bytes + bytes / 4;
    cudaError_t e = pinned_host ? cudaMallocHost(&b.p, want) : cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        b.p = nullptr;
        fm_set_error("scratch allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
        return FMIDX_ECUDA;
    }
    b.cap = want;
    b.pinned_host = pinned_host;
    b.epoch = 0;
    b.ctr_clean = false;
    b.row_clean = false;
    return FMIDX_OK;
}

void fm_release_index(fmidx *h) {
    h->scalar_spec_seq = 0;  // a speculated enumeration belongs to the index it ran on
    for (void *p : h->owned) cudaFree(p);
    h->owned.clear();
    h->dev_bytes = 0;
    h->dev = FmDev{};
    h->ready = false;
}

static int enter(fmidx *h, bool need_ready) {
    if (!h) {
        fm_set_error("null handle");
        return FMIDX_EINVAL;
    }
    FM_CUDA(cudaSetDevice(h->device));
    if (need_ready && !h->ready) {
        fm_set_error("index is not initialised");
        return FMIDX_ESTATE;
    }
    return FMIDX_OK;
}

// Lazy clearing (k_mask_top): the per-row slice occupancy of what slot `sl` last left in `d_mask`, or nullptr when this call must
// zero whole rows (feature off for this kind of buffer, or more than 32 slices per row).  A new (buffer, shape) association
// starts from "everything may be dirty".  `owned`: the buffer belongs to the library (level 1) or to the caller (level 2).
static int lazy_state(fmidx *h, fmidx::Slot &sl, const uint32_t *d_mask, size_t B, uint32_t vocab, bool owned, cudaStream_t s,
                      uint32_t **out) {
    *out = nullptr;
    if (h->tune_lazy_clear < (owned ? 1 : 2) || !d_mask || B == 0 || (vocab + 4095u) / 4096u > 32u) {
        sl.lazy_ptr = nullptr;  // the rows are about to be rewritten without book-keeping
        return FMIDX_OK;
    }
    if (sl.lazy_ptr != d_mask || sl.lazy_B != B || sl.lazy_vocab != vocab || !sl.occ.p) {
        if (sl.occ.p && B * 4 > sl.occ.cap) FM_CUDA(cudaStreamSynchronize(s));  // growing frees the old array
        FM_TRY(fm_scratch(h, sl.occ, B * 4, false));
        FM_CUDA(cudaMemsetAsync(sl.occ.p, 0xFF, B * 4, s));
        sl.lazy_ptr = d_mask;
        sl.lazy_B = B;
        sl.lazy_vocab = vocab;
    }
    *out = (uint32_t *)sl.occ.p;
    return FMIDX_OK;
}

// Unigram continuation table (fmidx_layout.cuh): built with the library's own enumeration kernels, once per index.
int fm_build_unigram_table(fmidx *h, cudaStream_t s) {
    FmDev &d = h->dev;
    d.uni_n = 0;
    d.uni_sp = d.uni_ep = d.uni_masks = nullptr;
    if (!d.bwt || h->tune_unigram_max == 0 || d.n < 2) return FMIDX_OK;  // resident layout only
    const uint32_t nsym = d.max_sym + 1;
    std::vector<uint32_t> C(nsym + 1);
    FM_CUDA(cudaMemcpyAsync(C.data(), d.C, (size_t)(nsym + 1) * 4, cudaMemcpyDeviceToHost, s));
    FM_CUDA(cudaStreamSynchronize(s));
    // the symbols whose block is too wide to be read row by row, most frequent first, at most tune_unigram_max
    const uint32_t min_rows = std::max<uint32_t>(h->tune_scan_max, 1024u);
    std::vector<uint32_t> cand;
    for (uint32_t c = 1; c < nsym; ++c)
        if (C[c + 1] - C[c] > min_rows) cand.push_back(c);
    if (cand.size() > h->tune_unigram_max) {
        std::nth_element(cand.begin(), cand.begin() + h->tune_unigram_max, cand.end(),
                         [&](uint32_t a, uint32_t b) { return C[a + 1] - C[a] > C[b + 1] - C[b]; });
        cand.resize(h->tune_unigram_max);
    }
    std::sort(cand.begin(), cand.end());
    const size_t n = cand.size();
    if (n == 0) return FMIDX_OK;
    const uint32_t words = (nsym + 31u) >> 5;
    std::vector<uint32_t> sp(n), ep(n);
    std::vector<uint64_t> sp64(n), ep64(n);
    for (size_t i = 0; i < n; ++i) {
        sp[i] = C[cand[i]]; ep[i] = C[cand[i] + 1];
        sp64[i] = sp[i]; ep64[i] = ep[i];
    }
    uint32_t *d_sp = nullptr, *d_ep = nullptr, *d_masks = nullptr;
    uint64_t *d_r = nullptr;
    FM_TRY(fm_dev_alloc(h, (void **)&d_sp, n * 4, true));
    FM_TRY(fm_dev_alloc(h, (void **)&d_ep, n * 4, true));
    FM_TRY(fm_dev_alloc(h, (void **)&d_masks, n * (size_t)words * 4, true));
    FM_CUDA(cudaMalloc((void **)&d_r, n * 16));
    FM_CUDA(cudaMemcpyAsync(d_sp, sp.data(), n * 4, cudaMemcpyHostToDevice, s));
    FM_CUDA(cudaMemcpyAsync(d_ep, ep.data(), n * 4, cudaMemcpyHostToDevice, s));
    FM_CUDA(cudaMemcpyAsync(d_r, sp64.data(), n * 8, cudaMemcpyHostToDevice, s));
    FM_CUDA(cudaMemcpyAsync(d_r + n, ep64.data(), n * 8, cudaMemcpyHostToDevice, s));
    // symbol axis: shift 0, "vocab" = every symbol; the table is not in use yet (uni_n == 0), so this is the plain walk
    int rc = fm_launch_mask(h, d_r, d_r + n, n, 0, nsym, nullptr, nullptr, 3, 0, nullptr, -1, nullptr, 0, d_masks, nullptr, s);
    cudaStreamSynchronize(s);
    cudaFree(d_r);
    FM_TRY(rc);
    d.uni_sp = d_sp; d.uni_ep = d_ep; d.uni_masks = d_masks; d.uni_n = (uint32_t)n; d.uni_words = words;
    return FMIDX_OK;
}

extern "C" {

const char *fmidx_last_error(void) { return g_err; }
int fmidx_abi_version(void) { return 1; }
uint64_t fmidx_launch_count(void) { return g_fm_launches.load(); }

int fmidx_new(int device, fmidx_t **out) {
    if (!out) { fm_set_error("out is NULL"); return FMIDX_EINVAL; }
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        fm_set_error("no CUDA device available (%s): libfmidx has no CPU fallback",
                     e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return FMIDX_ECUDA;
    }
    if (device < 0 || device >= ndev) { fm_set_error("device %d out of range [0,%d)", device, ndev); return FMIDX_EINVAL; }
    FM_CUDA(cudaSetDevice(device));
    std::unique_ptr<fmidx> h(new fmidx());
    h->device = device;
    FM_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    int sms = 0;
    FM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    h->sm_count = sms > 0 ? sms : 148;
    auto env_long = [](const char *name, long dflt) {
        const char *e = getenv(name);
        return e && *e ? atol(e) : dflt;
    };
    h->tune_scan_max = (uint32_t)std::min<long>(std::max<long>(env_long("FMIDX_SCAN_MAX", 16384), 0), 1l << 24);
    h->tune_bytes_occ = (int)env_long("FMIDX_BYTES_OCC", 5);
    h->tune_unigram = env_long("FMIDX_UNIGRAM_TABLE", 1) ? 1 : 0;
    h->tune_scalar_fast = env_long("FMIDX_SCALAR_FAST", 1) ? 1 : 0;
    h->tune_lazy_clear = (int)std::min<long>(2, std::max<long>(0, env_long("FMIDX_LAZY_CLEAR", 1)));
    h->tune_scalar_speculate = env_long("FMIDX_SCALAR_SPECULATE", 1) ? 1 : 0;
    h->tune_unigram_max = (uint32_t)std::min<long>(std::max<long>(env_long("FMIDX_UNIGRAM_MAX", 1024), 0), 65536);
    *out = h.release();
    return FMIDX_OK;
}

void fmidx_free(fmidx_t *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    fm_release_index(h);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->scalar_out) cudaFreeHost(h->scalar_out);
    if (h->scalar_d_range) cudaFree(h->scalar_d_range);
    for (auto &b : h->d_scratch) if (b.p) cudaFree(b.p);
    for (FmBuf *b : {&h->mask_ws, &h->mask_ws_host}) {
        if (b->last_use) { cudaEventSynchronize(b->last_use); cudaEventDestroy(b->last_use); }
        if (b->p) cudaFree(b->p);
    }
    for (auto &sl : h->slots) {
        if (sl.stream) { cudaStreamSynchronize(sl.stream); cudaStreamDestroy(sl.stream); }
        for (auto &g : sl.gexec) if (g) cudaGraphExecDestroy(g);
        if (sl.g_hin) cudaFreeHost(sl.g_hin);
        if (sl.g_hout) cudaFreeHost(sl.g_hout);
        if (sl.g_din) cudaFree(sl.g_din);
        if (sl.g_dout) cudaFree(sl.g_dout);
        if (sl.in.p) cudaFree(sl.in.p);
        if (sl.occ.p) cudaFree(sl.occ.p);
        if (sl.mask.p) cudaFree(sl.mask.p);
        if (sl.ws.p) cudaFree(sl.ws.p);
    }
    for (auto &b : h->h_scratch) if (b.p) cudaFreeHost(b.p);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int fmidx_set_options(fmidx_t *h, uint32_t sa_dens, uint32_t isa_dens, int keep_text) {
    if (!h) { fm_set_error("null handle"); return FMIDX_EINVAL; }
    if (sa_dens == 0 || isa_dens == 0) { fm_set_error("sample densities must be >= 1"); return FMIDX_EINVAL; }
    h->opt_sa_dens = sa_dens;
    h->opt_isa_dens = isa_dens;
    h->opt_keep_text = keep_text ? 1 : 0;
    return FMIDX_OK;
}

// ---- tunables ---------------------------------------------------------------------------------------
int fmidx_tune(fmidx_t *h, const char *key, int64_t value) {
    if (!h || !key) { fm_set_error("null argument"); return FMIDX_EINVAL; }
    if (!strcmp(key, "scan_max")) h->tune_scan_max = (uint32_t)std::min<int64_t>(std::max<int64_t>(value, 0), 1ll << 24);
    else if (!strcmp(key, "bytes_occ")) h->tune_bytes_occ = (int)value;
    else if (!strcmp(key, "unigram_table")) h->tune_unigram = value ? 1 : 0;
    else if (!strcmp(key, "probe_phase")) h->tune_probe_phase = (int)value;
    else if (!strcmp(key, "scalar_fast")) h->tune_scalar_fast = value ? 1 : 0;
    else if (!strcmp(key, "lazy_clear")) h->tune_lazy_clear = value < 0 ? 0 : (value > 2 ? 2 : (int)value);
    else if (!strcmp(key, "scalar_speculate")) { h->tune_scalar_speculate = value ? 1 : 0; h->scalar_spec_seq = 0; }
    else { fm_set_error("fmidx_tune: unknown key '%s'", key); return FMIDX_EINVAL; }
    return FMIDX_OK;
}

// ---- build ------------------------------------------------------------------------------------------
static int build_from_host_u32(fmidx *h, const std::vector<uint32_t> &sym) {
    uint32_t *d = nullptr;
    FM_CUDA(cudaMalloc((void **)&d, sym.size() * 4 + 16));
    cudaError_t e = cudaMemcpyAsync(d, sym.data(), sym.size() * 4, cudaMemcpyHostToDevice, h->stream);
    int rc = FMIDX_OK;
    if (e != cudaSuccess) {
        fm_set_error("H2D copy of the text failed: %s", cudaGetErrorString(e));
        rc = FMIDX_ECUDA;
    } else {
        rc = fm_build_from_device_text(h, d, sym.size(), h->stream);
        if (rc == FMIDX_OK) rc = fm_build_unigram_table(h, h->stream);
    }
    cudaStreamSynchronize(h->stream);
    cudaFree(d);
    return rc;
}

int fmidx_initialize(fmidx_t *h, const uint64_t *data, size_t n) {
    FM_TRY(enter(h, false));
    if (n && !data) { fm_set_error("data is NULL"); return FMIDX_EINVAL; }
    std::vector<uint32_t> sym(n);
    for (size_t i = 0; i < n; ++i) {
        if (data[i] == 0) { fm_set_error("text contains the reserved symbol 0 at position %zu", i); return FMIDX_EZERO; }
        if (data[i] >> (4 * FMIDX_MAX_LEVELS)) { fm_set_error("symbol %llu at position %zu does not fit %d bits", (unsigned long long)data[i], i, 4 * FMIDX_MAX_LEVELS); return FMIDX_ERANGE; }
        sym[i] = (uint32_t)data[i];
    }
    return build_from_host_u32(h, sym);
}

int fmidx_initialize_from_file(fmidx_t *h, const char *path, int width) {
    FM_TRY(enter(h, false));
    if (!path) { fm_set_error("path is NULL"); return FMIDX_EINVAL; }
    if (width != 1 && width != 2 && width != 4 && width != 8) {
        fm_set_error("initialize_from_file: width %d unsupported (1, 2, 4 or 8 bytes per symbol)", width);
        return FMIDX_EINVAL;
    }
    FILE *f = fopen(path, "rb");
    if (!f) { fm_set_error("cannot open %s: %s", path, strerror(errno)); return FMIDX_EIO; }
    fseek(f, 0, SEEK_END);
    long long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    size_t n = (size_t)bytes / (size_t)width;
    std::vector<unsigned char> raw((size_t)bytes);
    if (bytes && fread(raw.data(), 1, (size_t)bytes, f) != (size_t)bytes) {
        fclose(f);
        fm_set_error("short read on %s", path);
        return FMIDX_EIO;
    }
    fclose(f);
    std::vector<uint32_t> sym(n);
    for (size_t i = 0; i < n; ++i) {
        uint64_t v = 0;
        memcpy(&v, raw.data() + i * (size_t)width, (size_t)width);  // little-endian, like sdsl's int reader
        if (v == 0) { fm_set_error("file %s contains the reserved symbol 0 at position %zu", path, i); return FMIDX_EZERO; }
        if (v >> (4 * FMIDX_MAX_LEVELS)) { fm_set_error("symbol %llu does not fit %d bits", (unsigned long long)v, 4 * FMIDX_MAX_LEVELS); return FMIDX_ERANGE; }
        sym[i] = (uint32_t)v;
    }
    return build_from_host_u32(h, sym);
}

int fmidx_initialize_device(fmidx_t *h, const uint32_t *d_symbols, size_t n, void *stream) {
    FM_TRY(enter(h, false));
    if (n && !d_symbols) { fm_set_error("d_symbols is NULL"); return FMIDX_EINVAL; }
    FM_TRY(fm_build_from_device_text(h, d_symbols, n, (cudaStream_t)stream));
    return fm_build_unigram_table(h, (cudaStream_t)stream);
}

// ---- persistence ------------------------------------------------------------------------------------
static const char FM_MAGIC[8] = {'F', 'M', 'I', 'D', 'X', 'B', '2', '\0'};
struct FmFileHeader {
    char magic[8];
    uint32_t version, levels, max_sym, sa_dens, isa_dens, keep_text, nblocks, nbblocks;
    uint64_t n;
    uint64_t bytes[9];  // nibble cells, C, D, sa_marks, sa_samples, isa_samples, text, byte-level cells, bwt
};
constexpr int FM_SECTIONS = 9;

static void section_sizes(const FmDev &d, uint64_t b[FM_SECTIONS]) {
    const uint64_t N = d.n;
    b[0] = (uint64_t)d.levels * d.nblocks * 128;
    b[1] = ((uint64_t)d.max_sym + 2) * 4;
    b[2] = ((uint64_t)d.max_sym + 1) * 4;
    b[3] = d.sa_dens == 1 ? 0 : ((N + 31) / 32) * 8;
    b[4] = d.sa_dens == 1 ? N * 4 : ((N - 1) / d.sa_dens + 1) * 4;
    b[5] = ((N - 1) / d.isa_dens + 1) * 4;
    b[6] = d.text ? N * 4 : 0;
    b[7] = (uint64_t)d.nbblocks * FMIDX_BCELL_U4 * 16;
    b[8] = d.bwt ? N * 4 : 0;
}

int fmidx_save(fmidx_t *h, const char *path) {
    FM_TRY(enter(h, true));
    if (!path) { fm_set_error("path is NULL"); return FMIDX_EINVAL; }
    FILE *f = fopen(path, "wb");
    if (!f) { fm_set_error("cannot open %s for writing: %s", path, strerror(errno)); return FMIDX_EIO; }
    FmFileHeader hd{};
    memcpy(hd.magic, FM_MAGIC, 8);
    const FmDev &d = h->dev;
    hd.version = 3; hd.nbblocks = d.nbblocks; hd.levels = d.levels; hd.max_sym = d.max_sym; hd.sa_dens = d.sa_dens; hd.isa_dens = d.isa_dens;
    hd.keep_text = d.text ? 1 : 0; hd.nblocks = d.nblocks; hd.n = d.n;
    section_sizes(d, hd.bytes);
    const void *src[FM_SECTIONS] = {d.cells, d.C, d.D, d.sa_marks, d.sa_samples, d.isa_samples, d.text, d.bcells, d.bwt};
    int rc = FMIDX_OK;
    if (fwrite(&hd, sizeof(hd), 1, f) != 1) rc = FMIDX_EIO;
    const size_t CH = 64u << 20;
    std::vector<unsigned char> buf(CH);
    for (int sct = 0; sct < FM_SECTIONS && rc == FMIDX_OK; ++sct) {
        for (uint64_t off = 0; off < hd.bytes[sct] && rc == FMIDX_OK; off += CH) {
            size_t len = (size_t)std::min<uint64_t>(CH, hd.bytes[sct] - off);
            cudaError_t e = cudaMemcpy(buf.data(), (const char *)src[sct] + off, len, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { fm_set_error("D2H failed while saving: %s", cudaGetErrorString(e)); rc = FMIDX_ECUDA; }
            else if (fwrite(buf.data(), 1, len, f) != len) rc = FMIDX_EIO;
        }
    }
    if (fclose(f) != 0 && rc == FMIDX_OK) rc = FMIDX_EIO;
    if (rc == FMIDX_EIO) fm_set_error("write error on %s", path);
    return rc;
}

int fmidx_load(const char *path, int device, fmidx_t **out) {
    if (!path || !out) { fm_set_error("null argument"); return FMIDX_EINVAL; }
    *out = nullptr;
    fmidx_t *h = nullptr;
    FM_TRY(fmidx_new(device, &h));
    FILE *f = fopen(path, "rb");
    if (!f) { fm_set_error("cannot open %s: %s", path, strerror(errno)); fmidx_free(h); return FMIDX_EIO; }
    FmFileHeader hd{};
    int rc = FMIDX_OK;
    if (fread(&hd, sizeof(hd), 1, f) != 1 || memcmp(hd.magic, FM_MAGIC, 8) != 0 || hd.version != 3) {
        fm_set_error("%s is not an fmidx-b200 index file (bad magic/version)", path);
        rc = FMIDX_EIO;
    }
    if (rc == FMIDX_OK) {
        FmDev d{};
        d.levels = hd.levels; d.max_sym = hd.max_sym; d.sa_dens = hd.sa_dens; d.isa_dens = hd.isa_dens;
        d.nblocks = hd.nblocks; d.nbblocks = hd.nbblocks; d.n = (uint32_t)hd.n; d.level_stride = (uint64_t)hd.nblocks * FMIDX_CELL_U4;
        uint64_t expect[FM_SECTIONS];
        FmDev probe = d;
        probe.text = hd.keep_text ? (const uint32_t *)1 : nullptr;
        probe.bwt = hd.bytes[8] ? (const uint32_t *)1 : nullptr;
        section_sizes(probe, expect);
        for (int i = 0; i < FM_SECTIONS; ++i) if (expect[i] != hd.bytes[i]) { fm_set_error("%s: corrupt section table", path); rc = FMIDX_EIO; }
        void *dst[FM_SECTIONS] = {nullptr};
        const size_t CH = 64u << 20;
        std::vector<unsigned char> buf(CH);
        for (int sct = 0; sct < FM_SECTIONS && rc == FMIDX_OK; ++sct) {
            if (!hd.bytes[sct]) continue;
            rc = fm_dev_alloc(h, &dst[sct], hd.bytes[sct], true);
            for (uint64_t off = 0; off < hd.bytes[sct] && rc == FMIDX_OK; off += CH) {
                size_t len = (size_t)std::min<uint64_t>(CH, hd.bytes[sct] - off);
                if (fread(buf.data(), 1, len, f) != len) { fm_set_error("%s: truncated file", path); rc = FMIDX_EIO; break; }
                cudaError_t e = cudaMemcpy((char *)dst[sct] + off, buf.data(), len, cudaMemcpyHostToDevice);
                if (e != cudaSuccess) { fm_set_error("H2D failed while loading: %s", cudaGetErrorString(e)); rc = FMIDX_ECUDA; }
            }
        }
        if (rc == FMIDX_OK) {
            d.cells = (const uint4 *)dst[0]; d.C = (const uint32_t *)dst[1]; d.D = (const uint32_t *)dst[2];
            d.sa_marks = (const uint2 *)dst[3]; d.sa_samples = (const uint32_t *)dst[4];
            d.isa_samples = (const uint32_t *)dst[5]; d.text = (const uint32_t *)dst[6]; d.bcells = (const uint4 *)dst[7]; d.bwt = (const uint32_t *)dst[8];
            h->dev = d;
            h->opt_sa_dens = d.sa_dens; h->opt_isa_dens = d.isa_dens; h->opt_keep_text = hd.keep_text;
            h->ready = true;
            rc = fm_build_unigram_table(h, h->stream);  // derived data: rebuilt rather than stored
        }
    }
    fclose(f);
    if (rc != FMIDX_OK) { fmidx_free(h); return rc; }
    *out = h;
    return FMIDX_OK;
}

// ---- scalar queries -----------------------------------------------------------------------------------
int fmidx_size(fmidx_t *h, uint64_t *out) {
    if (!h || !out) { fm_set_error("null argument"); return FMIDX_EINVAL; }
    if (!h->ready) { fm_set_error("index is not initialised"); return FMIDX_ESTATE; }
    *out = h->dev.n;
    return FMIDX_OK;
}

int fmidx_info(fmidx_t *h, uint64_t info[8]) {
    if (!h || !info) { fm_set_error("null argument"); return FMIDX_EINVAL; }
    if (!h->ready) { fm_set_error("index is not initialised"); return FMIDX_ESTATE; }
    info[0] = h->dev.n; info[1] = h->dev.levels + 1; info[2] = h->dev.max_sym; info[3] = h->dev.sa_dens;
    info[4] = h->dev.isa_dens; info[5] = h->dev_bytes; info[6] = (uint64_t)h->dev.nblocks * 128; info[7] = h->dev.text ? 1 : 0;
    return FMIDX_OK;
}

int fmidx_gather_peak(fmidx_t *h, int unit_bytes, size_t n_loads, double *out_gbps) {
    FM_TRY(enter(h, true));
    return fm_launch_gather_probe(h, unit_bytes, n_loads, out_gbps, h->stream);
}

int fmidx_debug_copy_sa(fmidx_t *h, uint32_t *out, size_t cap) {
    FM_TRY(enter(h, true));
    if (h->dev.sa_dens != 1) { fm_set_error("debug_copy_sa needs sa_dens == 1"); return FMIDX_ESTATE; }
    if (cap < h->dev.n) { fm_set_error("buffer too small"); return FMIDX_ERANGE; }
    FM_CUDA(cudaMemcpy(out, h->dev.sa_samples, (size_t)h->dev.n * 4, cudaMemcpyDeviceToHost));
    return FMIDX_OK;
}

int fmidx_backward_search_step(fmidx_t *h, uint64_t symbol, uint64_t low, uint64_t high, uint64_t out[2]) {
    FM_TRY(enter(h, true));
    if (h->tune_scalar_fast) return fm_scalar_step(h, symbol, low, high, out);
    FM_TRY(fm_scratch(h, h->d_scratch[0], 64, false));
    FM_TRY(fm_scratch(h, h->h_scratch[0], 64, true));
    uint64_t *hp = (uint64_t *)h->h_scratch[0].p, *dp = (uint64_t *)h->d_scratch[0].p;
    hp[0] = symbol; hp[1] = low; hp[2] = high;
    FM_CUDA(cudaMemcpyAsync(dp, hp, 24, cudaMemcpyHostToDevice, h->stream));
    FM_TRY(fm_launch_step(h, dp, dp + 1, dp + 2, 1, dp + 4, dp + 5, h->stream));
    FM_CUDA(cudaMemcpyAsync(hp + 4, dp + 4, 16, cudaMemcpyDeviceToHost, h->stream));
    FM_CUDA(cudaStreamSynchronize(h->stream));
    out[0] = hp[4]; out[1] = hp[5];
    return FMIDX_OK;
}

int fmidx_backward_search_multi(fmidx_t *h, const uint64_t *query, size_t n, uint64_t out[2]) {
    FM_TRY(enter(h, true));
    if (h->tune_scalar_fast && n <= (size_t)fm_scalar_max_tokens()) {
        int64_t t[128];
        for (size_t i = 0; i < n; ++i) t[i] = query[i] > 0x7FFFFFFFull ? -1 : (int64_t)query[i];  // -1: absent
        return fm_scalar_search(h, t, (int)n, 0, 1, out);
    }
    const size_t tb = (n + 4) * 4;
    FM_TRY(fm_scratch(h, h->d_scratch[0], tb + 64, false));
    FM_TRY(fm_scratch(h, h->h_scratch[0], tb + 64, true));
    char *hp = (char *)h->h_scratch[0].p, *dp = (char *)h->d_scratch[0].p;
    int64_t *hoffs = (int64_t *)hp;             // [2]
    uint64_t *hres = (uint64_t *)(hp + 16);     // [2]
    int32_t *htok = (int32_t *)(hp + 32);       // [n]
    hoffs[0] = 0; hoffs[1] = (int64_t)n;
    for (size_t i = 0; i < n; ++i) htok[i] = query[i] > 0x7FFFFFFFull ? -1 : (int32_t)query[i];  // -1: absent
    FM_CUDA(cudaMemcpyAsync(dp, hp, 32 + n * 4, cudaMemcpyHostToDevice, h->stream));
    FM_TRY(fm_launch_search(h, (int32_t *)(dp + 32), (int64_t *)dp, 1, 0, (uint64_t *)(dp + 16), (uint64_t *)(dp + 24), 1, h->stream));
    FM_CUDA(cudaMemcpyAsync(hres, dp + 16, 16, cudaMemcpyDeviceToHost, h->stream));
    FM_CUDA(cudaStreamSynchronize(h->stream));
    out[0] = hres[0]; out[1] = hres[1];
    return FMIDX_OK;
}

// ranges -> per-range sparse (symbol,count) lists on the raw symbol axis, via bitmap + compaction
static int distinct_lists(fmidx *h, const uint64_t *lows, const uint64_t *highs, size_t nr, bool with_counts,
                          std::vector<uint64_t> &flat, std::vector<uint64_t> &offsets) {
    const uint32_t vocab = h->dev.max_sym + 1;
    const uint32_t words = (vocab + 31) / 32;
    offsets.assign(nr + 1, 0);
    flat.clear();
    size_t chunk = (size_t)((256ull << 20) / ((uint64_t)vocab * 4));
    if (chunk < 1) chunk = 1;
    if (chunk > 16384) chunk = 16384;
    cudaStream_t s = h->stream;
    for (size_t r0 = 0; r0 < nr; r0 += chunk) {
        const size_t rn = std::min(chunk, nr - r0);
        FM_TRY(fm_scratch(h, h->d_scratch[1], rn * 16, false));                       // sp | ep
        FM_TRY(fm_scratch(h, h->d_scratch[2], rn * (size_t)words * 4, false));         // bitmaps
        FM_TRY(fm_scratch(h, h->d_scratch[3], rn * (size_t)vocab * 4, false));         // counts
        FM_TRY(fm_scratch(h, h->d_scratch[4], (rn + 1) * 8 + rn * 4, false));          // offsets | n
        uint64_t *d_sp = (uint64_t *)h->d_scratch[1].p, *d_ep = d_sp + rn;
        uint32_t *d_mask = (uint32_t *)h->d_scratch[2].p, *d_counts = (uint32_t *)h->d_scratch[3].p;
        uint64_t *d_off = (uint64_t *)h->d_scratch[4].p;
        uint32_t *d_n = (uint32_t *)(d_off + rn + 1);
        FM_CUDA(cudaMemcpyAsync(d_sp, lows + r0, rn * 8, cudaMemcpyHostToDevice, s));
        FM_CUDA(cudaMemcpyAsync(d_ep, highs + r0, rn * 8, cudaMemcpyHostToDevice, s));
        FM_TRY(fm_launch_mask(h, d_sp, d_ep, rn, 0, vocab, nullptr, nullptr, 3, 0, nullptr, -1, nullptr, 0, d_mask,
                              with_counts ? d_counts : nullptr, s));
        FM_TRY(fm_launch_popcount(h, d_mask, rn, vocab, d_n, s));
        std::vector<uint32_t> hn(rn);
        FM_CUDA(cudaMemcpyAsync(hn.data(), d_n, rn * 4, cudaMemcpyDeviceToHost, s));
        FM_CUDA(cudaStreamSynchronize(s));
        std::vector<uint64_t> hoff(rn + 1, 0);
        for (size_t i = 0; i < rn; ++i) hoff[i + 1] = hoff[i] + hn[i];
        const uint64_t total = hoff[rn];
        FM_TRY(fm_scratch(h, h->d_scratch[5], (total + 1) * 16, false));
        int64_t *d_tok = (int64_t *)h->d_scratch[5].p;
        uint64_t *d_cnt = (uint64_t *)(d_tok + total + 1);
        FM_CUDA(cudaMemcpyAsync(d_off, hoff.data(), (rn + 1) * 8, cudaMemcpyHostToDevice, s));
        FM_TRY(fm_launch_compact(h, d_mask, with_counts ? d_counts : nullptr, rn, vocab, d_off, d_tok,
                                 with_counts ? d_cnt : nullptr, s));
        std::vector<int64_t> htok(total);
        std::vector<uint64_t> hcnt(with_counts ? total : 0);
        if (total) {
            FM_CUDA(cudaMemcpyAsync(htok.data(), d_tok, total * 8, cudaMemcpyDeviceToHost, s));
            if (with_counts) FM_CUDA(cudaMemcpyAsync(hcnt.data(), d_cnt, total * 8, cudaMemcpyDeviceToHost, s));
        }
        FM_CUDA(cudaStreamSynchronize(s));
        for (size_t i = 0; i < rn; ++i) {
            for (uint64_t k = hoff[i]; k < hoff[i + 1]; ++k) {
                flat.push_back((uint64_t)htok[k]);
                if (with_counts) flat.push_back(hcnt[k]);
            }
            offsets[r0 + i + 1] = flat.size();
        }
    }
    return FMIDX_OK;
}

static int distinct_one(fmidx_t *h, uint64_t low, uint64_t high, bool with_counts, uint64_t *out, size_t cap, size_t *k) {
    FM_TRY(enter(h, true));
    if (!k) { fm_set_error("k is NULL"); return FMIDX_EINVAL; }
    if (h->tune_scalar_fast) {
        bool fallback = false;
        int rc = fm_scalar_distinct(h, low, high, with_counts ? 1 : 0, out, cap, k, &fallback);
        if (rc != FMIDX_OK || !fallback) return rc;
    }
    std::vector<uint64_t> flat, offs;
    FM_TRY(distinct_lists(h, &low, &high, 1, with_counts, flat, offs));
    *k = flat.size();
    if (flat.size() > cap) { fm_set_error("output buffer too small: need %zu", flat.size()); return FMIDX_ERANGE; }
    if (!flat.empty()) memcpy(out, flat.data(), flat.size() * 8);
    return FMIDX_OK;
}

int fmidx_distinct(fmidx_t *h, uint64_t low, uint64_t high, uint64_t *out, size_t cap, size_t *k) {
    return distinct_one(h, low, high, false, out, cap, k);
}
int fmidx_distinct_count(fmidx_t *h, uint64_t low, uint64_t high, uint64_t *out, size_t cap, size_t *k) {
    return distinct_one(h, low, high, true, out, cap, k);
}

int fmidx_distinct_count_multi(fmidx_t *h, const uint64_t *lows, const uint64_t *highs, size_t nr, uint64_t *out,
                               size_t cap, uint64_t *offsets) {
    FM_TRY(enter(h, true));
    if (!offsets || (nr && (!lows || !highs))) { fm_set_error("null argument"); return FMIDX_EINVAL; }
    std::vector<uint64_t> flat, offs;
    FM_TRY(distinct_lists(h, lows, highs, nr, true, flat, offs));
    memcpy(offsets, offs.data(), (nr + 1) * 8);
    if (flat.size() > cap) { fm_set_error("output buffer too small: need %zu", flat.size()); return FMIDX_ERANGE; }
    if (!flat.empty()) memcpy(out, flat.data(), flat.size() * 8);
    return FMIDX_OK;
}

int fmidx_locate(fmidx_t *h, uint64_t row, uint64_t *out) {
    FM_TRY(enter(h, true));
    if (!out) { fm_set_error("null argument"); return FMIDX_EINVAL; }
    if (h->tune_scalar_fast) return fm_scalar_locate(h, row, out);
    FM_TRY(fm_scratch(h, h->d_scratch[0], 64, false));
    FM_TRY(fm_scratch(h, h->h_scratch[0], 64, true));
    uint64_t *hp = (uint64_t *)h->h_scratch[0].p, *dp = (uint64_t *)h->d_scratch[0].p;
    hp[0] = row;
    FM_CUDA(cudaMemcpyAsync(dp, hp, 8, cudaMemcpyHostToDevice, h->stream));
    FM_TRY(fm_launch_locate(h, dp, 1, dp + 1, h->stream));
    FM_CUDA(cudaMemcpyAsync(hp + 1, dp + 1, 8, cudaMemcpyDeviceToHost, h->stream));
    FM_CUDA(cudaStreamSynchronize(h->stream));
    *out = hp[1];
    return FMIDX_OK;
}

int fmidx_extract_text(fmidx_t *h, uint64_t begin, uint64_t end, uint64_t *out, size_t cap, size_t *k) {
    FM_TRY(enter(h, true));
    if (!k) { fm_set_error("k is NULL"); return FMIDX_EINVAL; }
    uint64_t e2 = end > h->dev.n ? h->dev.n : end;
    uint64_t m = e2 > begin ? e2 - begin : 0;
    *k = (size_t)m;
    if (m == 0) return FMIDX_OK;
    if (m > cap) { fm_set_error("output buffer too small: need %llu", (unsigned long long)m); return FMIDX_ERANGE; }
    FM_TRY(fm_scratch(h, h->d_scratch[0], 64 + m * 8, false));
    FM_TRY(fm_scratch(h, h->h_scratch[0], 64, true));
    uint64_t *hp = (uint64_t *)h->h_scratch[0].p, *dp = (uint64_t *)h->d_scratch[0].p;
    hp[0] = begin; hp[1] = end;
    FM_CUDA(cudaMemcpyAsync(dp, hp, 16, cudaMemcpyHostToDevice, h->stream));
    FM_TRY(fm_launch_extract(h, dp, dp + 1, 1, m, dp + 8, dp + 2, h->stream));
    FM_CUDA(cudaMemcpyAsync(out, dp + 8, m * 8, cudaMemcpyDeviceToHost, h->stream));
    FM_CUDA(cudaStreamSynchronize(h->stream));
    return FMIDX_OK;
}

// ---- batched device entry points ------------------------------------------------------------------------
int fmidx_search_batch(fmidx_t *h, const int32_t *d_toks, const int64_t *d_offs, size_t B, int64_t shift,
                       uint64_t *d_sp, uint64_t *d_ep, void *stream) {
    FM_TRY(enter(h, true));
    return fm_launch_search(h, d_toks, d_offs, B, shift, d_sp, d_ep, 0, (cudaStream_t)stream);
}

int fmidx_step_batch(fmidx_t *h, const uint64_t *d_symbols, const uint64_t *d_lows, const uint64_t *d_highs, size_t B,
                     uint64_t *d_sp, uint64_t *d_ep, void *stream) {
    FM_TRY(enter(h, true));
    return fm_launch_step(h, d_symbols, d_lows, d_highs, B, d_sp, d_ep, (cudaStream_t)stream);
}

int fmidx_decode_step(fmidx_t *h, const int64_t *d_last_tok, int64_t tok_stride, const int32_t *d_parent, const uint64_t *d_state_sp_in,
                      const uint64_t *d_state_ep_in, uint64_t *d_state_sp_out, uint64_t *d_state_ep_out, size_t B, int64_t shift,
                      int64_t eos, int64_t pad, int64_t stop_at_count, uint64_t *d_query_sp, uint64_t *d_query_ep,
                      int32_t *d_override_tok, void *stream) {
    FM_TRY(enter(h, true));
    if (!d_last_tok || !d_state_sp_in || !d_state_ep_in || !d_state_sp_out || !d_state_ep_out || !d_query_sp || !d_query_ep || !d_override_tok) {
        fm_set_error("null argument");
        return FMIDX_EINVAL;
    }
    if (d_state_sp_in == d_state_sp_out || d_state_ep_in == d_state_ep_out) { fm_set_error("decode_step: the state is double-buffered (rows are gathered by parent)"); return FMIDX_EINVAL; }
    return fm_launch_decode_step(h, d_last_tok, tok_stride, d_parent, d_state_sp_in, d_state_ep_in, d_state_sp_out, d_state_ep_out, B,
                                 shift, eos, pad, stop_at_count, d_query_sp, d_query_ep, d_override_tok, (cudaStream_t)stream);
}

int fmidx_distinct_mask_batch(fmidx_t *h, const uint64_t *d_sp, const uint64_t *d_ep, size_t B, int64_t shift,
                              uint32_t vocab, uint32_t *d_mask, uint32_t *d_counts, void *stream) {
    FM_TRY(enter(h, true));
    if (!d_mask) { fm_set_error("d_mask is NULL"); return FMIDX_EINVAL; }
    return fm_launch_mask(h, d_sp, d_ep, B, shift, vocab, nullptr, nullptr, 3, 0, nullptr, -1, nullptr, 0, d_mask, d_counts,
                          (cudaStream_t)stream);
}

int fmidx_query_batch(fmidx_t *h, int ws_slot, const int32_t *d_toks, const int64_t *d_offs, size_t B, int64_t shift,
                      uint32_t vocab, uint64_t *d_sp, uint64_t *d_ep, uint32_t *d_mask, void *stream) {
    FM_TRY(enter(h, true));
    if (ws_slot < 0 || ws_slot >= FMIDX_SLOTS) { fm_set_error("ws_slot must be 0..%d", FMIDX_SLOTS - 1); return FMIDX_EINVAL; }
    if (!d_sp || !d_ep) { fm_set_error("null argument"); return FMIDX_EINVAL; }
    // the search kernel also zeroes the mask rows (its stores are free: it is bound by its dependent-load chain)
    const bool kz = vocab != 0 && fm_search_can_zero(d_mask, vocab);
    // probe_phase (measurement only, benchmarks/overlap_probe.py): 1 = search (+ row zeroing) alone, 2 = enumeration alone on the
    // ranges already in d_sp / d_ep
    if (h->tune_probe_phase != 2)
    FM_TRY(fm_launch_search(h, d_toks, d_offs, B, shift, d_sp, d_ep, 0, (cudaStream_t)stream, kz ? d_mask : nullptr, (vocab + 31u) >> 5));
    if (!d_mask || vocab == 0 || h->tune_probe_phase == 1) return FMIDX_OK;
    h->mask_ws_override = &h->slots[ws_slot].ws;  // calls on different streams must not share the node lists
    struct Reset { fmidx_t *h; ~Reset() { h->mask_ws_override = nullptr; } } reset{h};
    // fmidx_tune("lazy_clear", 2): the caller promises that nobody else writes d_mask between this slot's calls -- then only the
    // slices the previous call dirtied are zeroed
    uint32_t *occ_state = nullptr;
    FM_TRY(lazy_state(h, h->slots[ws_slot], d_mask, B, vocab, false, (cudaStream_t)stream, &occ_state));
    return fm_launch_mask(h, d_sp, d_ep, B, shift, vocab, nullptr, nullptr, 3, 0, nullptr, -1, nullptr, 0, d_mask, nullptr,
                          (cudaStream_t)stream, nullptr, 0, ~(size_t)0, kz, nullptr, nullptr, kz ? nullptr : occ_state);
}

int fmidx_distinct_mask_batch_occ(fmidx_t *h, const uint64_t *d_sp, const uint64_t *d_ep, size_t B, int64_t shift,
                                  uint32_t vocab, uint32_t *d_mask, uint32_t *d_counts, uint32_t *d_occ, void *stream) {
    FM_TRY(enter(h, true));
    if (!d_mask || !d_occ) { fm_set_error("null argument"); return FMIDX_EINVAL; }
    return fm_launch_mask(h, d_sp, d_ep, B, shift, vocab, nullptr, nullptr, 3, 0, nullptr, -1, nullptr, 0, d_mask, d_counts,
                          (cudaStream_t)stream, d_occ);
}

// ---- peer-visible buffers (CUDA IPC) for the pull-style cross-shard exchange ------------------------------
int fmidx_ipc_alloc(fmidx_t *h, size_t bytes, void **d_ptr, unsigned char handle[64]) {
    FM_TRY(enter(h, false));
    if (!d_ptr || !handle) { fm_set_error("null argument"); return FMIDX_EINVAL; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    FM_CUDA(cudaMalloc(d_ptr, bytes ? bytes : 16));
    cudaIpcMemHandle_t hd;
    cudaError_t e = cudaIpcGetMemHandle(&hd, *d_ptr);
    if (e != cudaSuccess) {
        cudaFree(*d_ptr);
        *d_ptr = nullptr;
        fm_set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        return FMIDX_ECUDA;
    }
    memcpy(handle, &hd, 64);
    return FMIDX_OK;
}

int fmidx_ipc_open(fmidx_t *h, const unsigned char handle[64], void **d_ptr) {
    FM_TRY(enter(h, false));
    if (!d_ptr || !handle) { fm_set_error("null argument"); return FMIDX_EINVAL; }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle, 64);
    FM_CUDA(cudaIpcOpenMemHandle(d_ptr, hd, cudaIpcMemLazyEnablePeerAccess));
    return FMIDX_OK;
}

int fmidx_ipc_close(fmidx_t *h, void *d_ptr) {
    FM_TRY(enter(h, false));
    FM_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return FMIDX_OK;
}

int fmidx_ipc_free(fmidx_t *h, void *d_ptr) {
    FM_TRY(enter(h, false));
    FM_CUDA(cudaFree(d_ptr));
    return FMIDX_OK;
}

int fmidx_mask_logits(fmidx_t *h, const uint64_t *d_sp, const uint64_t *d_ep, size_t B, int64_t shift, uint32_t vocab,
                      const void *d_logits, void *d_out, int dtype, int64_t row_stride, const int32_t *d_override_tok,
                      int32_t always_allow, const uint32_t *d_extra_masks, uint32_t n_extra, uint32_t *d_mask_out,
                      void *stream) {
    FM_TRY(enter(h, true));
    if (!d_logits || !d_out) { fm_set_error("logits pointers are NULL"); return FMIDX_EINVAL; }
    if (dtype != FMIDX_F32 && dtype != FMIDX_BF16 && dtype != FMIDX_F16) { fm_set_error("unknown dtype %d", dtype); return FMIDX_EINVAL; }
    return fm_launch_mask(h, d_sp, d_ep, B, shift, vocab, d_logits, d_out, dtype, row_stride, d_override_tok, always_allow,
                          d_extra_masks, n_extra, d_mask_out, nullptr, (cudaStream_t)stream);
}

int fmidx_boost_logits(fmidx_t *h, const void *d_logits, void *d_out, int dtype, int64_t row_stride, size_t B, uint32_t vocab,
                       const uint32_t *d_dmask, const uint32_t *d_n_distinct, const uint32_t *d_unigram, const int32_t *d_kind,
                       const float *d_vd, const float *d_vu_nonempty, const float *d_vu_empty, const float *d_ve, int32_t eos,
                       void *stream) {
    FM_TRY(enter(h, false));
    if (!d_logits || !d_out || !d_dmask || !d_n_distinct || !d_unigram || !d_kind || !d_vd || !d_vu_nonempty || !d_vu_empty || !d_ve) {
        fm_set_error("null argument");
        return FMIDX_EINVAL;
    }
    return fm_launch_boost_apply(h, d_logits, d_out, dtype, row_stride, B, vocab, d_dmask, d_n_distinct, d_unigram, d_kind, d_vd,
                                 d_vu_nonempty, d_vu_empty, d_ve, eos, (cudaStream_t)stream);
}

int fmidx_mask_or(fmidx_t *h, const uint32_t *d_pieces, uint32_t n_pieces, size_t n_words, uint32_t *d_out, void *stream) {
    FM_TRY(enter(h, false));
    if (!d_pieces || !d_out) { fm_set_error("null argument"); return FMIDX_EINVAL; }
    return fm_launch_or(h, d_pieces, n_pieces, n_words, d_out, (cudaStream_t)stream);
}

int fmidx_mask_popcount(fmidx_t *h, const uint32_t *d_mask, size_t B, uint32_t vocab, uint32_t *d_n, void *stream) {
    FM_TRY(enter(h, false));
    return fm_launch_popcount(h, d_mask, B, vocab, d_n, (cudaStream_t)stream);
}

int fmidx_mask_compact(fmidx_t *h, const uint32_t *d_mask, const uint32_t *d_counts, size_t B, uint32_t vocab,
                       const uint64_t *d_list_offsets, int64_t *d_tokens, uint64_t *d_cnts, void *stream) {
    FM_TRY(enter(h, false));
    return fm_launch_compact(h, d_mask, d_counts, B, vocab, d_list_offsets, d_tokens, d_cnts, (cudaStream_t)stream);
}

int fmidx_locate_batch(fmidx_t *h, const uint64_t *d_rows, size_t n, uint64_t *d_pos, void *stream) {
    FM_TRY(enter(h, true));
    return fm_launch_locate(h, d_rows, n, d_pos, (cudaStream_t)stream);
}

int fmidx_extract_batch(fmidx_t *h, const uint64_t *d_begin, const uint64_t *d_end, size_t n, uint64_t max_len,
                        uint64_t *d_out, uint64_t *d_len, void *stream) {
    FM_TRY(enter(h, true));
    if (max_len == 0) { fm_set_error("max_len is 0"); return FMIDX_EINVAL; }
    return fm_launch_extract(h, d_begin, d_end, n, max_len, d_out, d_len, (cudaStream_t)stream);
}

// ---- host-buffer batch calls ----------------------------------------------------------------------------
static int query_host_impl(fmidx_t *h, const int32_t *toks, const int64_t *offs, size_t B, int64_t shift, uint32_t vocab,
                           uint64_t *sp, uint64_t *ep, uint32_t *mask, uint32_t *n_distinct, int slot = -1) {
    FM_TRY(enter(h, true));
    if (B == 0) return FMIDX_OK;
    if (!toks || !offs || !sp || !ep) { fm_set_error("null argument"); return FMIDX_EINVAL; }
    if (offs[0] != 0) { fm_set_error("offs[0] must be 0"); return FMIDX_EINVAL; }
    const size_t ntok = (size_t)offs[B];
    const uint32_t words = (vocab + 31) / 32;
    const bool want_mask = mask || n_distinct;
    // slot < 0: synchronous call on the handle's private stream; slot >= 0: one of the pipelined slots
    // (own stream, own buffers, own mask workspace), left running -- fmidx_query_wait collects it
    cudaStream_t s = h->stream;
    FmBuf *in = &h->d_scratch[6], *mb = &h->d_scratch[7];
    if (slot >= 0) {
        fmidx::Slot &sl = h->slots[slot];
        if (!sl.stream) FM_CUDA(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
        s = sl.stream;
        in = &sl.in;
        mb = &sl.mask;
        h->mask_ws_override = &sl.ws;
    }
    struct Reset { fmidx_t *h; ~Reset() { h->mask_ws_override = nullptr; } } reset{h};
    FM_TRY(fm_scratch(h, *in, ntok * 4 + 16 + (B + 1) * 8 + B * 16 + B * 4, false));
    char *dp = (char *)in->p;
    int64_t *d_offs = (int64_t *)dp;
    uint64_t *d_sp = (uint64_t *)(dp + (B + 1) * 8), *d_ep = d_sp + B;
    uint32_t *d_n = (uint32_t *)(d_ep + B);
    int32_t *d_toks = (int32_t *)(d_n + B);
    FM_CUDA(cudaMemcpyAsync(d_offs, offs, (B + 1) * 8, cudaMemcpyHostToDevice, s));
    if (ntok) FM_CUDA(cudaMemcpyAsync(d_toks, toks, ntok * 4, cudaMemcpyHostToDevice, s));
    uint32_t *d_mask = nullptr;
    if (want_mask) {
        FM_TRY(fm_scratch(h, *mb, B * (size_t)words * 4, false));
        d_mask = (uint32_t *)mb->p;
    }
    const bool kz = want_mask && fm_search_can_zero(d_mask, vocab);
    FM_TRY(fm_launch_search(h, d_toks, d_offs, B, shift, d_sp, d_ep, 0, s, kz ? d_mask : nullptr, words));
    if (want_mask) {
        // the distinct-token counts come out of the enumeration itself (rows it finishes by scanning count their newly set bits;
        // only the few wide rows are re-read by k_popcount_pending) instead of a popcount pass over all B x 16 KB
        uint32_t *occ_state = nullptr;  // pipelined slots own their mask buffer: lazy clearing
        if (slot >= 0 && !kz) FM_TRY(lazy_state(h, h->slots[slot], d_mask, B, vocab, true, s, &occ_state));
        FM_TRY(fm_launch_mask(h, d_sp, d_ep, B, shift, vocab, nullptr, nullptr, 3, 0, nullptr, -1, nullptr, 0, d_mask, nullptr, s,
                              nullptr, 0, ~(size_t)0, kz, nullptr, n_distinct ? d_n : nullptr, occ_state));
    }
    FM_CUDA(cudaMemcpyAsync(sp, d_sp, B * 8, cudaMemcpyDeviceToHost, s));
    FM_CUDA(cudaMemcpyAsync(ep, d_ep, B * 8, cudaMemcpyDeviceToHost, s));
    if (mask) FM_CUDA(cudaMemcpyAsync(mask, d_mask, B * (size_t)words * 4, cudaMemcpyDeviceToHost, s));
    if (n_distinct) FM_CUDA(cudaMemcpyAsync(n_distinct, d_n, B * 4, cudaMemcpyDeviceToHost, s));
    if (slot < 0) FM_CUDA(cudaStreamSynchronize(s));
    return FMIDX_OK;
}

// ---- graph-captured pipelined form -------------------------------------------------------------------------------------
static int graph_sequence(fmidx_t *h, fmidx::Slot &sl, int64_t shift, uint32_t vocab) {
    const size_t B = sl.g_B;
    cudaStream_t s = sl.stream;
    const uint32_t words = (vocab + 31) / 32;
    h->mask_ws_override = &sl.ws;
    struct Reset { fmidx_t *h; ~Reset() { h->mask_ws_override = nullptr; } } reset{h};
    if (sl.g_dev_mask) {  // device-resident form: the caller's buffers, no copies
        const bool kz = fm_search_can_zero(sl.g_dev_mask, vocab);
        if (h->tune_probe_phase != 2)  // probe_phase: measurement only (see fmidx_query_batch)
        FM_TRY(fm_launch_search(h, sl.g_dev_toks, sl.g_dev_offs, B, shift, sl.g_dev_sp, sl.g_dev_ep, 0, s, kz ? sl.g_dev_mask : nullptr, words));
        if (h->tune_probe_phase == 1) return FMIDX_OK;
        return fm_launch_mask(h, sl.g_dev_sp, sl.g_dev_ep, B, shift, vocab, nullptr, nullptr, 3, 0, nullptr, -1, nullptr, 0, sl.g_dev_mask,
                              nullptr, s, nullptr, 0, ~(size_t)0, kz, nullptr, nullptr,
                              (!kz && sl.lazy_ptr == sl.g_dev_mask) ? (uint32_t *)sl.occ.p : nullptr);
    }
    int64_t *d_offs = (int64_t *)sl.g_din;
    int32_t *d_toks = (int32_t *)((char *)sl.g_din + (B + 1) * 8);
    uint64_t *d_sp = (uint64_t *)sl.g_dout, *d_ep = d_sp + B;
    uint32_t *d_n = (uint32_t *)(d_ep + B);
    uint32_t *d_mask = (uint32_t *)sl.mask.p;
    FM_CUDA(cudaMemcpyAsync(sl.g_din, sl.g_hin, (B + 1) * 8 + sl.g_maxtok * 4, cudaMemcpyHostToDevice, s));
    // the search kernel also zeroes the mask rows (as in fmidx_query_batch)
    const bool kz = fm_search_can_zero(d_mask, vocab);
    FM_TRY(fm_launch_search(h, d_toks, d_offs, B, shift, d_sp, d_ep, 0, s, kz ? d_mask : nullptr, words));
    FM_TRY(fm_launch_mask(h, d_sp, d_ep, B, shift, vocab, nullptr, nullptr, 3, 0, nullptr, -1, nullptr, 0, d_mask, nullptr, s, nullptr, 0,
                          ~(size_t)0, kz, nullptr, d_n, (!kz && sl.lazy_ptr == d_mask) ? (uint32_t *)sl.occ.p : nullptr));
    FM_CUDA(cudaMemcpyAsync(sl.g_hout, sl.g_dout, B * 20, cudaMemcpyDeviceToHost, s));
    return FMIDX_OK;
}

// eager sizing pass + the two captures (the mask kernels alternate between two work-list counter pairs)
static int graph_capture(fmidx_t *h, fmidx::Slot &sl, int64_t shift, uint32_t vocab) {
    // one eager pass sizes every workspace (nothing may be allocated while a stream is being captured)
    FM_TRY(graph_sequence(h, sl, shift, vocab));
    FM_CUDA(cudaStreamSynchronize(sl.stream));
    for (int e = 0; e < 2; ++e) {
        const uint64_t l0 = g_fm_launches.load();
        FM_CUDA(cudaStreamBeginCapture(sl.stream, cudaStreamCaptureModeThreadLocal));
        int rc = graph_sequence(h, sl, shift, vocab);
        cudaGraph_t g = nullptr;
        cudaError_t ce = cudaStreamEndCapture(sl.stream, &g);
        if (rc != FMIDX_OK) { if (g) cudaGraphDestroy(g); return rc; }
        if (ce != cudaSuccess) { fm_set_error("stream capture failed: %s", cudaGetErrorString(ce)); return FMIDX_ECUDA; }
        ce = cudaGraphInstantiate(&sl.gexec[e], g, 0);
        cudaGraphDestroy(g);
        if (ce != cudaSuccess) { fm_set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ce)); return FMIDX_ECUDA; }
        sl.g_launches = g_fm_launches.load() - l0;
        g_fm_launches.store(l0);  // captured, not launched
    }
    sl.g_step = 0;
    return FMIDX_OK;
}

int fmidx_query_graph_prepare(fmidx_t *h, int slot, size_t B, size_t max_tokens, int64_t shift, uint32_t vocab, void **h_in,
                              void **h_out) {
    FM_TRY(enter(h, true));
    if (slot < 0 || slot >= FMIDX_SLOTS) { fm_set_error("slot must be 0..%d", FMIDX_SLOTS - 1); return FMIDX_EINVAL; }
    if (!B || !vocab || !h_in || !h_out) { fm_set_error("bad argument"); return FMIDX_EINVAL; }
    fmidx::Slot &sl = h->slots[slot];
    if (sl.gexec[0]) { fm_set_error("slot %d already holds a captured step", slot); return FMIDX_ESTATE; }
    if (!sl.stream) FM_CUDA(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
    FM_CUDA(cudaStreamSynchronize(sl.stream));
    sl.g_B = B;
    sl.g_maxtok = max_tokens;
    const size_t in_bytes = (B + 1) * 8 + max_tokens * 4;
    FM_CUDA(cudaMallocHost(&sl.g_hin, in_bytes));
    FM_CUDA(cudaMallocHost(&sl.g_hout, B * 20));
    FM_CUDA(cudaMalloc(&sl.g_din, in_bytes + 16));
    FM_CUDA(cudaMalloc(&sl.g_dout, B * 20 + 16));
    memset(sl.g_hin, 0, in_bytes);  // B empty prefixes: a valid input
    FM_TRY(fm_scratch(h, sl.mask, B * (size_t)((vocab + 31) / 32) * 4, false));
    uint32_t *occ_state = nullptr;  // the slot owns its mask buffer: lazy clearing, decided here and baked into the captures
    FM_TRY(lazy_state(h, sl, (const uint32_t *)sl.mask.p, B, vocab, true, sl.stream, &occ_state));
    FM_TRY(graph_capture(h, sl, shift, vocab));
    *h_in = sl.g_hin;
    *h_out = sl.g_hout;
    return FMIDX_OK;
}

int fmidx_query_graph_prepare_device(fmidx_t *h, int slot, const int32_t *d_toks, const int64_t *d_offs, size_t B, int64_t shift,
                                     uint32_t vocab, uint64_t *d_sp, uint64_t *d_ep, uint32_t *d_mask) {
    FM_TRY(enter(h, true));
    if (slot < 0 || slot >= FMIDX_SLOTS) { fm_set_error("slot must be 0..%d", FMIDX_SLOTS - 1); return FMIDX_EINVAL; }
    if (!B || !vocab || !d_toks || !d_offs || !d_sp || !d_ep || !d_mask) { fm_set_error("bad argument"); return FMIDX_EINVAL; }
    fmidx::Slot &sl = h->slots[slot];
    if (sl.gexec[0]) { fm_set_error("slot %d already holds a captured step", slot); return FMIDX_ESTATE; }
    if (!sl.stream) FM_CUDA(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
    FM_CUDA(cudaStreamSynchronize(sl.stream));
    sl.g_B = B;
    sl.g_dev_toks = d_toks; sl.g_dev_offs = d_offs; sl.g_dev_sp = d_sp; sl.g_dev_ep = d_ep; sl.g_dev_mask = d_mask;
    uint32_t *occ_state = nullptr;  // the caller's buffer: lazy clearing only on request (fmidx_tune "lazy_clear" 2)
    FM_TRY(lazy_state(h, sl, d_mask, B, vocab, false, sl.stream, &occ_state));
    int rc = graph_capture(h, sl, shift, vocab);
    if (rc != FMIDX_OK) sl.g_dev_mask = nullptr;
    return rc;
}

int fmidx_query_graph_launch_on(fmidx_t *h, int slot, void *stream) {
    if (!h) { fm_set_error("null handle"); return FMIDX_EINVAL; }
    if (slot < 0 || slot >= FMIDX_SLOTS || !h->slots[slot].gexec[0]) { fm_set_error("slot %d holds no captured step", slot); return FMIDX_ESTATE; }
    fmidx::Slot &sl = h->slots[slot];
    FM_CUDA(cudaGraphLaunch(sl.gexec[sl.g_step++ & 1u], stream ? (cudaStream_t)stream : sl.stream));
    g_fm_launches.fetch_add(sl.g_launches, std::memory_order_relaxed);
    return FMIDX_OK;
}

int fmidx_query_graph_launch(fmidx_t *h, int slot) {
    if (!h) { fm_set_error("null handle"); return FMIDX_EINVAL; }
    if (slot < 0 || slot >= FMIDX_SLOTS || !h->slots[slot].gexec[0]) { fm_set_error("slot %d holds no captured step", slot); return FMIDX_ESTATE; }
    fmidx::Slot &sl = h->slots[slot];
    FM_CUDA(cudaGraphLaunch(sl.gexec[sl.g_step++ & 1u], sl.stream));
    g_fm_launches.fetch_add(sl.g_launches, std::memory_order_relaxed);
    return FMIDX_OK;
}

int fmidx_query_submit(fmidx_t *h, int slot, const int32_t *toks, const int64_t *offs, size_t B, int64_t shift,
                       uint32_t vocab, uint64_t *sp, uint64_t *ep, uint32_t *n_distinct) {
    if (slot < 0 || slot >= FMIDX_SLOTS) { fm_set_error("slot must be 0..%d", FMIDX_SLOTS - 1); return FMIDX_EINVAL; }
    if (h && h->slots[slot].gexec[0]) { fm_set_error("slot %d is driven by a captured step (fmidx_query_graph_launch)", slot); return FMIDX_ESTATE; }
    return query_host_impl(h, toks, offs, B, shift, vocab, sp, ep, nullptr, n_distinct, slot);
}

int fmidx_query_wait(fmidx_t *h, int slot) {
    FM_TRY(enter(h, false));
    if (slot < 0 || slot >= FMIDX_SLOTS) { fm_set_error("slot must be 0..%d", FMIDX_SLOTS - 1); return FMIDX_EINVAL; }
    if (h->slots[slot].stream) FM_CUDA(cudaStreamSynchronize(h->slots[slot].stream));
    return FMIDX_OK;
}

int fmidx_query_host(fmidx_t *h, const int32_t *toks, const int64_t *offs, size_t B, int64_t shift, uint32_t vocab,
                     uint64_t *sp, uint64_t *ep, uint32_t *mask) {
    return query_host_impl(h, toks, offs, B, shift, vocab, sp, ep, mask, nullptr);
}

int fmidx_query_host_summary(fmidx_t *h, const int32_t *toks, const int64_t *offs, size_t B, int64_t shift,
                             uint32_t vocab, uint64_t *sp, uint64_t *ep, uint32_t *n_distinct) {
    if (!n_distinct) { fm_set_error("n_distinct is NULL"); return FMIDX_EINVAL; }
    return query_host_impl(h, toks, offs, B, shift, vocab, sp, ep, nullptr, n_distinct);
}

int fmidx_search_host(fmidx_t *h, const int32_t *toks, const int64_t *offs, size_t B, int64_t shift, uint64_t *sp,
                      uint64_t *ep) {
    if (h && B == 1 && h->tune_scalar_fast && toks && offs && sp && ep && offs[0] == 0 && offs[1] >= 0 && offs[1] <= fm_scalar_max_tokens()) {
        FM_TRY(enter(h, true));
        int64_t t[128];
        for (int64_t i = 0; i < offs[1]; ++i) t[i] = toks[i];
        uint64_t r[2];
        FM_TRY(fm_scalar_search(h, t, (int)offs[1], shift, 0, r));
        sp[0] = r[0];
        ep[0] = r[1];
        return FMIDX_OK;
    }
    return fmidx_query_host(h, toks, offs, B, shift, 0, sp, ep, nullptr);
}

int fmidx_locate_host(fmidx_t *h, const uint64_t *rows, size_t n, uint64_t *pos) {
    FM_TRY(enter(h, true));
    if (n == 0) return FMIDX_OK;
    cudaStream_t s = h->stream;
    FM_TRY(fm_scratch(h, h->d_scratch[6], n * 16, false));
    uint64_t *d_rows = (uint64_t *)h->d_scratch[6].p, *d_pos = d_rows + n;
    FM_CUDA(cudaMemcpyAsync(d_rows, rows, n * 8, cudaMemcpyHostToDevice, s));
    FM_TRY(fm_launch_locate(h, d_rows, n, d_pos, s));
    FM_CUDA(cudaMemcpyAsync(pos, d_pos, n * 8, cudaMemcpyDeviceToHost, s));
    FM_CUDA(cudaStreamSynchronize(s));
    return FMIDX_OK;
}

}  // extern "C"
