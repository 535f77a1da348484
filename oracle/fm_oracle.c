/*
 * fm_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * A plain-C restatement of the reference's FM-index query path:
 *   /root/reference/cpp_modules_linux/fm_index.cpp:27-209   (class FMIndex)
 *   /root/reference/index.py:16-207                        (SHIFT / reversal / get_range ...)
 * and of the third-party algorithms those call.  The arithmetic lives in sdsl-lite
 * (simongog/sdsl-lite, cloned UNPINNED at install time, reference README.md:23; last tag
 * v2.1.1), which is NOT vendored under /root/reference and not installed in this image, so
 * the sdsl parts below restate its *published* algorithms by name:
 *   wt_int.hpp::{rank, operator[], interval_symbols/_interval_symbols}   (level-wise pointerless
 *       binary wavelet tree, nodes of a level ordered by prefix, 3 rank probes per level)
 *   rank_support_v.hpp        (512-bit superblocks: u64 absolute + 7x9-bit relative counts)
 *   suffix_array_algorithm.hpp::backward_search
 *   csa_wt.hpp::{operator[], lf},  csa_sampling_strategy.hpp::{sa_order_sa_sampling(32),
 *   isa_sampling(64)},  csa_alphabet_strategy.hpp::int_alphabet,  construct.hpp (appends the
 *   0 sentinel, rejects 0 symbols).
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or built index for this path
 * and cannot be compiled here.  Every query result is however a pure
 * function of the text (the suffix array of a text with a unique smallest sentinel is
 * unique), so this oracle is pinned instead against (1) the hand-checkable worked example
 * of SURVEY.md section 9 and (2) an independent brute-force model (oracle/brute.py: sorted
 * suffixes + list counting) on random corpora -- see tests/test_oracle.py, tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (fm_index_constrained_decoding_b200/) never does.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>

typedef uint64_t u64;
typedef int64_t i64;
typedef uint32_t u32;
typedef int32_t i32;

#define SA_DENS 32  /* csa_wt_int<> default t_dens      [sdsl csa_wt.hpp] */
#define ISA_DENS 64 /* csa_wt_int<> default t_inv_dens  [sdsl csa_wt.hpp] */

typedef struct {
    u64 n;          /* csa.size(): text length INCLUDING the appended 0 sentinel */
    u32 max_level;  /* wt_int::m_max_level = bits::hi(max symbol)+1 */
    u64 sigma;      /* wt_int::sigma = number of distinct symbols (incl. sentinel) */
    u64 tree_bits;  /* max_level * n */
    u64 *tree;      /* wt_int::m_tree, the concatenated level bit-vectors */
    u64 *rank_bb;   /* rank_support_v::m_basic_block: 2 x u64 per 512 bits */
    /* int_alphabet: char2comp via sorted symbol list, C over comp */
    u64 max_sym;
    i32 *char2comp; /* [max_sym+1]; -1 when absent */
    u64 *C;         /* [sigma+1] */
    u32 *sa_sample; /* SA[i], i % SA_DENS == 0   (sa_order_sa_sampling) */
    u32 *isa_sample;/* ISA[j], j % ISA_DENS == 0 (isa_sampling, text order) */
} fmo_t;

/* ------------------------------------------------------------------------------------------
 * Suffix sorting: SA-IS (Nong, Zhang, Chan 2009) over an integer alphabet.  sdsl uses
 * qsufsort for integer alphabets; the result (the suffix array) is unique, so any correct
 * sorter gives the same index.
 * ---------------------------------------------------------------------------------------- */
#define TGET(i) ((t[(i) >> 3] >> ((i) & 7)) & 1)
#define TSET(i, b) (t[(i) >> 3] = (unsigned char)((b) ? (t[(i) >> 3] | (1u << ((i) & 7))) : (t[(i) >> 3] & ~(1u << ((i) & 7)))))
#define IS_LMS(i) ((i) > 0 && TGET(i) && !TGET((i) - 1))

static void sais_buckets(const i32 *s, i64 n, i64 K, i32 *bkt, int end) {
    memset(bkt, 0, (size_t)K * sizeof(i32));
    for (i64 i = 0; i < n; i++) bkt[s[i]]++;
    i32 sum = 0;
    for (i64 c = 0; c < K; c++) {
        sum += bkt[c];
        bkt[c] = end ? sum : sum - bkt[c];
    }
}

static void sais_induce(const unsigned char *t, i32 *SA, const i32 *s, i32 *bkt, i64 n, i64 K) {
    sais_buckets(s, n, K, bkt, 0);
    for (i64 i = 0; i < n; i++) {
        i64 j = (i64)SA[i] - 1;
        if (SA[i] > 0 && !TGET(j)) SA[bkt[s[j]]++] = (i32)j;
    }
    sais_buckets(s, n, K, bkt, 1);
    for (i64 i = n - 1; i >= 0; i--) {
        i64 j = (i64)SA[i] - 1;
        if (SA[i] > 0 && TGET(j)) SA[--bkt[s[j]]] = (i32)j;
    }
}

/* s[n-1] must be 0 and the unique smallest symbol; symbols in [0,K). */
static void sais(const i32 *s, i32 *SA, i64 n, i64 K) {
    if (n == 1) { SA[0] = 0; return; }
    unsigned char *t = (unsigned char *)calloc((size_t)(n / 8 + 1), 1);
    i32 *bkt = (i32 *)malloc((size_t)K * sizeof(i32));
    TSET(n - 1, 1);
    TSET(n - 2, 0);
    for (i64 i = n - 3; i >= 0; i--)
        TSET(i, (s[i] < s[i + 1] || (s[i] == s[i + 1] && TGET(i + 1))) ? 1 : 0);
    /* stage 1: sort LMS substrings */
    sais_buckets(s, n, K, bkt, 1);
    for (i64 i = 0; i < n; i++) SA[i] = -1;
    for (i64 i = 1; i < n; i++)
        if (IS_LMS(i)) SA[--bkt[s[i]]] = (i32)i;
    sais_induce(t, SA, s, bkt, n, K);
    i64 n1 = 0;
    for (i64 i = 0; i < n; i++)
        if (IS_LMS(SA[i])) SA[n1++] = SA[i];
    for (i64 i = n1; i < n; i++) SA[i] = -1;
    i64 name = 0, prev = -1;
    for (i64 i = 0; i < n1; i++) {
        i64 pos = SA[i];
        int diff = 0;
        for (i64 d = 0; d < n; d++) {
            if (prev == -1 || s[pos + d] != s[prev + d] || TGET(pos + d) != TGET(prev + d)) { diff = 1; break; }
            else if (d > 0 && (IS_LMS(pos + d) || IS_LMS(prev + d))) break;
        }
        if (diff) { name++; prev = pos; }
        SA[n1 + pos / 2] = (i32)(name - 1);
    }
    for (i64 i = n - 1, j = n - 1; i >= n1; i--)
        if (SA[i] >= 0) SA[j--] = SA[i];
    i32 *SA1 = SA, *s1 = SA + n - n1;
    if (name < n1) sais(s1, SA1, n1, name);
    else for (i64 i = 0; i < n1; i++) SA1[s1[i]] = (i32)i;
    /* stage 3: induce the full SA from the sorted LMS suffixes */
    sais_buckets(s, n, K, bkt, 1);
    for (i64 i = 1, j = 0; i < n; i++)
        if (IS_LMS(i)) s1[j++] = (i32)i;
    for (i64 i = 0; i < n1; i++) SA1[i] = s1[SA1[i]];
    for (i64 i = n1; i < n; i++) SA[i] = -1;
    for (i64 i = n1 - 1; i >= 0; i--) {
        i64 j = SA[i];
        SA[i] = -1;
        SA[--bkt[s[j]]] = (i32)j;
    }
    sais_induce(t, SA, s, bkt, n, K);
    free(bkt);
    free(t);
}

/* ------------------------------------------------------------------------------------------
 * rank_support_v  [sdsl rank_support_v.hpp]: per 512-bit superblock two u64: absolute number
 * of ones before the superblock, then seven 9-bit counts relative to it for words 1..7.
 * ---------------------------------------------------------------------------------------- */
static void rank_v_build(fmo_t *f) {
    u64 nbits = f->tree_bits;
    u64 nsb = (nbits >> 9) + 1;
    f->rank_bb = (u64 *)calloc((size_t)(2 * nsb), sizeof(u64));
    u64 nwords = (nbits + 63) >> 6;
    u64 abs = 0;
    for (u64 sb = 0; sb < nsb; sb++) {
        f->rank_bb[2 * sb] = abs;
        u64 rel = 0, second = 0;
        for (u64 w = 0; w < 8; w++) {
            u64 wi = sb * 8 + w;
            if (w > 0) second |= rel << (63 - 9 * w);
            if (wi < nwords) rel += (u64)__builtin_popcountll(f->tree[wi]);
        }
        f->rank_bb[2 * sb + 1] = second;
        abs += rel;
    }
}

static inline u64 rank_v(const fmo_t *f, u64 idx) {
    const u64 *p = f->rank_bb + ((idx >> 8) & 0xFFFFFFFFFFFFFFFEULL);
    u64 r = p[0] + ((p[1] >> (63 - 9 * ((idx & 0x1FF) >> 6))) & 0x1FF);
    if (idx & 0x3F) r += (u64)__builtin_popcountll(f->tree[idx >> 6] & ((1ULL << (idx & 0x3F)) - 1));
    return r;
}

/* wt_int::rank(i, c)  [sdsl wt_int.hpp]: occurrences of c in the prefix [0..i-1]. */
static u64 wt_rank(const fmo_t *f, u64 i, u64 c) {
    if (f->max_level < 64 && (1ULL << f->max_level) <= c) return 0;
    u64 offset = 0, mask = 1ULL << (f->max_level - 1), node_size = f->n;
    for (u32 k = 0; k < f->max_level && i; ++k) {
        u64 ones_before_o = rank_v(f, offset);
        u64 ones_before_i = rank_v(f, offset + i) - ones_before_o;
        u64 ones_before_end = rank_v(f, offset + node_size) - ones_before_o;
        if (c & mask) {
            offset += node_size - ones_before_end;
            node_size = ones_before_end;
            i = ones_before_i;
        } else {
            node_size = node_size - ones_before_end;
            i = i - ones_before_i;
        }
        offset += f->n;
        mask >>= 1;
    }
    return i;
}

/* wt_int::inverse_select(i)  [sdsl wt_int.hpp]: (rank of wt[i] in [0..i-1], wt[i]). */
static u64 wt_inverse_select(const fmo_t *f, u64 i, u64 *sym) {
    u64 offset = 0, res = 0, node_size = f->n;
    for (u32 k = 0; k < f->max_level; ++k) {
        u64 ones_before_o = rank_v(f, offset);
        u64 ones_before_i = rank_v(f, offset + i) - ones_before_o;
        u64 ones_before_end = rank_v(f, offset + node_size) - ones_before_o;
        res <<= 1;
        if ((f->tree[(offset + i) >> 6] >> ((offset + i) & 63)) & 1) {
            offset += node_size - ones_before_end;
            node_size = ones_before_end;
            i = ones_before_i;
            res |= 1;
        } else {
            node_size = node_size - ones_before_end;
            i = i - ones_before_i;
        }
        offset += f->n;
    }
    *sym = res;
    return i;
}

/* wt_int::_interval_symbols  [sdsl wt_int.hpp]: DFS, 0-child first, 4 rank probes per node. */
static void wt_interval_symbols_rec(const fmo_t *f, u64 i, u64 j, u64 *k, u64 *cs, u64 *rci, u64 *rcj,
                                    u32 level, u64 path, u64 node_size, u64 offset) {
    if (level >= f->max_level) {
        rci[*k] = i;
        rcj[*k] = j;
        cs[(*k)++] = path;
        return;
    }
    u64 ones_before_o = rank_v(f, offset);
    u64 ones_before_i = rank_v(f, offset + i) - ones_before_o;
    u64 ones_before_j = rank_v(f, offset + j) - ones_before_o;
    u64 ones_before_end = rank_v(f, offset + node_size) - ones_before_o;
    if ((j - i) - (ones_before_j - ones_before_i) > 0)
        wt_interval_symbols_rec(f, i - ones_before_i, j - ones_before_j, k, cs, rci, rcj, level + 1, path << 1,
                                node_size - ones_before_end, offset + f->n);
    if ((ones_before_j - ones_before_i) > 0)
        wt_interval_symbols_rec(f, ones_before_i, ones_before_j, k, cs, rci, rcj, level + 1, (path << 1) | 1,
                                ones_before_end, offset + (node_size - ones_before_end) + f->n);
}

static void wt_interval_symbols(const fmo_t *f, u64 i, u64 j, u64 *k, u64 *cs, u64 *rci, u64 *rcj) {
    if (i == j) { *k = 0; return; }
    if (i + 1 == j) {
        u64 sym;
        u64 r = wt_inverse_select(f, i, &sym);
        cs[0] = sym; rci[0] = r; rcj[0] = r + 1; *k = 1;
        return;
    }
    *k = 0;
    wt_interval_symbols_rec(f, i, j, k, cs, rci, rcj, 0, 0, f->n, 0);
}

/* int_alphabet::char2comp  [sdsl csa_alphabet_strategy.hpp]: 0 for symbols not in the text. */
static inline u64 char2comp(const fmo_t *f, u64 c) {
    if (c > f->max_sym || f->char2comp[c] < 0) return 0;
    return (u64)f->char2comp[c];
}

/* sdsl::backward_search(csa, l, r, c, l_res, r_res)  [suffix_array_algorithm.hpp], inclusive. */
static void sdsl_backward_search(const fmo_t *f, u64 l, u64 r, u64 c, u64 *l_res, u64 *r_res) {
    u64 cc = char2comp(f, c);
    if (cc == 0 && c > 0) { *l_res = 1; *r_res = 0; return; }
    u64 c_begin = f->C[cc];
    if (l == 0 && r + 1 == f->n) { *l_res = c_begin; *r_res = f->C[cc + 1] - 1; }
    else {
        *l_res = c_begin + wt_rank(f, l, c);
        *r_res = c_begin + wt_rank(f, r + 1, c) - 1;
    }
}

/* csa_wt::lf[i] = C[comp(bwt[i])] + rank_{bwt[i]}(i). */
static inline u64 csa_lf(const fmo_t *f, u64 i) {
    u64 sym;
    u64 r = wt_inverse_select(f, i, &sym);
    return f->C[char2comp(f, sym)] + r;
}

/* ------------------------------------------------------------------------------------------ */
static void fmo_build_from_text_sa(fmo_t *f, const i32 *s, const i32 *SA, u64 n) {
    f->n = n;
    u64 max_sym = 0;
    for (u64 i = 0; i < n; i++) if ((u64)s[i] > max_sym) max_sym = (u64)s[i];
    f->max_sym = max_sym;
    u32 hi = 0;
    while ((max_sym >> (hi + 1)) != 0) hi++;
    f->max_level = hi + 1; /* bits::hi(max)+1 ; hi(0)=0 -> 1 level */
    /* alphabet + C */
    u64 *cnt = (u64 *)calloc((size_t)(max_sym + 2), sizeof(u64));
    for (u64 i = 0; i < n; i++) cnt[s[i]]++;
    f->char2comp = (i32 *)malloc((size_t)(max_sym + 1) * sizeof(i32));
    u64 sigma = 0;
    for (u64 c = 0; c <= max_sym; c++) f->char2comp[c] = cnt[c] ? (i32)sigma++ : -1;
    f->sigma = sigma;
    f->C = (u64 *)calloc((size_t)(sigma + 1), sizeof(u64));
    u64 acc = 0, comp = 0;
    for (u64 c = 0; c <= max_sym; c++) if (cnt[c]) { f->C[comp++] = acc; acc += cnt[c]; }
    f->C[sigma] = acc;
    free(cnt);
    /* BWT */
    u32 *cur = (u32 *)malloc((size_t)n * sizeof(u32));
    u32 *nxt = (u32 *)malloc((size_t)n * sizeof(u32));
    for (u64 i = 0; i < n; i++) cur[i] = (u32)(SA[i] ? s[SA[i] - 1] : s[n - 1]);
    /* samples */
    f->sa_sample = (u32 *)malloc((size_t)((n + SA_DENS - 1) / SA_DENS) * sizeof(u32));
    f->isa_sample = (u32 *)malloc((size_t)((n + ISA_DENS - 1) / ISA_DENS) * sizeof(u32));
    for (u64 i = 0; i < n; i++) {
        if (i % SA_DENS == 0) f->sa_sample[i / SA_DENS] = (u32)SA[i];
        if ((u64)SA[i] % ISA_DENS == 0) f->isa_sample[(u64)SA[i] / ISA_DENS] = (u32)i;
    }
    /* wt_int: level k bits at [k*n,(k+1)*n); nodes ordered by prefix, stable inside nodes. */
    f->tree_bits = (u64)f->max_level * n;
    f->tree = (u64 *)calloc((size_t)((f->tree_bits + 63) / 64 + 1), sizeof(u64));
    for (u32 k = 0; k < f->max_level; k++) {
        u32 shift = f->max_level - 1 - k;
        u64 base = (u64)k * n;
        /* nodes of level k are the maximal runs of equal top-k-bit prefix; inside a node the
         * 0-children keep their order and precede the 1-children (stable partition). */
        for (u64 s0 = 0; s0 < n;) {
            u64 pre = (k == 0) ? 0 : (cur[s0] >> (shift + 1));
            u64 e0 = s0, zeros = 0;
            while (e0 < n && ((k == 0) ? 0 : (cur[e0] >> (shift + 1))) == pre) {
                zeros += !((cur[e0] >> shift) & 1);
                e0++;
            }
            u64 zi = s0, oi = s0 + zeros;
            for (u64 p = s0; p < e0; p++) {
                if ((cur[p] >> shift) & 1) {
                    f->tree[(base + p) >> 6] |= 1ULL << ((base + p) & 63);
                    nxt[oi++] = cur[p];
                } else nxt[zi++] = cur[p];
            }
            s0 = e0;
        }
        u32 *tmp = cur; cur = nxt; nxt = tmp;
    }
    free(cur); free(nxt);
    rank_v_build(f);
}

/* construct(): text = data || 0.  Returns NULL when data contains a 0 symbol (sdsl throws
 * std::logic_error "contains zero symbol") or a symbol that does not fit i32. */
void *fmo_build(const u64 *data, u64 n_data) {
    u64 n = n_data + 1;
    if (n >= 0x7FFFFFFFULL) return NULL;
    i32 *s = (i32 *)malloc((size_t)n * sizeof(i32));
    u64 max_sym = 0;
    for (u64 i = 0; i < n_data; i++) {
        if (data[i] == 0 || data[i] > 0x3FFFFFFFULL) { free(s); return NULL; }
        s[i] = (i32)data[i];
        if (data[i] > max_sym) max_sym = data[i];
    }
    s[n - 1] = 0;
    i32 *SA = (i32 *)malloc((size_t)n * sizeof(i32));
    sais(s, SA, (i64)n, (i64)max_sym + 1);
    fmo_t *f = (fmo_t *)calloc(1, sizeof(fmo_t));
    fmo_build_from_text_sa(f, s, SA, n);
    free(SA); free(s);
    return f;
}

/* bench.py cpu_baseline only: build the oracle's structures from a suffix array that was
 * computed elsewhere (the text INCLUDES the trailing 0 sentinel here). */
void *fmo_build_from_sa(const u32 *text_with_sentinel, const u32 *sa, u64 n) {
    if (n >= 0x7FFFFFFFULL) return NULL;
    fmo_t *f = (fmo_t *)calloc(1, sizeof(fmo_t));
    fmo_build_from_text_sa(f, (const i32 *)text_with_sentinel, (const i32 *)sa, n);
    return f;
}

/* Independent O(n) proof that `sa` is THE suffix array of `text` (n symbols, unique smallest sentinel 0 at n-1), after
 * Burkhardt & Karkkainen's checker: (1) sa is a permutation of 0..n-1; (2) adjacent suffixes are in order, i.e.
 * T[sa[i]] < T[sa[i+1]], or the first symbols are equal and ISA[sa[i]+1] < ISA[sa[i+1]+1].  Used by bench.py before it
 * hands a suffix array computed elsewhere (the GPU builder) to fmo_build_from_sa, so that the big-corpus parity check
 * does not rest on the product.  Returns 0 when sa is correct, a positive code otherwise. */
typedef struct { const u32 *t; const u32 *sa; const u32 *isa; u64 lo, hi, n; int bad; } verify_job_t;
static void *verify_worker(void *arg) {
    verify_job_t *j = (verify_job_t *)arg;
    const u32 *t = j->t, *sa = j->sa, *isa = j->isa;
    for (u64 i = j->lo; i < j->hi; i++) {
        u32 a = sa[i], b = sa[i + 1];
        if (t[a] > t[b]) { j->bad = 3; return NULL; }
        if (t[a] == t[b]) {
            /* equal first symbols: neither is the (unique) sentinel, so a+1 and b+1 exist */
            if ((u64)a + 1 >= j->n || (u64)b + 1 >= j->n || isa[a + 1] >= isa[b + 1]) { j->bad = 4; return NULL; }
        }
    }
    return NULL;
}
int fmo_verify_sa(const u32 *text, const u32 *sa, u64 n, int nthreads) {
    if (n == 0 || text[n - 1] != 0 || sa[0] != (u32)(n - 1)) return 1;
    u32 *isa = (u32 *)malloc((size_t)n * sizeof(u32));
    memset(isa, 0xFF, (size_t)n * sizeof(u32));
    for (u64 i = 0; i < n; i++) {
        if (sa[i] >= n || isa[sa[i]] != 0xFFFFFFFFu) { free(isa); return 2; }  /* not a permutation */
        isa[sa[i]] = (u32)i;
    }
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    pthread_t th[64];
    verify_job_t jobs[64];
    const u64 pairs = n - 1;
    for (int k = 0; k < nthreads; k++) {
        verify_job_t jb = {text, sa, isa, pairs * (u64)k / (u64)nthreads, pairs * (u64)(k + 1) / (u64)nthreads, n, 0};
        jobs[k] = jb;
        pthread_create(&th[k], NULL, verify_worker, &jobs[k]);
    }
    int bad = 0;
    for (int k = 0; k < nthreads; k++) {
        pthread_join(th[k], NULL);
        if (jobs[k].bad && !bad) bad = jobs[k].bad;
    }
    free(isa);
    return bad;
}

/* Suffix array of data||0 (for cross-checking the device builder). out has n_data+1 slots. */
int fmo_suffix_array(const u64 *data, u64 n_data, u32 *out) {
    u64 n = n_data + 1;
    if (n >= 0x7FFFFFFFULL) return -1;
    i32 *s = (i32 *)malloc((size_t)n * sizeof(i32));
    u64 max_sym = 0;
    for (u64 i = 0; i < n_data; i++) {
        if (data[i] == 0 || data[i] > 0x3FFFFFFFULL) { free(s); return -1; }
        s[i] = (i32)data[i];
        if (data[i] > max_sym) max_sym = data[i];
    }
    s[n - 1] = 0;
    sais(s, (i32 *)out, (i64)n, (i64)max_sym + 1);
    free(s);
    return 0;
}

void fmo_free(void *h) {
    fmo_t *f = (fmo_t *)h;
    if (!f) return;
    free(f->tree); free(f->rank_bb); free(f->char2comp); free(f->C); free(f->sa_sample); free(f->isa_sample);
    free(f);
}

/* FMIndex::size()  fm_index.cpp:50-52 */
u64 fmo_size(void *h) { return ((fmo_t *)h)->n; }
u64 fmo_sigma(void *h) { return ((fmo_t *)h)->sigma; }

/* FMIndex::backward_search_step  fm_index.cpp:67-97: inclusive [low,high] in, half-open out,
 * (0,0) for an invalid or empty range. */
void fmo_backward_search_step(void *h, u64 symbol, u64 low, u64 high, u64 out[2]) {
    const fmo_t *f = (const fmo_t *)h;
    u64 nl = 0, nh = 0;
    if (high >= f->n) high = f->n - 1;
    if (low > high) { out[0] = 0; out[1] = 0; return; }
    sdsl_backward_search(f, low, high, symbol, &nl, &nh);
    if (nl > nh) { out[0] = 0; out[1] = 0; return; }
    out[0] = nl;
    out[1] = nh + 1;
}

/* FMIndex::backward_search_multi  fm_index.cpp:55-65: no guards; returns (l, r+1) raw.  With
 * sdsl's asserts compiled out an empty intermediate range keeps flowing through rank();
 * this restatement does the same arithmetic (u64 wrap-around included). */
void fmo_backward_search_multi(void *h, const u64 *query, u64 qlen, u64 out[2]) {
    const fmo_t *f = (const fmo_t *)h;
    u64 l = 0, r = f->n - 1;
    for (u64 i = 0; i < qlen; i++) sdsl_backward_search(f, l, r, query[i], &l, &r);
    out[0] = l;
    out[1] = r + 1;
}

/* FMIndex::distinct_count  fm_index.cpp:119-144 (allocates 3 x sigma u64 per call, as the
 * reference does).  Writes [s0,c0,s1,c1,...] to out (capacity 2*sigma) and returns the number
 * of u64 written.  with_counts=0 gives FMIndex::distinct (fm_index.cpp:99-117). */
u64 fmo_distinct_count(void *h, u64 low, u64 high, u64 *out, int with_counts) {
    const fmo_t *f = (const fmo_t *)h;
    u64 *cs = (u64 *)calloc((size_t)f->sigma, sizeof(u64));
    u64 *rci = (u64 *)calloc((size_t)f->sigma, sizeof(u64));
    u64 *rcj = (u64 *)calloc((size_t)f->sigma, sizeof(u64));
    u64 w = 0;
    if (low != high) {
        if (high > f->n) high = f->n;
        if (low < high) {
            u64 k = 0;
            wt_interval_symbols(f, low, high, &k, cs, rci, rcj);
            for (u64 i = 0; i < k; i++) {
                out[w++] = cs[i];
                if (with_counts) out[w++] = rcj[i] - rci[i];
            }
        }
    }
    free(cs); free(rci); free(rcj);
    return w;
}

/* FMIndex::locate  fm_index.cpp:168-172 ; csa_wt::operator[] with sa_order_sa_sampling. */
u64 fmo_locate(void *h, u64 row) {
    const fmo_t *f = (const fmo_t *)h;
    if (row >= f->n) return (u64)-1;
    u64 off = 0;
    while (row % SA_DENS != 0) { row = csa_lf(f, row); ++off; }
    u64 result = f->sa_sample[row / SA_DENS];
    return (result + off < f->n) ? result + off : result + off - f->n;
}

/* csa.isa[i] from text-order samples: walk LF down from the next sampled position. */
static u64 csa_isa(const fmo_t *f, u64 i) {
    u64 q = ((i + ISA_DENS - 1) / ISA_DENS) * ISA_DENS;
    u64 row;
    if (q >= f->n) { q = f->n - 1; row = 0; /* ISA[n-1] = 0: the sentinel suffix is smallest */ }
    else row = f->isa_sample[q / ISA_DENS];
    for (; q > i; --q) row = csa_lf(f, row);
    return row;
}

/* FMIndex::extract_text  fm_index.cpp:174-194, as written (isa[end-1]). Returns count. */
u64 fmo_extract_text(void *h, u64 begin, u64 end, u64 *out) {
    const fmo_t *f = (const fmo_t *)h;
    if (end <= begin) return 0;
    if (end > f->n) end = f->n;
    if (end <= begin) return 0; /* cannot happen in the reference either: unsigned end-begin-1 loop guard */
    u64 w = 0, start = csa_isa(f, end - 1), symbol, sr[2];
    wt_inverse_select(f, start, &symbol);
    out[w++] = symbol;
    if (end - begin == 1) return w;
    for (u64 i = 0; i < end - begin - 1; i++) {
        fmo_backward_search_step(h, symbol, start, start + 1, sr);
        start = sr[0];
        if (start >= f->n) break;
        wt_inverse_select(f, start, &symbol);
        out[w++] = symbol;
    }
    return w;
}

/* index.FMIndex.get_range  index.py:103-114 */
void fmo_get_range(void *h, const i64 *tokens, u64 len, i64 shift, u64 out[2]) {
    u64 start = 0, end = fmo_size(h), r[2];
    for (u64 i = 0; i < len; i++) {
        fmo_backward_search_step(h, (u64)(tokens[i] + shift), start, end - 1, r);
        start = r[0]; end = r[1];
        if (start == 0 && end == 0) break;
    }
    out[0] = start; out[1] = end;
}

/* ---- batch drivers for the CPU baseline (std::async-per-range in the reference becomes a
 * fixed pthread pool over the batch; bench.py states the thread count) -------------------- */
typedef struct {
    void *h; const i64 *toks; const i64 *offs; u64 b0, b1; i64 shift; u64 *sp, *ep;
    u32 *mask; u64 mask_words; u64 vocab; u64 *n_distinct;
} batch_job_t;

static void *batch_worker(void *arg) {
    batch_job_t *j = (batch_job_t *)arg;
    const fmo_t *f = (const fmo_t *)j->h;
    u64 *buf = j->mask ? (u64 *)malloc((size_t)(2 * f->sigma + 2) * sizeof(u64)) : NULL;
    for (u64 b = j->b0; b < j->b1; b++) {
        u64 r[2];
        fmo_get_range(j->h, j->toks + j->offs[b], (u64)(j->offs[b + 1] - j->offs[b]), j->shift, r);
        j->sp[b] = r[0]; j->ep[b] = r[1];
        if (j->mask) {
            /* index.FMIndex.get_distinct_count index.py:147-159 -> allowed-token bitmap */
            u32 *m = j->mask + b * j->mask_words;
            memset(m, 0, (size_t)j->mask_words * sizeof(u32));
            u64 w = fmo_distinct_count(j->h, r[0], r[1], buf, 1), nd = 0;
            for (u64 i = 0; i < w; i += 2) {
                if (buf[i] > 0) {
                    i64 tok = (i64)buf[i] - j->shift;
                    if (tok >= 0 && (u64)tok < j->vocab) { m[tok >> 5] |= 1u << (tok & 31); nd++; }
                }
            }
            if (j->n_distinct) j->n_distinct[b] = nd;
        }
    }
    free(buf);
    return NULL;
}

/* get_range (+ optional distinct-token bitmap) for B ragged prefixes on nthreads threads. */
void fmo_query_batch(void *h, const i64 *toks, const i64 *offs, u64 B, i64 shift, u64 *sp, u64 *ep,
                     u32 *mask, u64 mask_words, u64 vocab, u64 *n_distinct, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if ((u64)nthreads > B) nthreads = (int)(B ? B : 1);
    pthread_t *th = (pthread_t *)malloc((size_t)nthreads * sizeof(pthread_t));
    batch_job_t *jobs = (batch_job_t *)malloc((size_t)nthreads * sizeof(batch_job_t));
    for (int t = 0; t < nthreads; t++) {
        batch_job_t jb = {h, toks, offs, B * (u64)t / (u64)nthreads, B * (u64)(t + 1) / (u64)nthreads, shift, sp, ep,
                          mask, mask_words, vocab, n_distinct};
        jobs[t] = jb;
        pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

typedef struct { void *h; const u64 *rows; u64 *pos; u64 b0, b1; } loc_job_t;
static void *loc_worker(void *arg) {
    loc_job_t *j = (loc_job_t *)arg;
    for (u64 b = j->b0; b < j->b1; b++) j->pos[b] = fmo_locate(j->h, j->rows[b]);
    return NULL;
}
void fmo_locate_batch(void *h, const u64 *rows, u64 B, u64 *pos, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if ((u64)nthreads > B) nthreads = (int)(B ? B : 1);
    pthread_t *th = (pthread_t *)malloc((size_t)nthreads * sizeof(pthread_t));
    loc_job_t *jobs = (loc_job_t *)malloc((size_t)nthreads * sizeof(loc_job_t));
    for (int t = 0; t < nthreads; t++) {
        loc_job_t jb = {h, rows, pos, B * (u64)t / (u64)nthreads, B * (u64)(t + 1) / (u64)nthreads};
        jobs[t] = jb;
        pthread_create(&th[t], NULL, loc_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

/* ------------------------------------------------------------------------------------------
 * store_to_file(index, path)  (FMIndex::save, fm_index.cpp:196-199) in sdsl-lite's serialisation
 * layout AS RECALLED from sdsl-lite v2.1.1 (csa_wt::serialize -> wt_int::serialize ->
 * int_vector / rank_support_v / select_support_mcl serialize; int_alphabet::serialize).
 * UNVERIFIED against a real sdsl-lite build (none available): it exists so that the product's
 * sdsl reader (fm_index_constrained_decoding_b200/sdsl_format.py) has files to parse in tests.
 * ---------------------------------------------------------------------------------------- */
static void w_u64(FILE *f, u64 v) { fwrite(&v, 8, 1, f); }
static u32 bits_hi_plus1(u64 x) { u32 b = 0; while (x >> 1) { x >>= 1; b++; } return b + 1; }

static void w_int_vector0(FILE *f, const u64 *vals, u64 count, unsigned width) {
    u64 nbits = count * width;
    w_u64(f, nbits);
    unsigned char w8 = (unsigned char)width;
    fwrite(&w8, 1, 1, f);
    u64 nwords = (nbits + 63) >> 6;
    u64 *buf = (u64 *)calloc((size_t)(nwords + 1), 8);
    for (u64 i = 0; i < count; i++) {
        u64 bit = i * width, v = width < 64 ? (vals[i] & ((1ULL << width) - 1)) : vals[i];
        buf[bit >> 6] |= v << (bit & 63);
        if ((bit & 63) + width > 64) buf[(bit >> 6) + 1] |= v >> (64 - (bit & 63));
    }
    fwrite(buf, 8, (size_t)nwords, f);
    free(buf);
}
static void w_bit_vector(FILE *f, const u64 *words, u64 nbits) {
    w_u64(f, nbits);
    fwrite(words, 8, (size_t)((nbits + 63) >> 6), f);
}
static void w_rank_v(FILE *f, const u64 *bv, u64 nbits) { /* rank_support_v::m_basic_block : int_vector<64> */
    /* sized from the bit vector's capacity (whole 64-bit words) [sdsl, recalled]: ((capacity >> 9) + 1) << 1 words */
    u64 nwords = (nbits + 63) >> 6, nsb = ((nwords << 6) >> 9) + 1, abs = 0;
    w_u64(f, nsb * 2 * 64);
    for (u64 sb = 0; sb < nsb; sb++) {
        u64 rel = 0, second = 0;
        for (u64 w = 0; w < 8; w++) {
            if (w > 0) second |= rel << (63 - 9 * w);
            if (sb * 8 + w < nwords) rel += (u64)__builtin_popcountll(bv[sb * 8 + w]);
        }
        w_u64(f, abs);
        w_u64(f, second);
        abs += rel;
    }
}
static void w_select_mcl(FILE *f, const u64 *bv, u64 nbits, int bit) {
    u64 cnt = 0;
    for (u64 i = 0; i < nbits; i++) cnt += (((bv[i >> 6] >> (i & 63)) & 1) == (u64)bit);
    w_u64(f, cnt);
    if (!cnt) return;
    u64 sb = (cnt + 4095) >> 12;
    u32 logn = bits_hi_plus1(nbits);
    u64 logn4 = (u64)logn * logn * logn * logn;
    u64 *pos = (u64 *)malloc((size_t)cnt * 8), k = 0;
    for (u64 i = 0; i < nbits; i++)
        if (((bv[i >> 6] >> (i & 63)) & 1) == (u64)bit) pos[k++] = i;
    u64 *sup = (u64 *)malloc((size_t)sb * 8), *mol = (u64 *)calloc((size_t)((sb + 63) / 64 + 1), 8);
    for (u64 s = 0; s < sb; s++) {
        sup[s] = pos[s * 4096];
        u64 last = (s + 1) * 4096 - 1 < cnt ? (s + 1) * 4096 - 1 : cnt - 1;
        if (!(pos[last] - pos[s * 4096] > logn4)) mol[s >> 6] |= 1ULL << (s & 63); /* 1 = miniblock */
    }
    w_int_vector0(f, sup, sb, logn);
    w_bit_vector(f, mol, sb);
    for (u64 s = 0; s < sb; s++) {
        u64 first = s * 4096, last = first + 4095 < cnt ? first + 4095 : cnt - 1;
        if ((mol[s >> 6] >> (s & 63)) & 1) {
            u64 mini[64];
            memset(mini, 0, sizeof(mini));
            for (u64 j = 0; j < 64 && first + j * 64 <= last; j++) mini[j] = pos[first + j * 64] - pos[first];
            w_int_vector0(f, mini, 64, bits_hi_plus1(pos[last] - pos[first]));
        } else {
            u64 *lng = (u64 *)calloc(4096, 8);
            for (u64 j = 0; first + j <= last; j++) lng[j] = pos[first + j];
            w_int_vector0(f, lng, 4096, bits_hi_plus1(pos[last]));
            free(lng);
        }
    }
    free(pos); free(sup); free(mol);
}

int fmo_store_sdsl(void *h, const char *path) {
    const fmo_t *x = (const fmo_t *)h;
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    /* wt_int */
    w_u64(f, x->n);
    w_u64(f, x->sigma);
    w_bit_vector(f, x->tree, x->tree_bits);
    w_rank_v(f, x->tree, x->tree_bits);
    w_select_mcl(f, x->tree, x->tree_bits, 1);
    w_select_mcl(f, x->tree, x->tree_bits, 0);
    fwrite(&x->max_level, 4, 1, f);
    /* sa / isa samples */
    unsigned wn = bits_hi_plus1(x->n);
    u64 nsa = (x->n + SA_DENS - 1) / SA_DENS, nisa = (x->n + ISA_DENS - 1) / ISA_DENS;
    u64 *tmp = (u64 *)malloc((size_t)(nsa > nisa ? nsa : nisa) * 8 + 8);
    for (u64 i = 0; i < nsa; i++) tmp[i] = x->sa_sample[i];
    w_int_vector0(f, tmp, nsa, wn);
    for (u64 i = 0; i < nisa; i++) tmp[i] = x->isa_sample[i];
    w_int_vector0(f, tmp, nisa, wn);
    free(tmp);
    /* int_alphabet: m_char, m_char_rank, m_char_select, m_C, m_sigma */
    u64 nchar = x->max_sym + 1;
    u64 *cb = (u64 *)calloc((size_t)((nchar + 63) / 64 + 1), 8);
    for (u64 c = 0; c < nchar; c++)
        if (x->char2comp[c] >= 0) cb[c >> 6] |= 1ULL << (c & 63);
    w_bit_vector(f, cb, nchar);
    w_rank_v(f, cb, nchar);
    w_select_mcl(f, cb, nchar, 1);
    free(cb);
    w_int_vector0(f, x->C, x->sigma + 1, wn);
    w_u64(f, x->sigma);
    fclose(f);
    return 0;
}
