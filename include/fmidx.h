/*
 * fmidx.h -- C ABI of libfmidx.so: the B200-native FM-index query path.
 *
 * Drop-in boundary for the SWIG + sdsl-lite binding of the reference
 * (cpp_modules_linux/fm_index.i:7-21 exporting cpp_modules_linux/fm_index.hpp:20-44).
 * Plain pointers and sizes only; no exceptions and no torch types cross this boundary.
 *
 * Conventions
 *   - every function returns 0 on success, a negative FMIDX_E* code on failure;
 *     fmidx_last_error() returns a thread-local message for the last failure.
 *   - "host" pointers are ordinary CPU memory, "d_" pointers are CUDA device pointers on the
 *     index's device; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - batched device entry points only ENQUEUE work on `stream`; they do not synchronise.
 *   - rows / positions are uint64_t at the ABI like the reference's `unsigned long`
 *     (fm_index.hpp:16-18); one index holds fewer than 2^32 symbols (shard larger corpora).
 *   - there is NO CPU fallback: without a CUDA device every call fails with FMIDX_ECUDA.
 *   - a handle is used by ONE host thread at a time (like the reference object under the GIL); work is
 *     made concurrent on the DEVICE: device entry points that take a ws_slot / slot may be in flight on different
 *     streams at once, the ones without share one scratch workspace and must be stream-ordered with each other.
 */
#ifndef FMIDX_H
#define FMIDX_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fmidx fmidx_t;

#define FMIDX_OK 0
#define FMIDX_EINVAL (-1)  /* bad argument */
#define FMIDX_ECUDA (-2)   /* CUDA runtime error / no device */
#define FMIDX_EIO (-3)     /* file error */
#define FMIDX_ESTATE (-4)  /* index not initialised */
#define FMIDX_ERANGE (-5)  /* output buffer too small / symbol too large */
#define FMIDX_EZERO (-6)   /* text contains the reserved symbol 0 (sdsl construct rejects it) */

#define FMIDX_F32 0
#define FMIDX_BF16 1
#define FMIDX_F16 2

const char *fmidx_last_error(void);
int fmidx_abi_version(void);

/* ---- lifecycle: FMIndex::FMIndex / ~FMIndex (fm_index.cpp:27-31) ------------------------- */
int fmidx_new(int device, fmidx_t **out);
void fmidx_free(fmidx_t *h);
/* build options, set before initialize/load: sa_dens 1 = full SA resident (locate is one
 * gather), >1 = text-order SA samples (reference default 32, csa_wt_int<>); isa_dens likewise
 * (reference default 64); keep_text keeps the symbol text resident so extract is a gather. */
int fmidx_set_options(fmidx_t *h, uint32_t sa_dens, uint32_t isa_dens, int keep_text);

/* run-time tunables of a handle (defaults come from the environment when the handle is created): "scan_max" -- ranges of up
 * to this many rows are read off the resident BWT instead of walking the matrix ($FMIDX_SCAN_MAX, 16384); "bytes_occ" -- CTAs
 * per SM of the byte-level enumeration kernel ($FMIDX_BYTES_OCC, 5); "unigram_table" -- 1/0: use / ignore the unigram
 * continuation table at query time ($FMIDX_UNIGRAM_TABLE, 1); "scalar_fast" -- 1/0: scalar calls as one zero-copy launch with the
 * result polled from mapped pinned memory ($FMIDX_SCALAR_FAST, 1); "scalar_speculate" -- 1/0: a scalar search also enumerates the
 * range it returns, for the distinct_count call that usually follows ($FMIDX_SCALAR_SPECULATE, 1); "lazy_clear" -- 0/1/2: zero
 * only the mask slices a buffer's previous use dirtied: 1 = buffers the library owns (pipelined / captured host forms), 2 = also
 * the caller's d_mask of fmidx_query_batch / fmidx_query_graph_prepare_device, under the caller's promise that nothing else
 * writes that buffer between the calls of its slot ($FMIDX_LAZY_CLEAR, 1); "probe_phase" -- measurement only.  Results never
 * depend on them (lazy_clear 2: as long as the promise holds). */
int fmidx_tune(fmidx_t *h, const char *key, int64_t value);

/* ---- build: FMIndex::initialize (fm_index.cpp:33-41), initialize_from_file (:43-48) ------- */
int fmidx_initialize(fmidx_t *h, const uint64_t *data, size_t n);
int fmidx_initialize_from_file(fmidx_t *h, const char *path, int width);
/* same, text already on the device as u32 symbols (synthetic corpora generated on the GPU) */
int fmidx_initialize_device(fmidx_t *h, const uint32_t *d_symbols, size_t n, void *stream);

/* ---- persistence: FMIndex::save (fm_index.cpp:196-199), load_FMIndex (:201-209) ----------- */
int fmidx_save(fmidx_t *h, const char *path);
int fmidx_load(const char *path, int device, fmidx_t **out);

/* ---- scalar queries, host in/out, synchronous --------------------------------------------- */
/* FMIndex::size (fm_index.cpp:50-52) */
int fmidx_size(fmidx_t *h, uint64_t *out);
/* FMIndex::backward_search_step (fm_index.cpp:67-97): inclusive [low,high] in, half-open out,
 * (0,0) when invalid/empty */
int fmidx_backward_search_step(fmidx_t *h, uint64_t symbol, uint64_t low, uint64_t high, uint64_t out[2]);
/* FMIndex::backward_search_multi (fm_index.cpp:55-65): returns (l, r+1) with no guards */
int fmidx_backward_search_multi(fmidx_t *h, const uint64_t *query, size_t n, uint64_t out[2]);
/* FMIndex::distinct (fm_index.cpp:99-117): ascending symbols of BWT[low,high) */
int fmidx_distinct(fmidx_t *h, uint64_t low, uint64_t high, uint64_t *out, size_t cap, size_t *k);
/* FMIndex::distinct_count (fm_index.cpp:119-144): interleaved [s0,c0,s1,c1,...]; *k = #u64 */
int fmidx_distinct_count(fmidx_t *h, uint64_t low, uint64_t high, uint64_t *out, size_t cap, size_t *k);
/* FMIndex::distinct_count_multi (fm_index.cpp:146-166): ranges in one device batch; interleaved
 * lists concatenated in `out`, list r occupying out[offsets[r] .. offsets[r+1]).  If cap is too
 * small returns FMIDX_ERANGE with offsets[nr] = required size. */
int fmidx_distinct_count_multi(fmidx_t *h, const uint64_t *lows, const uint64_t *highs, size_t nr,
                               uint64_t *out, size_t cap, uint64_t *offsets);
/* FMIndex::locate (fm_index.cpp:168-172): SA[row], or 2^64-1 when row >= size */
int fmidx_locate(fmidx_t *h, uint64_t row, uint64_t *out);
/* FMIndex::extract_text (fm_index.cpp:174-194), as written (starts from isa[end-1]) */
int fmidx_extract_text(fmidx_t *h, uint64_t begin, uint64_t end, uint64_t *out, size_t cap, size_t *k);

/* ---- batched device entry points (the hot path) -------------------------------------------- */
/* index.FMIndex.get_range (index.py:103-114) for B ragged prefixes: prefix b is
 * d_toks[d_offs[b] .. d_offs[b+1]); symbols are token + shift.  d_sp/d_ep receive [sp,ep). */
int fmidx_search_batch(fmidx_t *h, const int32_t *d_toks, const int64_t *d_offs, size_t B, int64_t shift,
                       uint64_t *d_sp, uint64_t *d_ep, void *stream);
/* B independent FMIndex::backward_search_step calls (incremental decoding: one LF step per
 * beam per generated token instead of re-searching the history). Inclusive highs in. */
int fmidx_step_batch(fmidx_t *h, const uint64_t *d_symbols, const uint64_t *d_lows, const uint64_t *d_highs,
                     size_t B, uint64_t *d_sp, uint64_t *d_ep, void *stream);
/* Incremental constrained decoding: the reference re-searches the whole history twice per beam per step
 * (document_constrained_generation_vllm.py:91-92); here the per-beam state (sp, ep) = get_range(force + history) stays on
 * the device and one step costs ONE backward_search_step per beam.  Row r continues the state of row d_parent[r] (NULL:
 * r itself) with the token d_last_tok[r * tok_stride] (the last column of input_ids, read on the device); the new state goes
 * to d_state_*_out (double-buffered).  The kernel also emits what fmidx_mask_logits needs for this step, decided row by row
 * like vllm.py:78-118: d_query_sp/ep ((0,0) for rows whose last token is eos / pad) and d_override_tok ({eos} when
 * stop_at_count > 0 and the PARENT state is at most stop_at_count wide -- that width is count(history[:-1]) -- else {pad}
 * for finished rows, else -1).  No host synchronisation, no host-side token lists. */
int fmidx_decode_step(fmidx_t *h, const int64_t *d_last_tok, int64_t tok_stride, const int32_t *d_parent, const uint64_t *d_state_sp_in,
                      const uint64_t *d_state_ep_in, uint64_t *d_state_sp_out, uint64_t *d_state_ep_out, size_t B, int64_t shift,
                      int64_t eos, int64_t pad, int64_t stop_at_count, uint64_t *d_query_sp, uint64_t *d_query_ep,
                      int32_t *d_override_tok, void *stream);
/* index.FMIndex.get_distinct_count_multi (index.py:161-174) as bitmaps: for each half-open
 * range, bit v of row b of d_mask ([B, ceil(vocab/32)] u32) is set iff symbol v+shift occurs in
 * BWT[sp,ep) and 0 <= v < vocab (the sentinel and symbols < shift are dropped like index.py's
 * `d > 0` filter).  d_counts (optional, [B, vocab] u32) receives the occurrence count at every
 * SET position (other entries are left untouched). */
int fmidx_distinct_mask_batch(fmidx_t *h, const uint64_t *d_sp, const uint64_t *d_ep, size_t B, int64_t shift,
                              uint32_t vocab, uint32_t *d_mask, uint32_t *d_counts, void *stream);
/* One decoding step's lookups for B beams in ONE call: fmidx_search_batch followed by
 * fmidx_distinct_mask_batch on the same stream (what index.py:93-131 `get_range` + index.py:146-159
 * `get_distinct_count` do per beam in document_constrained_generation_vllm.py:72-137), d_sp/d_ep and d_mask as
 * in those two calls (d_mask NULL or vocab 0: ranges only).  ws_slot (0..31) selects the scratch workspace the
 * enumeration uses: calls that may run concurrently on DIFFERENT streams must use different slots, calls that
 * share a slot must be stream-ordered.  The slots are the ones fmidx_query_submit uses: do not mix the two on one
 * slot number. */
int fmidx_query_batch(fmidx_t *h, int ws_slot, const int32_t *d_toks, const int64_t *d_offs, size_t B, int64_t shift,
                      uint32_t vocab, uint64_t *d_sp, uint64_t *d_ep, uint32_t *d_mask, void *stream);
/* same, additionally filling d_occ ([B][ceil(nslices/32)] u32, nslices = ceil(vocab/4096)): bit s of a row is set
 * when the row's 4096-token slice s may hold set bits (conservative).  The sharded step publishes these words to the peers. */
int fmidx_distinct_mask_batch_occ(fmidx_t *h, const uint64_t *d_sp, const uint64_t *d_ep, size_t B, int64_t shift,
                                  uint32_t vocab, uint32_t *d_mask, uint32_t *d_counts, uint32_t *d_occ, void *stream);
/* ---- document-sharded step; no reference counterpart: the reference's only parallelism is one
 * std::async per range (fm_index.cpp:146-166) --------------------------------------------------------------------
 * One process per GPU, each holding the index of its own document shard; the B prefixes of a step are replicated.
 * A shard context owns n_sets (<= 32) independent buffer sets in peer-visible memory (CUDA IPC); a set holds TWO
 * [B][ceil(vocab/32)] bitmap buffers and TWO inboxes (one slot per shard: slice-occupancy words + occurrence counts), used
 * alternately by consecutive steps, and one data flag per shard.  Set-up: every rank creates its context, exports the
 * 64-byte handles of its buffers (kind 0 = bitmap buffer buf, 1 = inbox buf, buf = 2 * set + parity; 2 = the flag block),
 * the host language exchanges them (torch.distributed all_gather_object), every rank imports every peer's, then
 * fmidx_shard_ready.
 * fmidx_shard_step enqueues ONE whole step of set `set` on `stream` (NULL: the set's own stream) with six launches and
 * no NCCL: search -> enumeration into bitmaps + occupancy + counts (the CTA that finishes last publishes the side data
 * into every shard's inbox and raises the data flags) -> per peer: wait for its data flag, pull its occupied 4096-token
 * slices over NVLink, OR them into the rows this rank owns; total counts.  No acknowledgements: when this rank's pull of
 * step k-1 has returned, every peer has published step k-1 and therefore finished its pull of step k-2, the last reader of
 * the buffers step k overwrites.  Afterwards (stream-ordered) rank r owns the final bitmaps of prefixes
 * [r*B/world, (r+1)*B/world) in the buffer of that step's parity and the total occurrence counts of all B prefixes
 * (fmidx_shard_result; consume them before the set's second-next step).  d_sp/d_ep receive THIS shard's ranges (ranges are
 * shard-local and never reduced).  Steps on different sets may be in flight on different streams; every rank must issue the
 * same steps on the same sets in the same order.
 * fmidx_shard_submit / fmidx_shard_wait: host-buffer form on the set's own stream (H2D of the prefixes and D2H of the
 * ranges, total counts and distinct-token counts of the owned prefixes inside the call). */
typedef struct fmidx_shard fmidx_shard_t;
int fmidx_shard_create(fmidx_t *h, uint32_t world, uint32_t rank, size_t B, uint32_t vocab, uint32_t n_sets,
                       fmidx_shard_t **out);
int fmidx_shard_export(fmidx_shard_t *x, uint32_t buf, int kind, unsigned char handle[64]);
int fmidx_shard_import(fmidx_shard_t *x, uint32_t buf, int kind, uint32_t peer, const unsigned char handle[64]);
int fmidx_shard_ready(fmidx_shard_t *x);
int fmidx_shard_step(fmidx_shard_t *x, uint32_t set, const int32_t *d_toks, const int64_t *d_offs, int64_t shift, uint64_t *d_sp,
                     uint64_t *d_ep, void *stream);
/* step number k (1, 2, ...) of a set writes bitmap buffer k & 1; parity 0 / 1 selects that buffer, -1 the one of the step issued last */
/* The step of one buffer set captured in CUDA graphs for FIXED device buffers (the caller rewrites d_toks / d_offs in place
 * between launches): the step number lives in device memory (advanced by the step's first enumeration kernel, read by the
 * publication and by the pull kernel), so a step has no per-step host argument.  prepare runs ONE real step (every shard must
 * prepare the same sets in the same order) and captures one graph per buffer parity; launch replays them alternately on the
 * set's own stream (fmidx_shard_wait synchronises it).  Eager steps (fmidx_shard_step / fmidx_shard_submit) may be mixed in:
 * every step of a set, captured or not, advances the same step number and the same scratch epoch. */
int fmidx_shard_graph_prepare(fmidx_shard_t *x, uint32_t set, const int32_t *d_toks, const int64_t *d_offs, int64_t shift, uint64_t *d_sp,
                              uint64_t *d_ep);
int fmidx_shard_graph_launch(fmidx_shard_t *x, uint32_t set);
int fmidx_shard_result(fmidx_shard_t *x, uint32_t set, int parity, uint32_t **d_own_masks, uint64_t **d_counts, void **stream);
int fmidx_shard_submit(fmidx_shard_t *x, uint32_t set, const int32_t *toks, const int64_t *offs, int64_t shift, uint64_t *sp,
                       uint64_t *ep, uint64_t *counts, uint32_t *n_distinct_own);
int fmidx_shard_wait(fmidx_shard_t *x, uint32_t set);
/* every rank: after a barrier that follows its last step (peers may still be pulling otherwise) */
void fmidx_shard_destroy(fmidx_shard_t *x);
/* peer-visible device buffers (CUDA IPC): alloc on the owner, ship the 64-byte handle, open on the peers */
int fmidx_ipc_alloc(fmidx_t *h, size_t bytes, void **d_ptr, unsigned char handle[64]);
int fmidx_ipc_open(fmidx_t *h, const unsigned char handle[64], void **d_ptr);
int fmidx_ipc_close(fmidx_t *h, void *d_ptr);
int fmidx_ipc_free(fmidx_t *h, void *d_ptr);
/* Fused constrained-decoding step (document_constrained_generation_vllm.py:52,105-137):
 *   out[b,v] = logits[b,v] + (v in S_b ? 0.0 : -inf)
 * S_b = distinct next tokens of [sp_b,ep_b); replaced by {override_tok[b]} when
 * d_override_tok != NULL and override_tok[b] >= 0 (the eos/pad/stop_at_count rows); token
 * `always_allow` (>= 0) is additionally allowed in every row (always_allow_eos).
 * d_extra_masks (optional): n_extra bitmaps [n_extra][B][ceil(vocab/32)] OR-ed into S_b (the
 * per-shard masks gathered from the other GPUs).  d_out may alias d_logits (in place).
 * row_stride in elements.  d_mask_out (optional) receives the final bitmap S_b. */
int fmidx_mask_logits(fmidx_t *h, const uint64_t *d_sp, const uint64_t *d_ep, size_t B, int64_t shift,
                      uint32_t vocab, const void *d_logits, void *d_out, int dtype, int64_t row_stride,
                      const int32_t *d_override_tok, int32_t always_allow, const uint32_t *d_extra_masks,
                      uint32_t n_extra, uint32_t *d_mask_out, void *stream);
/* Soft-boost arithmetic of the causal processors as ONE streaming kernel over bitmaps
 * (document_constrained_generation_causal.py:137-178, _causal_qwen.py:159-212):  out[b,v] = logits[b,v] + m[b,v],
 *   d_kind[b] == 0: m = 0;   == 1: m = d_vu_empty[b] on the corpus unigrams (bitmap d_unigram, [ceil(vocab/32)]);
 *   == 2: m = d_vd[b] on the row's continuation bitmap d_dmask[b] ([B][ceil(vocab/32)], from fmidx_distinct_mask_batch /
 *         fmidx_query_batch), and on the OTHER unigrams d_vu_nonempty[b], or d_vu_empty[b] when the row's bitmap is empty
 *         (d_n_distinct[b] == 0, from fmidx_mask_popcount);
 *   finally m[eos] = d_ve[b] unless d_ve[b] is NaN (kind 0 rows are never touched).
 * The values are float32 holding numbers already rounded to the dtype of the logits; the sum is formed in fp32 and rounded
 * once.  d_out may alias d_logits. */
int fmidx_boost_logits(fmidx_t *h, const void *d_logits, void *d_out, int dtype, int64_t row_stride, size_t B, uint32_t vocab,
                       const uint32_t *d_dmask, const uint32_t *d_n_distinct, const uint32_t *d_unigram, const int32_t *d_kind,
                       const float *d_vd, const float *d_vu_nonempty, const float *d_vu_empty, const float *d_ve, int32_t eos,
                       void *stream);
/* bitmap + counts -> sparse ascending (token,count) lists (index.py:161-174 return shape).
 * d_list_offsets [B+1] must hold the exclusive prefix sums of the per-row popcounts, which
 * fmidx_mask_popcount computes into d_n ([B] u32). */
/* out[i] = OR over p < n_pieces of pieces[p][i], i < n_words (n_words % 4 == 0): the local half of
 * the cross-shard mask OR-reduction (NCCL has no bitwise-OR op; the pieces arrive by all-to-all) */
int fmidx_mask_or(fmidx_t *h, const uint32_t *d_pieces, uint32_t n_pieces, size_t n_words, uint32_t *d_out, void *stream);
int fmidx_mask_popcount(fmidx_t *h, const uint32_t *d_mask, size_t B, uint32_t vocab, uint32_t *d_n, void *stream);
int fmidx_mask_compact(fmidx_t *h, const uint32_t *d_mask, const uint32_t *d_counts, size_t B, uint32_t vocab,
                       const uint64_t *d_list_offsets, int64_t *d_tokens, uint64_t *d_cnts, void *stream);
/* FMIndex::locate for n rows */
int fmidx_locate_batch(fmidx_t *h, const uint64_t *d_rows, size_t n, uint64_t *d_pos, void *stream);
/* FMIndex::extract_text for n (begin,end) pairs; row i of d_out ([n, max_len] u64) receives
 * min(len_i, max_len) symbols, d_len[i] the count */
int fmidx_extract_batch(fmidx_t *h, const uint64_t *d_begin, const uint64_t *d_end, size_t n, uint64_t max_len,
                        uint64_t *d_out, uint64_t *d_len, void *stream);

/* ---- host-buffer batch calls (what a non-CUDA caller binds; copies are inside the call) ---- */
int fmidx_search_host(fmidx_t *h, const int32_t *toks, const int64_t *offs, size_t B, int64_t shift,
                      uint64_t *sp, uint64_t *ep);
/* search + distinct-token bitmaps in one call; mask [B, ceil(vocab/32)] may be NULL */
int fmidx_query_host(fmidx_t *h, const int32_t *toks, const int64_t *offs, size_t B, int64_t shift,
                     uint32_t vocab, uint64_t *sp, uint64_t *ep, uint32_t *mask);
/* same, but the masks stay in device memory (where the logits kernel consumes them); the host
 * receives the ranges and n_distinct[b] = number of distinct next tokens of prefix b */
int fmidx_query_host_summary(fmidx_t *h, const int32_t *toks, const int64_t *offs, size_t B, int64_t shift,
                             uint32_t vocab, uint64_t *sp, uint64_t *ep, uint32_t *n_distinct);
/* pipelined form of fmidx_query_host_summary: up to 32 batches in flight, each in its own slot (own stream,
 * buffers and workspace).  submit enqueues H2D copies, kernels and D2H copies and returns; wait blocks until the
 * slot's host buffers (which should be pinned) hold the results.  n_distinct may be NULL (ranges only). */
int fmidx_query_submit(fmidx_t *h, int slot, const int32_t *toks, const int64_t *offs, size_t B, int64_t shift,
                       uint32_t vocab, uint64_t *sp, uint64_t *ep, uint32_t *n_distinct);
int fmidx_query_wait(fmidx_t *h, int slot);
/* The same step as ONE CUDA-graph launch per batch: a slot prepared here owns pinned host staging buffers -- the caller
 * writes a batch into *h_in ([offs: (B+1) int64][toks: up to max_tokens int32], offs[0] = 0, B fixed) and, after
 * fmidx_query_graph_launch + fmidx_query_wait, reads *h_out ([sp: B u64][ep: B u64][n_distinct: B u32]).  The captured
 * sequence is H2D of the input block -> search -> enumeration -> popcount -> D2H of the result block on the slot's
 * stream; ten driver calls of fmidx_query_submit become one.  A prepared slot cannot be used with fmidx_query_submit. */
int fmidx_query_graph_prepare(fmidx_t *h, int slot, size_t B, size_t max_tokens, int64_t shift, uint32_t vocab, void **h_in,
                              void **h_out);
int fmidx_query_graph_launch(fmidx_t *h, int slot);
/* The device-resident step (what fmidx_query_batch enqueues: search + row zeroing -> enumeration into d_mask) captured
 * once for FIXED device buffers -- the serving-loop pattern: the caller rewrites d_toks / d_offs in place (B fixed,
 * offs[0] = 0) and replays.  fmidx_query_graph_launch_on replays the slot's captured step on `stream` (NULL: the slot's
 * own stream, which fmidx_query_wait synchronises); replays of ONE slot must be stream-ordered, different slots are
 * independent.  Replaces four kernel launches per step by one graph launch (the step is otherwise bound by the host's
 * launch rate).  No reference counterpart (the reference issues one SWIG call per token, index.py:103-114). */
int fmidx_query_graph_prepare_device(fmidx_t *h, int slot, const int32_t *d_toks, const int64_t *d_offs, size_t B, int64_t shift,
                                     uint32_t vocab, uint64_t *d_sp, uint64_t *d_ep, uint32_t *d_mask);
int fmidx_query_graph_launch_on(fmidx_t *h, int slot, void *stream);
int fmidx_locate_host(fmidx_t *h, const uint64_t *rows, size_t n, uint64_t *pos);

/* ---- introspection -------------------------------------------------------------------------- */
/* info[0]=size n, [1]=levels, [2]=max_sym, [3]=sa_dens, [4]=isa_dens, [5]=device bytes,
 * [6]=cell bytes per level, [7]=keep_text */
int fmidx_info(fmidx_t *h, uint64_t info[8]);
/* copy the suffix array / BWT of the last build to the host (testing the builder); the SA is
 * only available when sa_dens == 1 */
int fmidx_debug_copy_sa(fmidx_t *h, uint32_t *out, size_t cap);
/* Measurement aid (SURVEY.md 8(d) "random-gather peak"): n_loads independent random loads of unit_bytes (32: one
 * lane per 32-B unit; 128: one octet per aligned 128-B line) over this index's large arrays (cells, and SA / ISA /
 * text when resident) taken as one address space; *out_gbps receives
 * n_loads * unit_bytes / time.  Synchronous, on the handle's private stream.  No reference counterpart. */
int fmidx_gather_peak(fmidx_t *h, int unit_bytes, size_t n_loads, double *out_gbps);
/* number of kernels this library has launched in this process (bench.py gpu_launches) */
uint64_t fmidx_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
