"""Hard-mask logits processor over the fused GPU path.

Same constructor arguments, call conventions and masking semantics as the reference's
`IndexBasedLogitsProcessor`:
  * vLLM per-request style   `proc(input_ids: Tuple[int, ...], scores[V]) -> scores[V]`
    (document_constrained_generation_vllm.py:46-137)
  * HF `generate()` batch style `proc(input_ids: LongTensor[B*beams, L], scores[B*beams, V]) -> scores`
    with the decoder-start token in column 0 (document_constrained_generation_seq.py:35-123);
    pass `skip_first_token=True` for that variant.
but one step is three kernel launches for the whole batch -- K1 ranges of the beam histories
(and of the histories minus their last token when `stop_at_count` is in force), then the fused
enumerate + -inf mask -- instead of per-beam Python loops, SWIG calls and index scatters.

out[b, v] = scores[b, v] + (0.0 if v in S_b else -inf), BOOST = 0.0 as in the reference.
"""
import torch

from .index import FMIndex, SHIFT, ragged


class IndexBasedLogitsProcessor:
    """Hard-mask processor (document_constrained_generation_vllm.py:7-137 / _seq.py:6-123).

    CUDA `input_ids` tensors take the DEVICE path: no `.tolist()`, no ragged host lists, no host synchronisation.  The
    per-beam state (sp, ep) = get_range(force + history) is kept on the GPU between calls; a call whose histories are one
    token longer than the previous call's continues it with ONE backward-search step per beam (`fmidx_decode_step`: rows
    gathered by `beam_idx`, or by parents matched on the device when it is not given), where the reference re-searches
    every history twice (vllm.py:91-92).  Otherwise the parent states come from one batched search of the histories minus
    their last token, and the same step kernel finishes -- so both ways share every decision rule.  `reset()` forgets
    the state."""

    def __init__(self, index: FMIndex, num_beams: int, pad_token_id: int = 0, eos_token_id: int = 2,
                 force_decoding_from=None, stop_at_count: int = 0, always_allow_eos: bool = False,
                 forced_bos_token_id=None, skip_first_token: bool = False, incremental: bool = True):
        self.index = index
        self._num_beams = num_beams
        self.pad_token_id = pad_token_id
        self.eos_token_id = eos_token_id
        self.force_decoding_from = force_decoding_from
        self.stop_at_count = stop_at_count
        self.always_allow_eos = always_allow_eos
        self.forced_bos_token_id = forced_bos_token_id
        self.skip_first_token = skip_first_token
        self.incremental = incremental
        self.BOOST = 0.0
        self._state = None       # (sp, ep, history tensor [R, L]) of the previous device-path call
        self.steps_incremental = 0
        self.steps_searched = 0

    def clone(self):
        return IndexBasedLogitsProcessor(
            index=self.index, num_beams=self._num_beams, pad_token_id=self.pad_token_id, eos_token_id=self.eos_token_id,
            force_decoding_from=self.force_decoding_from, stop_at_count=self.stop_at_count,
            always_allow_eos=self.always_allow_eos, forced_bos_token_id=self.forced_bos_token_id,
            skip_first_token=self.skip_first_token, incremental=self.incremental)

    def reset(self):
        self._state = None

    # ---- call styles ------------------------------------------------------------------------------
    def __call__(self, input_ids, scores, beam_idx=None):
        if torch.is_tensor(input_ids) and input_ids.dim() == 2:
            hist = input_ids[:, 1:] if self.skip_first_token else input_ids
            if hist.is_cuda and scores.is_cuda and isinstance(self.index, FMIndex):
                return self.process_device(hist, scores, beam_idx)
            return self.process_batch(hist.tolist(), scores)
        single = scores.dim() == 1
        hist = list(input_ids)
        out = self.process_batch([hist], scores.unsqueeze(0) if single else scores)
        return out.squeeze(0) if single else out

    # ---- the step, device-resident -------------------------------------------------------------------------------------
    def process_device(self, hist, scores, beam_idx=None):
        """hist: int64 CUDA tensor [R, L] (equal-length histories); scores [R, V] CUDA; beam_idx: optional int tensor [R],
        row r of this call continues row beam_idx[r] of the previous call."""
        import ctypes as C
        from . import _native as N
        ix = self.index
        R, V = scores.shape
        dev = scores.device
        scores = scores.contiguous()
        eos, pad = self.eos_token_id, self.pad_token_id
        allow = eos if self.always_allow_eos else -1
        L = hist.shape[1]
        if self.forced_bos_token_id is not None:
            if L == 0:  # only the forced BOS token (vllm.py:55-58); always_allow_eos is NOT applied on this early return
                z = torch.zeros(R, dtype=torch.int64, device=dev)
                ovr = torch.full((R,), int(self.forced_bos_token_id), dtype=torch.int32, device=dev)
                self._state = None
                return ix.mask_logits(z, z, scores, out=torch.empty_like(scores), override_tok=ovr, always_allow=-1)
            hist = hist[:, 1:]
            L -= 1
        hist = hist.to(torch.int64)
        force = list(self.force_decoding_from) if self.force_decoding_from is not None else []
        if L == 0:
            # S = occurring_distinct = distinct symbols of rows [0, len(index)) (vllm.py:62-66, index.py:206)
            sp = torch.zeros(R, dtype=torch.int64, device=dev)
            ep = torch.full((R,), len(ix), dtype=torch.int64, device=dev)
            self._state = None
            return ix.mask_logits(sp, ep, scores, out=torch.empty_like(scores), always_allow=allow)
        st = self._state
        parent = None
        if self.incremental and st is not None and st[2].shape == (R, L - 1) and st[2].device == dev:
            psp, pep = st[0], st[1]
            if beam_idx is not None:
                parent = beam_idx.to(device=dev, dtype=torch.int32).contiguous()
            elif L > 1:  # match every row's history[:-1] against the previous call's rows, on the device
                eq = (hist[:, None, :-1] == st[2][None, :, :]).all(-1)
                parent = eq.to(torch.int8).argmax(1).to(torch.int32).contiguous()
            self.steps_incremental += 1
        else:
            # parent states by ONE batched search of force + history[:-1] (device-side ragged layout, no host lists)
            head = hist[:, :-1]
            if force:
                head = torch.cat([torch.tensor(force, dtype=torch.int64, device=dev).expand(R, -1), head], 1)
            Lh = head.shape[1]
            offs = torch.arange(R + 1, dtype=torch.int64, device=dev) * Lh
            psp, pep = ix.get_range_batch(head.to(torch.int32).contiguous().view(-1), offs)
            self.steps_searched += 1
        buf = torch.empty((4, R), dtype=torch.int64, device=dev)
        override = torch.empty(R, dtype=torch.int32, device=dev)
        last = hist[:, -1]
        N.check(N.lib().fmidx_decode_step(
            ix._h, C.c_void_p(last.data_ptr()), last.stride(0), C.c_void_p(parent.data_ptr()) if parent is not None else None,
            C.c_void_p(psp.data_ptr()), C.c_void_p(pep.data_ptr()), C.c_void_p(buf[0].data_ptr()), C.c_void_p(buf[1].data_ptr()), R,
            SHIFT, eos, pad, int(self.stop_at_count), C.c_void_p(buf[2].data_ptr()), C.c_void_p(buf[3].data_ptr()),
            C.c_void_p(override.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        self._state = (buf[0], buf[1], hist.clone()) if self.incremental else None
        return ix.mask_logits(buf[2], buf[3], scores, out=torch.empty_like(scores), override_tok=override, always_allow=allow)

    # ---- the step, from host token lists ---------------------------------------------------------------------------------
    def process_batch(self, histories, scores):
        """histories: B token lists (all of the same length in the reference's use); scores [B, V]."""
        ix = self.index
        assert scores.is_cuda or not isinstance(ix, FMIndex), "scores must live on the GPU that holds the index"
        B, V = scores.shape
        dev = scores.device
        scores = scores.contiguous()
        eos, pad = self.eos_token_id, self.pad_token_id
        allow = eos if self.always_allow_eos else -1
        L = len(histories[0]) if B else 0
        if self.forced_bos_token_id is not None:
            if L == 0:
                # only the forced BOS token (vllm.py:55-58); always_allow_eos is NOT applied on this early return
                z = torch.zeros(B, dtype=torch.int64, device=dev)
                ovr = torch.full((B,), int(self.forced_bos_token_id), dtype=torch.int32, device=dev)
                return ix.mask_logits(z, z, scores, out=torch.empty_like(scores), override_tok=ovr, always_allow=-1)
            histories = [h[1:] for h in histories]
            L -= 1
        if L == 0:
            # S = occurring_distinct = distinct symbols of rows [0, len(index)) (vllm.py:62-66, index.py:206)
            sp = torch.zeros(B, dtype=torch.int64, device=dev)
            ep = torch.full((B,), len(ix), dtype=torch.int64, device=dev)
            return ix.mask_logits(sp, ep, scores, out=torch.empty_like(scores), always_allow=allow)
        force = list(self.force_decoding_from) if self.force_decoding_from is not None else []
        last = torch.tensor([h[-1] for h in histories], device=dev)
        finished = (last == eos) | (last == pad)
        toks, offs = ragged([force + list(h) for h in histories])
        sp, ep = ix.get_range_batch(torch.from_numpy(toks).to(dev), torch.from_numpy(offs).to(dev))
        zero = torch.zeros_like(sp)
        sp = torch.where(finished, zero, sp)  # (low, high, count) = (0, 0, 0) for finished rows, vllm.py:81-84
        ep = torch.where(finished, zero, ep)
        override = torch.full((B,), -1, dtype=torch.int32, device=dev)
        override[finished] = pad  # vllm.py:110-118
        if self.stop_at_count > 0:
            ptoks, poffs = ragged([force + list(h[:-1]) for h in histories])
            psp, pep = ix.get_range_batch(torch.from_numpy(ptoks).to(dev), torch.from_numpy(poffs).to(dev))
            count = torch.where(finished, zero, pep - psp)
            override[count <= self.stop_at_count] = eos  # checked first in the reference, vllm.py:105-108
        return ix.mask_logits(sp, ep, scores, out=torch.empty_like(scores), override_tok=override, always_allow=allow)


class SoftBoostLogitsProcessor:
    """The reference's soft-boost ("causal") processors over the batched GPU path:

      variant "causal": document_constrained_generation_causal.py:7-180
      variant "qwen"  : document_constrained_generation_causal_qwen.py:7-214 (thinking prefix, min_new_tokens)

    out = scores + mask, mask = 0 everywhere except, per row whose trailing corpus n-gram `s` is non-empty,
    BOOST * factor**len(s) on the tokens that continue `s` in the corpus and BOOST/2 on the other corpus unigrams
    (qwen: BOOST when nothing continues `s`); rows without a trailing n-gram get BOOST on every corpus unigram; EOS is
    boosted by BOOST * factor**generated_length (qwen: only after min_new_tokens and only for rows that finished
    thinking).  Reference behaviours kept as written: the trailing-n-gram search only ever looks up the LAST token
    (causal.py:68-73 returns inside the first loop iteration), override rows ({eos} under stop_at_count, {pad} after
    eos/pad), `count` is the count of the n-gram itself.

    One step = host bookkeeping on the token lists, ONE batched search (2 queries per row: the last token and the
    whole n-gram), ONE enumeration launch group for all rows, and elementwise tensor ops in the dtype of `scores`
    (so the additions round exactly like the reference's `scores + mask`).  Every row must contain the chat-template
    end marker exactly once (the only case in which the reference's own row bookkeeping is well defined).
    """
    END_MARKERS = {"qwen": [151645, 198, 151644, 77091, 198], "llama": [78191, 128007, 271]}
    THINK_PREFIX = [151667, 271, 151668, 271]
    THINK_DONE = (151668, 271)

    def __init__(self, index, num_beams, variant="causal", model_name="llama", min_new_tokens=5, pad_token_id=0,
                 eos_token_id=2, force_decoding_from=None, stop_at_count=0, always_allow_eos=False,
                 forced_bos_token_id=None, length_reward_factor=2.0, end_marker=None):
        assert variant in ("causal", "qwen")
        self.index = index
        self._num_beams = num_beams
        self.variant = variant
        self.min_new_tokens = min_new_tokens
        self.pad_token_id = pad_token_id
        self.eos_token_id = eos_token_id
        self.force_decoding_from = force_decoding_from
        self.stop_at_count = stop_at_count
        self.always_allow_eos = always_allow_eos
        self.forced_bos_token_id = forced_bos_token_id
        self.length_reward_factor = length_reward_factor
        self.BOOST = 10.0
        if end_marker is not None:
            self.end_marker = list(end_marker)
        elif variant == "qwen" or str(model_name).lower().startswith("qwen"):
            self.end_marker = self.END_MARKERS["qwen"]
        else:
            self.end_marker = self.END_MARKERS["llama"]
        self._unigrams = {}  # (device, V) -> bool [V]

    def __call__(self, input_ids, scores):
        return self._process(input_ids, scores)

    # ---- host bookkeeping (token lists) ----------------------------------------------------------------------------
    def _after_end_marker(self, rows):
        m, marker, out = len(self.end_marker), self.end_marker, []
        for row in rows:
            hits = [i for i in range(len(row) - m + 1) if row[i:i + m] == marker]
            if len(hits) != 1:
                raise ValueError("every row of input_ids must contain the end marker exactly once")
            out.append(row[hits[0] + m:])
        return out

    def _candidate(self, row):
        """None: row is skipped (qwen: still thinking); otherwise the sequence whose last token decides between
        'no trailing n-gram' and 'the whole sequence is the n-gram'."""
        if self.variant == "causal":
            return row
        if not (len(row) >= 2 and any((row[i], row[i + 1]) == self.THINK_DONE for i in range(len(row) - 1))):
            return None
        k = len(self.THINK_PREFIX)
        if len(row) < k:
            return [] if row == self.THINK_PREFIX[:len(row)] else row
        return row[k:] if row[:k] == self.THINK_PREFIX else row

    def _unigram_mask(self, device, V):
        key = (str(device), V)
        if key not in self._unigrams:
            u = torch.zeros(V, dtype=torch.bool)
            toks = [t for t in self.index.occurring_distinct if 0 <= t < V]
            u[torch.tensor(toks, dtype=torch.long)] = True
            self._unigrams[key] = u.to(device)
        return self._unigrams[key]

    # ---- the step ----------------------------------------------------------------------------------------------------
    def _process(self, input_ids, scores):
        rows = self._after_end_marker(input_ids.tolist() if torch.is_tensor(input_ids) else [list(r) for r in input_ids])
        R, V = scores.shape
        dev, dt = scores.device, scores.dtype
        eos, pad, BOOST = self.eos_token_id, self.pad_token_id, self.BOOST
        mask = torch.zeros_like(scores)
        if self.variant == "causal" and self.forced_bos_token_id is not None:
            if len(rows[0]) == 0:
                mask[:, self.forced_bos_token_id] = BOOST
                return scores + mask
            rows = [r[1:] for r in rows]
        width = len(rows[0])
        unig = self._unigram_mask(dev, V)
        col = torch.arange(V, device=dev)
        if width == 0:
            if self.variant == "causal":
                mask = torch.where(unig.unsqueeze(0), torch.tensor(BOOST, dtype=dt, device=dev), mask)
                if self.always_allow_eos:
                    mask[:, eos] = BOOST
            return scores + mask
        cands = [self._candidate(r) for r in rows]
        force = list(self.force_decoding_from) if self.force_decoding_from is not None else []
        # two queries per row: [last token] and force + candidate (dummy empty queries where there is no candidate)
        queries = [[c[-1]] if c else [] for c in cands] + [force + c if c else [] for c in cands]
        toks, offs = ragged(queries)
        sp2, ep2 = self.index.get_range_batch(torch.from_numpy(toks).to(dev), torch.from_numpy(offs).to(dev))
        skip = torch.tensor([c is None for c in cands], device=dev)
        has = torch.tensor([bool(c) for c in cands], device=dev)
        fin_last = torch.tensor([bool(c) and c[-1] in (eos, pad) for c in cands], device=dev)
        length = torch.tensor([len(c) if c else 0 for c in cands], device=dev)
        ngram = has & ((ep2[:R] - sp2[:R]) > 0)   # the last token occurs in the corpus -> the n-gram is the whole candidate
        zero = torch.zeros_like(sp2[:R])
        live = ngram & ~fin_last
        sp, ep = torch.where(live, sp2[R:], zero), torch.where(live, ep2[R:], zero)
        words = self.index.distinct_mask_batch(sp, ep, V)
        if self.stop_at_count > 0:
            ov_eos = ngram & ((ep - sp) <= self.stop_at_count)
        else:
            ov_eos = torch.zeros_like(ngram)
        ov_pad = ngram & ~ov_eos & fin_last
        # boosts are computed as Python floats (like the reference) and only then cast to the dtype of `scores`
        f = self.length_reward_factor
        table = torch.tensor([BOOST * (f ** n) for n in range(width + len(self.THINK_PREFIX) + 1)], dtype=torch.float64).to(dt).to(dev)
        b_len = table[length]
        full, half = torch.tensor(BOOST, dtype=dt, device=dev), torch.tensor(BOOST / 2, dtype=dt, device=dev)
        eos_rule = self.always_allow_eos and (self.variant == "causal" or width >= self.min_new_tokens)
        e = torch.tensor(BOOST * (f ** width), dtype=torch.float64).to(dt).to(dev) if eos_rule else None
        if scores.is_cuda and hasattr(self.index, "_h"):
            return self._apply_on_device(scores, words, ngram, skip, ov_eos, ov_pad, b_len, full, half, e)
        sh = torch.arange(32, device=dev, dtype=words.dtype)
        D = ((words.unsqueeze(-1) >> sh) & 1).bool().view(R, -1)[:, :V]
        D = torch.where(ov_eos.unsqueeze(1), (col == eos).unsqueeze(0), D)
        D = torch.where(ov_pad.unsqueeze(1), (col == pad).unsqueeze(0), D)
        if self.variant == "qwen":
            add = torch.where(D.any(1), half, full)
        else:
            add = half.expand(R)
        z = torch.zeros((), dtype=dt, device=dev)
        m_ngram = torch.where(D, b_len.unsqueeze(1), torch.where(unig.unsqueeze(0), add.unsqueeze(1), z))
        m_uni = torch.where(unig, full, z).unsqueeze(0)
        mask = torch.where(ngram.unsqueeze(1), m_ngram, torch.where(skip.unsqueeze(1), z, m_uni))
        if e is not None:
            if self.variant == "causal":
                mask[:, eos] = e
            else:
                mask[:, eos] = torch.where(skip, mask[:, eos], e)
        return scores + mask

    def _apply_on_device(self, scores, words, ngram, skip, ov_eos, ov_pad, b_len, full, half, e):
        """The arithmetic as ONE streaming kernel over bitmaps (fmidx_boost_logits): no [R, V] intermediate tensors.  Override
        rows ({eos} under stop_at_count, {pad} after eos / pad) are patched into the [R, V/32] continuation bitmaps."""
        import ctypes as C
        from . import _native as N
        R, V = scores.shape
        dev = scores.device
        eos, pad = self.eos_token_id, self.pad_token_id
        ov = ov_eos | ov_pad
        tok = torch.where(ov_eos, torch.full_like(ov_eos, eos, dtype=torch.int64), torch.full_like(ov_eos, pad, dtype=torch.int64))
        one = torch.ones((), dtype=torch.int64, device=dev)
        bit = torch.bitwise_left_shift(one, tok & 31)
        bit = torch.where(bit >= 2 ** 31, bit - 2 ** 32, bit).to(torch.int32)  # bit 31 lives in the sign of the int32 word
        words = torch.where(ov.unsqueeze(1), torch.zeros_like(words), words)
        rows = torch.arange(R, device=dev)
        in_vocab = ov & (tok >= 0) & (tok < V)
        patched = torch.where(in_vocab, bit, words[rows, (tok >> 5).clamp(0, words.shape[1] - 1)])
        words[rows, (tok >> 5).clamp(0, words.shape[1] - 1)] = patched
        words = words.contiguous()
        nd = torch.empty(R, dtype=torch.int32, device=dev)
        sptr = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        L = N.lib()
        N.check(L.fmidx_mask_popcount(self.index._h, C.c_void_p(words.data_ptr()), R, V, C.c_void_p(nd.data_ptr()), sptr))
        kind = torch.where(ngram, 2, torch.where(skip, 0, 1)).to(torch.int32)
        vd = b_len.to(torch.float32).contiguous()
        vu_ne = half.to(torch.float32).expand(R).contiguous()
        if self.variant == "qwen":
            vu_e = full.to(torch.float32).expand(R).contiguous()
        else:
            vu_e = torch.where(ngram, half.to(torch.float32), full.to(torch.float32)).contiguous()
        ve = (e.to(torch.float32) if e is not None else torch.tensor(float("nan"), device=dev)).expand(R).contiguous()
        ub = self._unigram_words(dev, V)
        scores = scores.contiguous()
        out = torch.empty_like(scores)
        code = N.DTYPE_CODES[str(scores.dtype).replace("torch.", "")]
        N.check(L.fmidx_boost_logits(self.index._h, C.c_void_p(scores.data_ptr()), C.c_void_p(out.data_ptr()), code, scores.stride(0), R, V,
                                     C.c_void_p(words.data_ptr()), C.c_void_p(nd.data_ptr()), C.c_void_p(ub.data_ptr()),
                                     C.c_void_p(kind.data_ptr()), C.c_void_p(vd.data_ptr()), C.c_void_p(vu_ne.data_ptr()),
                                     C.c_void_p(vu_e.data_ptr()), C.c_void_p(ve.data_ptr()), eos, sptr))
        return out

    def _unigram_words(self, device, V):
        """The corpus unigrams (index.occurring_distinct) as a resident [ceil(V/32)] int32 bitmap."""
        key = ("words", str(device), V)
        if key not in self._unigrams:
            import numpy as np
            w = np.zeros((V + 31) // 32, dtype=np.uint32)
            t = np.asarray([t for t in self.index.occurring_distinct if 0 <= t < V], dtype=np.int64)
            np.bitwise_or.at(w, t >> 5, (np.uint32(1) << (t & 31).astype(np.uint32)))
            self._unigrams[key] = torch.from_numpy(w.view(np.int32)).to(device)
        return self._unigrams[key]


class CausalIndexBasedLogitsProcessor(SoftBoostLogitsProcessor):
    """Constructor of document_constrained_generation_causal.py:8-20."""

    def __init__(self, index, num_beams, model_name, pad_token_id=0, eos_token_id=2, force_decoding_from=None,
                 stop_at_count=0, always_allow_eos=False, forced_bos_token_id=None, length_reward_factor=2.0):
        super().__init__(index, num_beams, variant="causal", model_name=model_name, pad_token_id=pad_token_id,
                         eos_token_id=eos_token_id, force_decoding_from=force_decoding_from, stop_at_count=stop_at_count,
                         always_allow_eos=always_allow_eos, forced_bos_token_id=forced_bos_token_id,
                         length_reward_factor=length_reward_factor)


class QwenIndexBasedLogitsProcessor(SoftBoostLogitsProcessor):
    """Constructor of document_constrained_generation_causal_qwen.py:8-20."""

    def __init__(self, index, num_beams, min_new_tokens=5, pad_token_id=0, eos_token_id=2, force_decoding_from=None,
                 stop_at_count=0, always_allow_eos=False, forced_bos_token_id=None, length_reward_factor=2.0):
        super().__init__(index, num_beams, variant="qwen", min_new_tokens=min_new_tokens, pad_token_id=pad_token_id,
                         eos_token_id=eos_token_id, force_decoding_from=force_decoding_from, stop_at_count=stop_at_count,
                         always_allow_eos=always_allow_eos, forced_bos_token_id=forced_bos_token_id,
                         length_reward_factor=length_reward_factor)
