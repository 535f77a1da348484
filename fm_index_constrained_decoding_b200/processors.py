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
    def __init__(self, index: FMIndex, num_beams: int, pad_token_id: int = 0, eos_token_id: int = 2,
                 force_decoding_from=None, stop_at_count: int = 0, always_allow_eos: bool = False,
                 forced_bos_token_id=None, skip_first_token: bool = False):
        self.index = index
        self._num_beams = num_beams
        self.pad_token_id = pad_token_id
        self.eos_token_id = eos_token_id
        self.force_decoding_from = force_decoding_from
        self.stop_at_count = stop_at_count
        self.always_allow_eos = always_allow_eos
        self.forced_bos_token_id = forced_bos_token_id
        self.skip_first_token = skip_first_token
        self.BOOST = 0.0

    def clone(self):
        return IndexBasedLogitsProcessor(
            index=self.index, num_beams=self._num_beams, pad_token_id=self.pad_token_id, eos_token_id=self.eos_token_id,
            force_decoding_from=self.force_decoding_from, stop_at_count=self.stop_at_count,
            always_allow_eos=self.always_allow_eos, forced_bos_token_id=self.forced_bos_token_id,
            skip_first_token=self.skip_first_token)

    # ---- call styles ------------------------------------------------------------------------------
    def __call__(self, input_ids, scores):
        if torch.is_tensor(input_ids) and input_ids.dim() == 2:
            hist = input_ids[:, 1:] if self.skip_first_token else input_ids
            return self.process_batch(hist.tolist(), scores)
        single = scores.dim() == 1
        hist = list(input_ids)
        out = self.process_batch([hist], scores.unsqueeze(0) if single else scores)
        return out.squeeze(0) if single else out

    # ---- the step ----------------------------------------------------------------------------------
    def process_batch(self, histories, scores):
        """histories: B token lists (all of the same length in the reference's use); scores [B, V]."""
        ix = self.index
        assert scores.is_cuda, "scores must live on the GPU that holds the index"
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
