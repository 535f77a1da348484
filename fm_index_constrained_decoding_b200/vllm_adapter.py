"""vLLM v1 adapter: the hard-mask constraint as a BATCH-level logits processor.

The reference hands vLLM one Python callable per request (`SamplingParams(logits_processors=[proc])`,
index_for_vllm.py:30-48 with document_constrained_generation_vllm.py:46-137 as the callable): vLLM's v0 interface,
one call per request per step, each with its own index lookups and host->device scatter.  vLLM >= 0.10 (0.22 is what
this image carries) dropped that interface for `vllm.v1.sample.logits_processor.LogitsProcessor`: one object per
engine, told about batch changes through `update_state(BatchUpdate)` and asked to rewrite the whole `[num_reqs, V]`
logits tensor in `apply` -- which is exactly the shape of the fused path: one batched search and one fused
enumerate + mask launch for every constrained request of the step.

    from fm_index_constrained_decoding_b200.vllm_adapter import make_index_logits_processor
    Proc = make_index_logits_processor(index, pad_token_id=pad, eos_token_id=eos, always_allow_eos=True)
    llm = LLM(model=..., logits_processors=[Proc])
    llm.generate(prompts, SamplingParams(max_tokens=64, extra_args={"fm_index": True}))

A request opts in with `extra_args={"fm_index": True}` (or a dict overriding `stop_at_count`); requests without the
key are left untouched.  As in the reference, the constraint sees the GENERATED tokens of the request only.
"""
import torch

from .processors import IndexBasedLogitsProcessor


def make_index_logits_processor(index, pad_token_id, eos_token_id, stop_at_count=0, always_allow_eos=False,
                                force_decoding_from=None, extra_arg="fm_index"):
    """Returns a `vllm.v1.sample.logits_processor.LogitsProcessor` subclass bound to `index`."""
    from vllm.v1.sample.logits_processor.builtin import process_dict_updates
    from vllm.v1.sample.logits_processor.interface import LogitsProcessor

    def hard(stop):
        return IndexBasedLogitsProcessor(index, num_beams=1, pad_token_id=pad_token_id, eos_token_id=eos_token_id,
                                         force_decoding_from=force_decoding_from, stop_at_count=stop,
                                         always_allow_eos=always_allow_eos)

    class FMIndexLogitsProcessor(LogitsProcessor):
        @classmethod
        def validate_params(cls, sampling_params):
            v = (sampling_params.extra_args or {}).get(extra_arg)
            if v is not None and not isinstance(v, (bool, dict)):
                raise ValueError("extra_args[%r] must be a bool or a dict" % extra_arg)

        def __init__(self, vllm_config, device, is_pin_memory):
            self.device = device
            self.rows = {}    # batch index -> (stop_at_count, live list of the request's output token ids)
            self.procs = {}   # stop_at_count -> processor

        def is_argmax_invariant(self):
            return False  # masking changes the arg max

        @staticmethod
        def _new_state(params, _prompt_tok_ids, output_tok_ids):
            v = (params.extra_args or {}).get(extra_arg)
            if not v:
                return None
            stop = v.get("stop_at_count", stop_at_count) if isinstance(v, dict) else stop_at_count
            return int(stop), output_tok_ids

        def update_state(self, batch_update):
            process_dict_updates(self.rows, batch_update, self._new_state)

        def apply(self, logits):
            if not self.rows:
                return logits
            # one fused call per (stop_at_count, empty / non-empty history) group: the empty history takes the
            # reference's `occurring_distinct` branch (vllm.py:62-66), which process_batch decides per call
            groups = {}
            for i, (stop, out) in self.rows.items():
                if i < logits.shape[0]:
                    groups.setdefault((stop, len(out) == 0), []).append(i)
            for (stop, _), idx in groups.items():
                if stop not in self.procs:
                    self.procs[stop] = hard(stop)
                sel = torch.tensor(idx, device=logits.device)
                hist = [list(self.rows[i][1]) for i in idx]
                logits[sel] = self.procs[stop].process_batch(hist, logits[sel])
            return logits

    return FMIndexLogitsProcessor
