"""The part of logits_processor_zoo.transformers.base.BaseLogitsProcessor that the reference's processors rely on
(document_constrained_generation_seq.py:2,6,18,35): a constructor and `__call__` delegating to `_process`."""


class BaseLogitsProcessor:
    def __init__(self):
        self.prompt_token_ids = None

    def _reset(self):
        pass

    def __call__(self, input_ids, scores):
        return self._process(input_ids, scores)
