"""CPU suite: the vLLM v1 batch-level adapter -- batch bookkeeping (added / removed / moved requests, live output
lists, per-request opt-in) and row selection -- against the reference's per-request arithmetic
(oracle.hard_mask_reference).  The index is a stand-in answering the facade's batched calls from the CPU oracle."""
import random

import pytest
import torch

V, EOS, PAD = 5000, 4990, 4991


class OracleBackedIndex:
    def __init__(self, o):
        self.o = o

    def __len__(self):
        return len(self.o)

    def get_range_batch(self, toks, offs):
        t, f = toks.tolist(), offs.tolist()
        r = [self.o.get_range(t[f[i]:f[i + 1]]) for i in range(len(f) - 1)]
        return torch.tensor([x[0] for x in r], dtype=torch.int64), torch.tensor([x[1] for x in r], dtype=torch.int64)

    def mask_logits(self, sp, ep, logits, out=None, override_tok=None, always_allow=-1):
        out = logits if out is None else out
        for b in range(logits.shape[0]):
            m = torch.full_like(logits[b], float("-inf"))
            ov = -1 if override_tok is None else int(override_tok[b])
            allowed = [ov] if ov >= 0 else [t for t in self.o.get_distinct_count(int(sp[b]), int(ep[b]))[0] if 0 <= t < logits.shape[1]]
            m[torch.tensor(allowed, dtype=torch.long)] = 0.0
            if always_allow >= 0:
                m[always_allow] = 0.0
            out[b] = logits[b] + m
        return out


def test_batch_level_adapter_follows_the_batch():
    interface = pytest.importorskip("vllm.v1.sample.logits_processor.interface")
    from vllm import SamplingParams
    from fm_index_constrained_decoding_b200.vllm_adapter import make_index_logits_processor
    from oracle.oracle import OracleFMIndex, hard_mask_reference
    rng = random.Random(11)
    docs = [[rng.randrange(2000) for _ in range(rng.randint(15, 60))] + [EOS] for _ in range(50)]
    o = OracleFMIndex()
    o.initialize(docs)
    Proc = make_index_logits_processor(OracleBackedIndex(o), pad_token_id=PAD, eos_token_id=EOS, always_allow_eos=True)
    proc = Proc(None, torch.device("cpu"), False)
    assert proc.is_argmax_invariant() is False
    on, off = SamplingParams(extra_args={"fm_index": True}), SamplingParams()
    stop2 = SamplingParams(extra_args={"fm_index": {"stop_at_count": 2}})
    with pytest.raises(ValueError):
        Proc.validate_params(SamplingParams(extra_args={"fm_index": 3}))

    def expect(rows, logits):
        out = logits.clone()
        for i, (stop, hist) in rows.items():
            out[i] = hard_mask_reference(o, [list(hist)], logits[i:i + 1], 1, EOS, PAD, stop, True)[0]
        return out

    def same(a, b):
        return torch.equal(torch.nan_to_num(a, nan=0.0), torch.nan_to_num(b, nan=0.0))

    g = torch.Generator().manual_seed(0)
    # step 0: four requests, #1 not constrained, all with empty outputs
    outs = [[] for _ in range(4)]
    proc.update_state(interface.BatchUpdate(4, removed=[], moved=[], added=[(0, on, [1, 2], outs[0]), (1, off, [3], outs[1]),
                                                                            (2, stop2, [4], outs[2]), (3, on, [5], outs[3])]))
    logits = torch.randn(4, V, generator=g)
    assert same(proc.apply(logits.clone()), expect({0: (0, []), 2: (2, []), 3: (0, [])}, logits))
    # steps 1..3: the engine appends to the SAME lists; no batch change -> update_state(None)
    d = docs[7]
    for step in range(3):
        outs[0].append(d[3 + step])
        outs[1].append(17)
        outs[2].append(docs[9][5 + step])
        outs[3].append(EOS if step == 2 else docs[2][step])
        proc.update_state(None)
        logits = torch.randn(4, V, generator=g)
        assert same(proc.apply(logits.clone()), expect({0: (0, outs[0]), 2: (2, outs[2]), 3: (0, outs[3])}, logits))
    # request 3 finishes and 0 is moved into... a swap of 0 and 2, removal of 3, a new request at index 3
    new = [docs[1][0]]
    proc.update_state(interface.BatchUpdate(4, removed=[], added=[(3, on, [9], new)],
                                            moved=[(0, 2, interface.MoveDirectionality.SWAP)]))
    logits = torch.randn(4, V, generator=g)
    assert same(proc.apply(logits.clone()), expect({2: (0, outs[0]), 0: (2, outs[2]), 3: (0, new)}, logits))
    # shrink: request at index 3 removed, the one at index 2 moved down to index 1 (replacing the unconstrained one)
    proc.update_state(interface.BatchUpdate(2, removed=[3], added=[], moved=[(2, 1, interface.MoveDirectionality.UNIDIRECTIONAL)]))
    logits = torch.randn(2, V, generator=g)
    assert same(proc.apply(logits.clone()), expect({1: (0, outs[0]), 0: (2, outs[2])}, logits))
