"""HTTP service (next-row f3): same JSON contracts as the reference's fm_index_api_service.py, with request
coalescing.  CPU part drives the app with the oracle as the engine; the GPU part with the real index."""
import random
import threading

import pytest
from starlette.testclient import TestClient

from fm_index_constrained_decoding_b200.service import create_app
from oracle.oracle import OracleFMIndex


def corpus():
    rng = random.Random(13)
    docs = [[rng.randrange(500) for _ in range(rng.randint(10, 80))] for _ in range(30)]
    docs[9] = docs[2][3:30] + docs[9]
    return rng, docs


def check_endpoints(client, o, rng, docs):
    assert client.get("/health").json() == {"status": "healthy", "n_docs": o.n_docs, "n_tokens": len(o)}
    assert client.get("/occurring_distinct").json() == o.occurring_distinct
    st = client.get("/stats").json()
    assert st["n_docs"] == o.n_docs and st["n_tokens"] == len(o) and st["has_labels"] is False
    for _ in range(20):
        d = rng.choice(docs)
        i = rng.randrange(len(d))
        q = d[i:i + rng.randint(1, 5)]
        r = client.post("/get_range", json={"sequence": q}).json()
        assert tuple(r["range"]) == o.get_range(q) and r["took_ms"] >= 0
        assert client.post("/get_count", json={"sub_sequence": q}).json()["count"] == o.get_count(q)
        assert client.post("/continuations", json={"sequence": q}).json()["continuations"] == o.get_continuations(q)
        s = client.post("/search", json={"sequence": q, "limit": 5}).json()
        assert s["count"] == o.get_count(q) and s["doc_indices"] == list(o.get_doc_indices(q))[:5]
    lows, highs = [0, 3, 7], [len(o), 9, 7]
    got = client.post("/get_distinct_count_multi", json={"lows": lows, "highs": highs}).json()["distinct_list"]
    assert [tuple(map(list, x)) for x in got] == [tuple(map(list, x)) for x in o.get_distinct_count_multi(lows, highs)]
    assert client.get("/docs/3").json() == {"doc_index": 3, "tokens": o.get_doc(3), "length": len(docs[3])}
    assert client.get("/docs/999").status_code == 404


def test_service_contracts_and_coalescing_cpu():
    rng, docs = corpus()
    o = OracleFMIndex()
    o.initialize(docs)
    app = create_app(o, window_ms=20.0)
    with TestClient(app) as client:
        check_endpoints(client, o, rng, docs)
        # concurrent requests are merged into few engine batches
        before = client.get("/stats").json()
        qs = [rng.choice(docs)[:rng.randint(1, 4)] for _ in range(24)]
        out = [None] * len(qs)

        def hit(k):
            out[k] = client.post("/get_range", json={"sequence": qs[k]}).json()["range"]
        th = [threading.Thread(target=hit, args=(k,)) for k in range(len(qs))]
        [t.start() for t in th]
        [t.join() for t in th]
        assert [tuple(x) for x in out] == [o.get_range(q) for q in qs]
        after = client.get("/stats").json()
        assert after["requests_coalesced"] - before["requests_coalesced"] == len(qs)
        assert after["device_batches"] - before["device_batches"] < len(qs)


@pytest.mark.gpu
def test_service_on_gpu_index(cuda):
    from fm_index_constrained_decoding_b200 import FMIndex
    rng, docs = corpus()
    o = OracleFMIndex()
    o.initialize(docs)
    ix = FMIndex()
    ix.initialize(docs, in_memory=True)
    assert ix.get_range_many([d[:3] for d in docs[:10]] + [[]]) == [o.get_range(d[:3]) for d in docs[:10]] + [o.get_range([])]
    with TestClient(create_app(ix, window_ms=2.0)) as client:
        check_endpoints(client, o, rng, docs)


def test_remote_index_drives_the_soft_boost_processor():
    """causal_qwen_api.py as a composition: the soft-boost processor over RemoteFMIndex over the HTTP service gives the
    scores of the restated reference over the local index."""
    import torch
    from fm_index_constrained_decoding_b200.client import RemoteFMIndex
    from fm_index_constrained_decoding_b200.processors import SoftBoostLogitsProcessor
    from oracle.oracle import soft_boost_reference
    rng, docs = corpus()
    o = OracleFMIndex()
    o.initialize(docs)
    with TestClient(create_app(o)) as client:
        remote = RemoteFMIndex(session=client)
        assert len(remote) == len(o) and remote.occurring_distinct == o.occurring_distinct
        q = docs[4][2:5]
        assert remote.get_range(q) == o.get_range(q) and remote.get_count(q) == o.get_count(q)
        assert remote.get_continuations(q) == o.get_continuations(q)
        V, EOS, PAD = 600, 590, 591
        marker = [151645, 198, 151644, 77091, 198]
        think = [151667, 271, 151668, 271]
        rows = []
        for k in range(6):
            d = rng.choice(docs)
            i = rng.randrange(len(d) - 4)
            c = d[i:i + 4]
            if k == 1:
                c = c[:-1] + [EOS]
            if k == 2:
                c = c[:-1] + [555]  # not in the corpus
            rows.append([1, 2] + marker + (think[:2] + c + [7, 8] if k == 3 else think + c))
        scores = torch.randn(len(rows), V, generator=torch.Generator().manual_seed(4))
        proc = SoftBoostLogitsProcessor(remote, 2, variant="qwen", min_new_tokens=2, pad_token_id=PAD, eos_token_id=EOS,
                                        stop_at_count=1, always_allow_eos=True)
        exp = soft_boost_reference(o, rows, scores, 2, "qwen", EOS, PAD, marker, 1, True, None, None, 2.0, 2)
        assert torch.equal(proc(torch.tensor(rows), scores), exp)
