"""Index facade over the HTTP service (`service.create_app`, or the reference's fm_index_api_service.py -- same JSON).

The reference's `document_constrained_generation_causal_qwen_api.py` is its soft-boost processor with the index calls
replaced by HTTP requests (:35, :87-97).  Here that is a composition: `RemoteFMIndex(url)` answers the calls the
processors make (`get_range_batch`, `distinct_mask_batch`, `occurring_distinct`) and the scalar calls of `index.py`
(`get_range`, `get_count`, `get_distinct_count_multi`, `get_continuations`) from the service, so

    SoftBoostLogitsProcessor(RemoteFMIndex("http://host:8000"), num_beams, variant="qwen", ...)

is the drop-in for that file.  Tensors come back on the CPU; the processors move nothing else.
"""
import numpy as np
import torch


class RemoteFMIndex:
    def __init__(self, index_url=None, session=None):
        """index_url: base URL of the service (default http://0.0.0.0:8000 like causal_qwen_api.py:11).  session: anything
        with requests' get/post -- a `requests.Session`, or an in-process client (starlette TestClient) that takes bare
        paths, in which case index_url stays None."""
        if session is None:
            import requests
            session = requests.Session()
            index_url = index_url or "http://0.0.0.0:8000"
        self.index_url = (index_url or "").rstrip("/")
        self._s = session
        self.occurring_distinct = self._get("/occurring_distinct")
        self._n_tokens = self._get("/health")["n_tokens"]

    def _get(self, path):
        r = self._s.get(self.index_url + path)
        r.raise_for_status()
        return r.json()

    def _post(self, path, payload):
        r = self._s.post(self.index_url + path, json=payload)
        r.raise_for_status()
        return r.json()

    def __len__(self):
        return self._n_tokens

    # ---- index.py surface (what causal_qwen_api.py:87-97 calls) ------------------------------------------------------
    def get_range(self, sequence):
        return tuple(self._post("/get_range", {"sequence": list(sequence)})["range"])

    def get_count(self, sequence):
        return self._post("/get_count", {"sub_sequence": list(sequence)})["count"]

    def get_distinct_count_multi(self, lows, highs):
        out = self._post("/get_distinct_count_multi", {"lows": list(lows), "highs": list(highs)})["distinct_list"]
        return [(list(d), list(c)) for d, c in out]

    def get_continuations(self, sequence):
        return self._post("/continuations", {"sequence": list(sequence)})["continuations"]

    # ---- the batched calls the processors make ------------------------------------------------------------------------
    def get_range_batch(self, toks, offs):
        t, f = toks.tolist(), offs.tolist()
        r = [self.get_range(t[f[i]:f[i + 1]]) for i in range(len(f) - 1)]
        dev = toks.device  # results live where the inputs live (the processors combine them with tensors on scores.device)
        return (torch.tensor([x[0] for x in r], dtype=torch.int64, device=dev),
                torch.tensor([x[1] for x in r], dtype=torch.int64, device=dev))

    def distinct_mask_batch(self, sp, ep, vocab):
        words = (vocab + 31) // 32
        m = np.zeros((sp.numel(), words), np.uint32)
        res = self.get_distinct_count_multi(sp.tolist(), ep.tolist()) if sp.numel() else []
        for i, (dist, _) in enumerate(res):
            t = np.asarray([x for x in dist if 0 <= x < vocab], dtype=np.int64)
            np.bitwise_or.at(m[i], t >> 5, (np.uint32(1) << (t & 31).astype(np.uint32)))
        return torch.from_numpy(m.view(np.int32)).to(sp.device)
