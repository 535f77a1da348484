"""HTTP service over the index with the JSON contracts of the reference's `fm_index_api_service.py`
(:101-255): /health, /search, /occurring_distinct, /get_count, /get_range, /get_distinct_count_multi,
/continuations, /docs/{i}, /stats -- same request/response fields (`took_ms` included).

The reference serves one request at a time (single event loop, GIL held through SWIG).  Here concurrent
range/count/continuation/search requests are COALESCED: whatever arrives within `window_ms` (or `max_batch`
requests) becomes ONE batched device call (`get_range_many` -> fmidx_search_host, then one
`get_distinct_count_multi`), which is how the GPU path is meant to be fed.

    from fm_index_constrained_decoding_b200 import FMIndex
    from fm_index_constrained_decoding_b200.service import create_app
    app = create_app(FMIndex.load(os.environ["FM_INDEX_PATH"]))     # uvicorn module:app

`create_app` accepts any engine with the index.py method surface.
"""
import asyncio
import concurrent.futures
import time
from typing import List, Optional, Tuple

from fastapi import FastAPI, HTTPException
from pydantic import BaseModel


class SearchRequest(BaseModel):
    sequence: List[int]
    limit: Optional[int] = 100


class SearchResponse(BaseModel):
    count: int
    doc_indices: List[int]
    took_ms: float


class ContinuationRequest(BaseModel):
    sequence: List[int]


class ContinuationResponse(BaseModel):
    continuations: List[int]
    took_ms: float


class CountRequest(BaseModel):
    sub_sequence: List[int]


class CountResponse(BaseModel):
    count: int
    took_ms: float


class RangeRequest(BaseModel):
    sequence: List[int]


class RangeResponse(BaseModel):
    range: Tuple[int, int]
    took_ms: float


class DistinctRequest(BaseModel):
    lows: List[int]
    highs: List[int]


class DistinctResponse(BaseModel):
    distinct_list: List[Tuple[List[int], List[int]]]
    took_ms: float


class RangeCoalescer:
    """Collects concurrent prefix lookups and resolves them with one batched engine call."""

    def __init__(self, engine, max_batch=4096, window_ms=1.0, executor=None):
        self.engine = engine
        # include/fmidx.h: a handle is used by ONE host thread at a time.  Every engine call of the service -- coalesced
        # batches, locate, extract, distinct lists -- runs on this single worker thread, so two batches can never be
        # inside the library at once (ctypes releases the GIL) and share its stream and scratch buffers.
        self.executor = executor or concurrent.futures.ThreadPoolExecutor(max_workers=1, thread_name_prefix="fmidx")
        self.max_batch = max_batch
        self.window = window_ms / 1000.0
        self.pending = []  # (sequence, want_continuations, future)
        self.flusher = None
        self.n_requests = 0
        self.n_batches = 0

    async def lookup(self, sequence, want_continuations=False):
        loop = asyncio.get_running_loop()
        fut = loop.create_future()
        self.pending.append((list(sequence), want_continuations, fut))
        self.n_requests += 1
        if len(self.pending) >= self.max_batch:
            self._flush_now(loop)
        elif self.flusher is None:
            self.flusher = loop.call_later(self.window, self._flush_now, loop)
        return await fut

    def _flush_now(self, loop):
        if self.flusher is not None:
            self.flusher.cancel()
            self.flusher = None
        batch, self.pending = self.pending, []
        if batch:
            self.n_batches += 1
            loop.create_task(self._run(loop, batch))

    async def _run(self, loop, batch):
        try:
            results = await loop.run_in_executor(self.executor, self._resolve, batch)
            for (_, _, fut), res in zip(batch, results):
                if not fut.done():
                    fut.set_result(res)
        except Exception as e:  # noqa: BLE001 -- every waiter must hear about it
            for _, _, fut in batch:
                if not fut.done():
                    fut.set_exception(e)

    def _resolve(self, batch):
        eng = self.engine
        seqs = [s for s, _, _ in batch]
        if hasattr(eng, "get_range_many"):
            ranges = eng.get_range_many(seqs)
        else:
            ranges = [eng.get_range(s) for s in seqs]
        conts = {}
        idx = [i for i, (_, want, _) in enumerate(batch) if want]
        if idx:
            multi = eng.get_distinct_count_multi([ranges[i][0] for i in idx], [ranges[i][1] for i in idx])
            for i, (tokens, _) in zip(idx, multi):
                conts[i] = tokens
        return [(tuple(ranges[i]), conts.get(i)) for i in range(len(batch))]


def create_app(index, max_batch=4096, window_ms=1.0):
    app = FastAPI(title="FM-Index Search Service (B200)", version="1.0.0")
    co = RangeCoalescer(index, max_batch=max_batch, window_ms=window_ms)
    app.state.coalescer = co

    def ms(t0):
        return (time.time() - t0) * 1000

    @app.get("/health")
    async def health_check():
        return {"status": "healthy", "n_docs": index.n_docs, "n_tokens": len(index)}

    @app.post("/search", response_model=SearchResponse)
    async def search(request: SearchRequest):
        t0 = time.time()
        try:
            (lo, hi), _ = await co.lookup(request.sequence)
            rows = range(lo, hi if not request.limit else min(hi, lo + request.limit))

            def locate():
                if hasattr(index, "doc_indices_of_rows"):
                    return index.doc_indices_of_rows(lo, rows.stop)
                return [index.get_doc_index_from_row(r) for r in rows]
            docs = await asyncio.get_running_loop().run_in_executor(co.executor, locate)
            return SearchResponse(count=hi - lo, doc_indices=docs, took_ms=ms(t0))
        except Exception as e:  # noqa: BLE001
            raise HTTPException(status_code=500, detail=str(e))

    @app.get("/occurring_distinct")
    async def get_occurring_distinct():
        return index.occurring_distinct

    @app.post("/get_count")
    async def get_count(request: CountRequest):
        t0 = time.time()
        try:
            (lo, hi), _ = await co.lookup(request.sub_sequence)
            return CountResponse(count=hi - lo, took_ms=ms(t0))
        except Exception as e:  # noqa: BLE001
            raise HTTPException(status_code=500, detail=str(e))

    @app.post("/get_range")
    async def get_range(request: RangeRequest):
        t0 = time.time()
        try:
            rng, _ = await co.lookup(request.sequence)
            return RangeResponse(range=rng, took_ms=ms(t0))
        except Exception as e:  # noqa: BLE001
            raise HTTPException(status_code=500, detail=str(e))

    @app.post("/get_distinct_count_multi")
    async def get_distinct_count_multi(request: DistinctRequest):
        t0 = time.time()
        try:
            loop = asyncio.get_running_loop()
            out = await loop.run_in_executor(co.executor, index.get_distinct_count_multi, request.lows, request.highs)
            return DistinctResponse(distinct_list=out, took_ms=ms(t0))
        except Exception as e:  # noqa: BLE001
            raise HTTPException(status_code=500, detail=str(e))

    @app.post("/continuations", response_model=ContinuationResponse)
    async def get_continuations(request: ContinuationRequest):
        t0 = time.time()
        try:
            _, conts = await co.lookup(request.sequence, want_continuations=True)
            return ContinuationResponse(continuations=conts, took_ms=ms(t0))
        except Exception as e:  # noqa: BLE001
            raise HTTPException(status_code=500, detail=str(e))

    @app.get("/docs/{doc_index}")
    async def get_document(doc_index: int):
        if doc_index < 0 or doc_index >= index.n_docs:
            raise HTTPException(status_code=404, detail="Document not found")
        try:
            doc = await asyncio.get_running_loop().run_in_executor(co.executor, index.get_doc, doc_index)
            return {"doc_index": doc_index, "tokens": doc, "length": len(doc)}
        except Exception as e:  # noqa: BLE001
            raise HTTPException(status_code=500, detail=str(e))

    @app.get("/stats")
    async def get_stats():
        return {"n_docs": index.n_docs, "n_tokens": len(index), "occurring_tokens": len(index.occurring),
                "has_labels": index.labels is not None, "requests_coalesced": co.n_requests, "device_batches": co.n_batches}

    return app
