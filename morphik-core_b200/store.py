"""B200MultiVectorStore -- the drop-in ``BaseVectorStore`` for Morphik's ColPali (multi-vector) retrieval.

Keeps the plugin contract of core/vector_store/base_vector_store.py:7-65 byte for byte (four async methods, same
argument names/meaning, same return shapes) plus the de-facto extras callers use: ``initialize()``
(core/app_factory.py:78-80), ``close()``, and the ``.storage`` / ``.uri`` attributes
(core/vector_store/dual_multivector_store.py:224-232).  Semantics follow the two reference providers:

  provider "postgres" (MultiVectorStore, multi_vector_store.py)   -> ``mode="binary"``: sign-bit quantise, exhaustive
      Hamming MaxSim, score = SQL max_sim value (multiples of 1/128), ORDER BY score DESC LIMIT k.
  provider "morphik" (FastMultiVectorStore, fast_multivector_store.py) -> ``mode="bf16"``: float MaxSim; the reference
      reranks <=75 ANN candidates, this store scores EVERY authorised page exactly (a superset of that behaviour).
  ``mode="int8"`` is new (BASELINE config 3).

Everything heavy runs on the GPU through libb200ms (see index.py); this class only owns the id <-> (document, chunk)
catalogue, the doc_ids -> page-mask conversion, payload bookkeeping and DocumentChunk construction.  GPU calls are
serialised with a lock and pushed off the event loop with asyncio.to_thread (SURVEY 8b "Threading").

Concurrent ``query_similar`` coroutines (the API server's situation) are COALESCED: while one GPU pass is running, the
requests that arrive queue up and the next pass scores them together, each against its own authorised-page mask
(``b200ms_search_host_masked``).  A single 32-token query leaves 3/4 of every 128-row MMA tile empty and the scan is
HBM-bound, so up to ~8 queries ride for the price of one; a lone query is not delayed (there is no batching window).
"""
from __future__ import annotations

import asyncio
import collections
import json
import logging
import threading
import time
import weakref
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .catalog import PageCatalog, PageRecord
from .models import DocumentChunk

logger = logging.getLogger(__name__)

try:  # inside a Morphik checkout the reference ABC is the real base class
    from core.vector_store.base_vector_store import BaseVectorStore  # type: ignore
except Exception:  # standalone: same abstract surface
    from abc import ABC, abstractmethod

    class BaseVectorStore(ABC):  # mirrors core/vector_store/base_vector_store.py:7-65
        @abstractmethod
        async def store_embeddings(self, chunks, app_id=None): ...

        @abstractmethod
        async def query_similar(self, query_embedding, k, doc_ids=None, app_id=None, skip_image_content=False): ...

        @abstractmethod
        async def get_chunks_by_id(self, chunk_identifiers, app_id=None, skip_image_content=False): ...

        @abstractmethod
        async def delete_chunks_by_document_id(self, document_id, app_id=None): ...


def build_store_metrics(**kw) -> Dict[str, Any]:
    """Same keys as core/vector_store/utils.py:73-103 so ingestion telemetry keeps working."""
    base = dict(chunk_payload_upload_s=0.0, chunk_payload_objects=0, chunk_payload_bytes=0, chunk_payload_backend="memory",
                multivector_upload_s=0.0, multivector_objects=0, multivector_bytes=0, multivector_backend="b200-hbm",
                vector_store_write_s=0.0, vector_store_backend="b200-hbm", vector_store_rows=0, cache_write_s=0.0,
                cache_write_objects=0)
    base.update(kw)
    return base


def as_query_matrix(query_embedding) -> np.ndarray:
    """np.ndarray | torch.Tensor | List[np.ndarray] | List[torch.Tensor] | nested lists -> float32 [T,128]
    (accepted input forms: multi_vector_store.py:723,334-337; fast_multivector_store.py:506,515-518)."""
    try:
        import torch

        if isinstance(query_embedding, torch.Tensor):
            query_embedding = query_embedding.detach().float().cpu().numpy()
        elif isinstance(query_embedding, (list, tuple)) and len(query_embedding) and isinstance(query_embedding[0], torch.Tensor):
            query_embedding = np.stack([t.detach().float().cpu().numpy() for t in query_embedding])
    except ImportError:  # pragma: no cover
        pass
    q = np.asarray(query_embedding, dtype=np.float32)
    if q.ndim == 1:
        q = q[None, :]
    if q.ndim != 2 or q.shape[1] != 128:
        raise ValueError(f"query embedding must be [T,128], got {q.shape}")
    return np.ascontiguousarray(q)


def as_page_matrix(embedding):
    """Page embedding -> [P,128] matrix for ``add_pages``.

    CUDA torch tensors (what ColpaliEmbeddingModel holds before its ``.to(float32).numpy(force=True)``,
    colpali_embedding_model.py:262,291) stay on the device as bf16/float32 and go straight into the pack kernel -- the
    ingest fast path: no fp32 host copy, no Rust pack, no ``Bit`` objects, no ``executemany``
    (multi_vector_store.py:681-703).  Everything else becomes a host float32 array, as the reference does (:334-337)."""
    if hasattr(embedding, "detach"):
        t = embedding.detach()
        if t.is_cuda:
            import torch

            if t.dtype not in (torch.bfloat16, torch.float32):
                t = t.float()
            return t[None, :] if t.ndim == 1 else t
        embedding = t.float().cpu().numpy()
    emb = np.asarray(embedding, dtype=np.float32)
    return emb[None, :] if emb.ndim == 1 else emb


class _QueryRequest:
    __slots__ = ("q", "k", "doc_ids", "app_id", "future")

    def __init__(self, q, k, doc_ids, app_id, future):
        self.q, self.k, self.doc_ids, self.app_id, self.future = q, k, doc_ids, app_id, future


class _LoopQueue:
    """Pending query_similar requests of ONE event loop and the task that drains them."""
    __slots__ = ("pending", "task", "__weakref__")

    def __init__(self):
        self.pending = collections.deque()
        self.task = None


class QueryCoalescer:
    """Mixin: concurrent ``query_similar`` coroutines share GPU passes.  The host class provides
    ``_search_coalesced_locked(batch) -> List[List[DocumentChunk]]`` (runs in a worker thread) and calls
    ``_init_coalescer`` in its constructor."""

    def _init_coalescer(self, coalesce_queries: bool, max_coalesced_tokens: int, max_coalesced_queries: int) -> None:
        # 1024 tokens = one CTA-pair pass over the corpus
        self.coalesce_queries = bool(coalesce_queries)
        self.max_coalesced_tokens = int(max_coalesced_tokens)
        self.max_coalesced_queries = int(max_coalesced_queries)
        self._queues: "weakref.WeakKeyDictionary[Any, _LoopQueue]" = weakref.WeakKeyDictionary()
        self._queues_lock = threading.Lock()
        self.last_coalesced_batch = 0

    async def _enqueue_query(self, q: np.ndarray, k: int, doc_ids, app_id) -> List[DocumentChunk]:
        loop = asyncio.get_running_loop()
        with self._queues_lock:
            lq = self._queues.get(loop)
            if lq is None:
                lq = self._queues[loop] = _LoopQueue()
        req = _QueryRequest(q, int(k), doc_ids, app_id, loop.create_future())
        lq.pending.append(req)
        if lq.task is None or lq.task.done():
            lq.task = loop.create_task(self._drain(lq))
        return await req.future

    async def _drain(self, lq: _LoopQueue) -> None:
        """Score everything that is pending, one GPU pass per batch; requests arriving meanwhile form the next batch."""
        while lq.pending:
            batch, tokens = [], 0
            while lq.pending and len(batch) < self.max_coalesced_queries:
                nxt = lq.pending[0]
                if batch and tokens + nxt.q.shape[0] > self.max_coalesced_tokens:
                    break
                batch.append(lq.pending.popleft())
                tokens += nxt.q.shape[0]
            try:
                results = await asyncio.to_thread(self._search_coalesced_locked, batch)
            except BaseException as e:  # noqa: BLE001  (every waiter must be released; errors propagate like the reference's)
                for r in batch:
                    if not r.future.done():
                        r.future.set_exception(e if isinstance(e, Exception) else RuntimeError(repr(e)))
                if not isinstance(e, Exception):
                    raise
                continue
            self.last_coalesced_batch = len(batch)
            for r, res in zip(batch, results):
                if not r.future.done():
                    r.future.set_result(res)


class B200MultiVectorStore(QueryCoalescer, BaseVectorStore):
    """Exhaustive ColPali MaxSim store on one B200 (see module docstring)."""

    def __init__(self, uri: str = "b200://0", device: int = 0, mode: str = "bf16", storage: Any = None,
                 auto_initialize: bool = True, compact_dead_fraction: float = 0.3, index: Any = None,
                 fde_candidates: Optional[int] = None, coalesce_queries: bool = True, max_coalesced_tokens: int = 1024,
                 max_coalesced_queries: int = 64, zero_pad_compat: int = 0):
        self.uri = uri
        self.device = int(device)
        self.mode = mode
        self.storage = storage  # attribute kept for DualMultiVectorStore / DocumentService compatibility
        self.compact_dead_fraction = float(compact_dead_fraction)
        self.catalog = PageCatalog()
        self._index = index  # injectable for host-logic tests; the product builds a MaxSimIndex in initialize()
        # None: exhaustive MaxSim over every authorised page (Postgres-provider behaviour).  An int: two-stage search like
        # the "morphik" provider -- FDE candidates (the reference asks Turbopuffer for min(10*k, 75)) then MaxSim rerank.
        self.fde_candidates = fde_candidates
        self._two_stage = None
        # > 0: reproduce colpali_engine score_multi_vector's zero-padding quirk with this batch size (128 upstream): a page
        # shorter than the longest page of its scoring batch scores sum_t max(max_r <q_t,d_r>, 0)  (SURVEY App. A.2;
        # processing_colpali.py:350-362).  0 (default) = the clean MaxSim.
        self.zero_pad_compat = int(zero_pad_compat)
        self._init_coalescer(coalesce_queries, max_coalesced_tokens, max_coalesced_queries)  # see module docstring
        self._lock = threading.Lock()
        self._journal = None  # shardfile.StoreJournal when the store is durable (B200MultiVectorStore.open)
        self._last_store_metrics: Dict[str, Any] = {}
        self.last_query_timing: Dict[str, float] = {}
        if auto_initialize:
            self.initialize()

    # ------------------------------------------------------------------ lifecycle
    def initialize(self) -> bool:
        if self._index is None:
            from .index import MaxSimIndex  # raises when libb200ms.so or the GPU is missing: no silent fallback

            if self.fde_candidates:
                from .fde import TwoStageIndex

                self._two_stage = TwoStageIndex(device=self.device, dtype=self.mode)
                self._index = self._two_stage.index
            else:
                self._index = MaxSimIndex(device=self.device, dtype=self.mode)
        elif self.fde_candidates and self._two_stage is None and hasattr(self._index, "h"):
            # a loaded / injected CUDA index in two-stage mode: rebuild the FDE matrix from its packed rows
            from .fde import TwoStageIndex

            self._two_stage = TwoStageIndex(index=self._index)
            self._two_stage.rebuild_from_index()
        if self.zero_pad_compat and hasattr(self._index, "set_option"):
            self._index.set_option("zero_pad_compat", self.zero_pad_compat)
        return True

    def close(self) -> None:
        if self._index is not None and hasattr(self._index, "close"):
            self._index.close()
        self._index = None

    def latest_store_metrics(self) -> Dict[str, Any]:
        return dict(self._last_store_metrics)

    # ------------------------------------------------------------------ write path
    async def store_embeddings(self, chunks: List[DocumentChunk], app_id: Optional[str] = None
                               ) -> Tuple[bool, List[str], Dict[str, Any]]:
        valid = []
        for c in chunks:
            if getattr(c, "embedding", None) is None:
                logger.error("Missing embeddings for chunk %s-%s", c.document_id, c.chunk_number)
                continue
            emb = as_page_matrix(c.embedding)
            if emb.ndim != 2 or emb.shape[1] != 128:
                raise ValueError(f"chunk {c.document_id}-{c.chunk_number}: embedding must be [P,128], got {tuple(emb.shape)}")
            valid.append((c, emb))
        if not valid:
            self._last_store_metrics = build_store_metrics()
            return True, [], self._last_store_metrics
        t0 = time.perf_counter()
        await asyncio.to_thread(self._add_pages_locked, valid, app_id)
        dt = time.perf_counter() - t0
        self._last_store_metrics = build_store_metrics(
            vector_store_write_s=dt, vector_store_rows=len(valid), multivector_objects=len(valid), multivector_upload_s=dt,
            multivector_bytes=int(sum(e.shape[0] for _, e in valid)) * self._index.row_bytes)
        return True, [f"{c.document_id}-{c.chunk_number}" for c, _ in valid], self._last_store_metrics

    def _add_pages_locked(self, valid, app_id):
        with self._lock:
            first, n = (self._two_stage or self._index).add_pages([e for _, e in valid])
            recs = []
            for i, (c, e) in enumerate(valid):
                rec = PageRecord(c.document_id, int(c.chunk_number), c.content, dict(c.metadata or {}), app_id, int(e.shape[0]))
                pid = self.catalog.add(rec)
                assert pid == first + i, "catalogue and device corpus out of step"
                recs.append(rec)
            if self._journal is not None:  # durable store: this call becomes one append-only segment
                fde = self._two_stage.fde_rows(first, n) if self._two_stage is not None else None
                self._journal.log_add(self._index, first, n, [self._record_json(r) for r in recs], fde)

    @staticmethod
    def _record_json(r: PageRecord) -> Dict[str, Any]:
        return {"document_id": r.document_id, "chunk_number": r.chunk_number, "content": r.content, "metadata": r.metadata,
                "app_id": r.app_id, "n_rows": r.n_rows}

    # ------------------------------------------------------------------ read path
    async def query_similar(self, query_embedding, k: int, doc_ids: Optional[List[str]] = None,
                            app_id: Optional[str] = None, skip_image_content: bool = False) -> List[DocumentChunk]:
        if not self.coalesce_queries:
            results = await self.query_similar_batch([query_embedding], k, doc_ids, app_id, skip_image_content)
            return results[0]
        q = as_query_matrix(query_embedding)
        if k <= 0:
            return []
        return await self._enqueue_query(q, k, doc_ids, app_id)

    def _search_coalesced_locked(self, batch: List[_QueryRequest]) -> List[List[DocumentChunk]]:
        """One pass for a batch of independent requests; masks are built under the lock so they match the corpus."""
        t0 = time.perf_counter()
        with self._lock:
            n = len(self.catalog)
            out: List[Optional[List[DocumentChunk]]] = [None] * len(batch)
            live = []
            for i, r in enumerate(batch):
                visible, words = self.catalog.allow_words(r.doc_ids, r.app_id) if n else (False, None)
                if not visible:
                    out[i] = []
                else:
                    live.append((i, r, words))
            if live:
                kk = min(max(r.k for _, r, _ in live), n, 4096)
                if max(r.k for _, r, _ in live) > 4096:
                    logger.warning("k > 4096 requested; truncated to 4096 (B200MS_MAX_K)")
                queries = [r.q for _, r, _ in live]
                masks = [w for _, _, w in live]
                if self._two_stage is None and hasattr(self._index, "search_host_masked") and len(live) > 1:
                    ts, ti, tc = self._index.search_host_masked(queries, kk, masks)
                else:  # two-stage search / injected test index: one call per distinct mask
                    ts = np.full((len(live), kk), -np.inf, dtype=np.float32)
                    ti = np.full((len(live), kk), -1, dtype=np.int64)
                    tc = np.zeros((len(live),), dtype=np.int32)
                    groups: Dict[Optional[bytes], List[int]] = {}
                    for j, w in enumerate(masks):
                        groups.setdefault(None if w is None else w.tobytes(), []).append(j)
                    for idxs in groups.values():
                        a, b, c = self._search_unlocked([queries[j] for j in idxs], kk, masks[idxs[0]])
                        ts[idxs], ti[idxs], tc[idxs] = a, b, c
                for row, (i, r, _) in enumerate(live):
                    out[i] = self._chunks(ts[row], ti[row], min(int(tc[row]), r.k))
        self.last_query_timing = {"coalesced_queries": float(len(batch)), "total_ms": (time.perf_counter() - t0) * 1e3}
        return out  # type: ignore[return-value]

    def _chunks(self, scores, ids, count: int, min_score: Optional[float] = None) -> List[DocumentChunk]:
        hits = []
        for j in range(count):
            if min_score is not None and float(scores[j]) < min_score:
                break  # sorted by score descending
            rec = self.catalog.records[int(ids[j])]
            hits.append(DocumentChunk(document_id=rec.document_id, chunk_number=rec.chunk_number, content=rec.content,
                                      embedding=[], metadata=dict(rec.metadata), score=float(scores[j])))
        return hits

    async def query_similar_batch(self, query_embeddings: Sequence[Any], k: int, doc_ids: Optional[List[str]] = None,
                                  app_id: Optional[str] = None, skip_image_content: bool = False,
                                  min_score: Optional[float] = None) -> List[List[DocumentChunk]]:
        """Batched form (extension): one corpus pass scores every query of the batch.

        ``min_score`` drops hits scoring below it (the reference accepts ``min_score`` in retrieve_chunks but never applies
        it, document_service.py:381-383) -- scores are MaxSim sums, so the threshold is on that scale.
        Mask construction, the GPU search and the id -> DocumentChunk resolution run in ONE worker-thread call under the
        store lock, so an ingest or compaction can never slip between them."""
        queries = [as_query_matrix(q) for q in query_embeddings]
        if k <= 0:
            return [[] for _ in queries]
        return await asyncio.to_thread(self._search_batch_locked, queries, int(k), doc_ids, app_id, min_score)

    def _search_batch_locked(self, queries, k, doc_ids, app_id, min_score):
        t0 = time.perf_counter()
        with self._lock:
            n = len(self.catalog)
            if n == 0:
                return [[] for _ in queries]
            visible, words = self.catalog.allow_words(doc_ids, app_id)
            if not visible:
                return [[] for _ in queries]
            if k > 4096:
                logger.warning("k=%d requested; truncated to 4096 (B200MS_MAX_K)", k)
            kk = min(int(k), n, 4096)
            t1 = time.perf_counter()
            ts, ti, tc = self._search_unlocked(queries, kk, words)
            t2 = time.perf_counter()
            out = [self._chunks(ts[qi], ti[qi], int(tc[qi]), min_score) for qi in range(len(queries))]
            t3 = time.perf_counter()
        self.last_query_timing = {"prepare_ms": (t1 - t0) * 1e3, "gpu_search_ms": (t2 - t1) * 1e3,
                                  "build_chunks_ms": (t3 - t2) * 1e3, "total_ms": (t3 - t0) * 1e3}
        if self._two_stage is not None:  # the reference's stage names (fast_multivector_store.py:523,534,550,574,589,604)
            st = self._two_stage.last_timing_ms
            self.last_query_timing.update(encode_query_ms=st.get("encode_query_ms", 0.0), ns_query_ms=st.get("ns_query_ms", 0.0),
                                          load_multivectors_ms=0.0, rerank_scoring_ms=st.get("rerank_scoring_ms", 0.0),
                                          load_contents_ms=0.0)
        logger.debug("query_similar timing %s", json.dumps(self.last_query_timing))
        return out

    def _search_locked(self, queries, k, words):
        with self._lock:
            return self._search_unlocked(queries, k, words)

    def _search_unlocked(self, queries, k, words):
        if self._two_stage is not None:
            return self._two_stage.search(queries, k, n_candidates=max(int(self.fde_candidates), k), allow_mask=words)
        return self._index.search_host(queries, k, allow_mask=words)

    async def get_chunks_by_id(self, chunk_identifiers: List[Tuple[str, int]], app_id: Optional[str] = None,
                               skip_image_content: bool = False) -> List[DocumentChunk]:
        if not chunk_identifiers:
            return []
        out = []
        for doc_id, chunk_num in dict.fromkeys((d, int(n)) for d, n in chunk_identifiers):
            pid = self.catalog.lookup(doc_id, chunk_num)
            if pid is None:
                continue
            rec = self.catalog.records[pid]
            if app_id is not None and rec.app_id is not None and rec.app_id != app_id:
                continue
            out.append(DocumentChunk(document_id=rec.document_id, chunk_number=rec.chunk_number, content=rec.content,
                                     embedding=[], metadata=dict(rec.metadata), score=0.0))
        return out

    async def delete_chunks_by_document_id(self, document_id: str, app_id: Optional[str] = None) -> bool:
        try:  # off the event loop: the lock may be held by a GPU pass, and a compaction copies device memory
            await asyncio.to_thread(self._delete_locked, document_id)
            return True
        except Exception as e:  # noqa: BLE001  (reference returns False on error: multi_vector_store.py:949-951)
            logger.error("Error deleting chunks for document %s: %s", document_id, e)
            return False

    def _delete_locked(self, document_id: str) -> None:
        with self._lock:
            self.catalog.delete_document(document_id)
            if self._journal is not None:
                self._journal.log_delete(document_id)
            if self.catalog.dead_fraction > self.compact_dead_fraction:
                self._compact_locked()

    def _compact_locked(self):
        """Drop tombstoned pages: rebuild the device corpus from the surviving packed rows (device-to-device)."""
        keep, _ = self.catalog.compaction_plan()
        target = self._two_stage or self._index
        if hasattr(target, "compact"):
            target.compact(keep)
        self.catalog.apply_compaction(keep)

    # ------------------------------------------------------------------ persistence
    def save(self, directory: str) -> None:
        """Checkpoint: tombstoned pages are compacted away and the live corpus is written as ONE base segment of a journal
        directory (shardfile.StoreJournal); the legacy pair corpus.b2ms + catalog.jsonl is written too so that
        ``load`` of older callers keeps working."""
        import os

        from . import shardfile

        os.makedirs(directory, exist_ok=True)
        with self._lock:
            if self.catalog.dead_fraction > 0:
                self._compact_locked()
            shardfile.save_index(self._index, os.path.join(directory, "corpus.b2ms"))
            with open(os.path.join(directory, "catalog.jsonl"), "w") as f:
                for r in self.catalog.records:
                    f.write(json.dumps(self._record_json(r)) + "\n")
            if self._journal is not None and os.path.abspath(self._journal.dir) == os.path.abspath(directory):
                self._journal.reset()
                if len(self.catalog):
                    fde = self._two_stage.fde_rows(0, len(self.catalog)) if self._two_stage is not None else None
                    self._journal.log_add(self._index, 0, len(self.catalog), [self._record_json(r) for r in self.catalog.records], fde)

    @classmethod
    def load(cls, directory: str, device: int = 0, **kw) -> "B200MultiVectorStore":
        import os

        from . import shardfile

        index = shardfile.load_index(os.path.join(directory, "corpus.b2ms"), device=device)
        store = cls(device=device, mode=index.dtype_name, auto_initialize=False, index=index, **kw)
        with open(os.path.join(directory, "catalog.jsonl")) as f:
            for line in f:
                d = json.loads(line)
                store.catalog.add(PageRecord(d["document_id"], d["chunk_number"], d["content"], d["metadata"], d["app_id"], d["n_rows"]))
        if len(store.catalog) != index.n_pages:
            raise ValueError("catalogue and shard file disagree on the page count")
        store.initialize()  # two-stage mode: rebuilds the FDE matrix from the packed rows; applies zero_pad_compat
        return store

    @classmethod
    def open(cls, directory: str, device: int = 0, mode: str = "bf16", **kw) -> "B200MultiVectorStore":
        """Durable store: replay ``directory``'s journal (segments appended by earlier store_embeddings calls, tombstones of
        earlier deletes -- fast_multivector_store.py:673-707 / multi_vector_store.py:929-933 persist the same two operations)
        and keep journaling.  Nothing is ever rewritten in place; ``save(directory)`` checkpoints."""
        import os

        import torch

        from . import shardfile

        journal = shardfile.StoreJournal(directory)
        store = cls(device=device, mode=mode, **kw)
        for op in journal.read_ops():
            if op["op"] == "segment":
                seg, cat, fde = journal.segment_files(op["seq"])
                first = store._index.n_pages
                shardfile.load_index(seg, device=device, into=store._index)
                with open(cat) as f:
                    for line in f:
                        d = json.loads(line)
                        store.catalog.add(PageRecord(d["document_id"], d["chunk_number"], d["content"], d["metadata"], d["app_id"], d["n_rows"]))
                if store._two_stage is not None:
                    if os.path.exists(fde):
                        blob = torch.load(fde)
                        store._two_stage.append_fde_rows(blob["rows"].view(torch.bfloat16), blob["inv"])
                    else:
                        store._two_stage.rebuild_from_index()
                if store._index.n_pages - first != int(op["pages"]) or len(store.catalog) != store._index.n_pages:
                    raise ValueError(f"{seg}: segment and journal disagree on the page count")
            elif op["op"] == "delete":
                store.catalog.delete_document(op["document_id"])
        with store._lock:
            if store.catalog.dead_fraction > store.compact_dead_fraction:
                store._compact_locked()
        store._journal = journal
        return store
