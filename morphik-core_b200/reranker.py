"""MaxSim reranker adapter.

The reference's ``core/reranker`` is a TEXT cross-encoder (``BaseReranker.rerank(query: str, chunks)`` /
``compute_score(query: str, text)``, core/reranker/base_reranker.py:7-26) and is bypassed when ColPali retrieval is on
(core/services/document_service.py:381-383).  A late-interaction reranker needs embeddings, not strings, so
this adapter keeps the method names and result conventions (chunks come back sorted by descending ``score``) but takes the
query's multivector: "retrieve with any store, rerank with ColPali MaxSim on the GPU".
"""
from __future__ import annotations

import asyncio
import threading
from typing import List, Optional, Sequence, Union

import numpy as np

from .index import MaxSimIndex
from .models import DocumentChunk
from .store import as_page_matrix, as_query_matrix


class B200MaxSimReranker:
    def __init__(self, device: int = 0, mode: str = "bf16"):
        self.device, self.mode = int(device), mode
        self._idx: Optional[MaxSimIndex] = None
        self._lock = threading.Lock()

    def _scores(self, query_embedding, page_embeddings: Sequence) -> np.ndarray:
        """One scratch index is reused across calls (its device buffer and native handle stay); candidate embeddings may be
        host arrays / lists or CUDA tensors (those stay on the device, store.as_page_matrix)."""
        with self._lock:
            if self._idx is None:
                self._idx = MaxSimIndex(device=self.device, dtype=self.mode)
            idx = self._idx
            idx.clear()
            idx.add_pages([as_page_matrix(e) for e in page_embeddings])
            return idx.score_matrix([as_query_matrix(query_embedding)])[0]

    def close(self) -> None:
        with self._lock:
            if self._idx is not None:
                self._idx.close()
                self._idx = None

    async def rerank(self, query_embedding, chunks: List[DocumentChunk], min_score: Optional[float] = None
                     ) -> List[DocumentChunk]:
        """Chunks must carry their multivector in ``embedding``; returns them sorted by MaxSim score (descending, stable).
        ``min_score`` (optional) drops chunks scoring below it -- the threshold retrieve_chunks accepts but never applies
        (document_service.py:381-383)."""
        if not chunks:
            return []
        scores = await asyncio.to_thread(self._scores, query_embedding, [c.embedding for c in chunks])  # GPU work off the loop
        order = sorted(range(len(chunks)), key=lambda i: (-scores[i], i))
        out = []
        for i in order:
            if min_score is not None and scores[i] < min_score:
                break
            c = chunks[i].model_copy() if hasattr(chunks[i], "model_copy") else chunks[i]
            c.score = float(scores[i])
            out.append(c)
        return out

    async def compute_score(self, query_embedding, page_embedding: Union[np.ndarray, Sequence]) -> Union[float, List[float]]:
        """One page ([P,128]) -> float; a list of pages -> list of floats (mirrors BaseReranker.compute_score's two forms)."""
        if isinstance(page_embedding, (list, tuple)) and len(page_embedding) and np.ndim(page_embedding[0]) == 2:
            return [float(s) for s in await asyncio.to_thread(self._scores, query_embedding, page_embedding)]
        return float((await asyncio.to_thread(self._scores, query_embedding, [page_embedding]))[0])
