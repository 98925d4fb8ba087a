"""ShardedB200MultiVectorStore -- the BaseVectorStore plugin over all GPUs of one box (SURVEY 8e at the plugin level).

One process per GPU (torchrun).  Rank 0 is the process the Morphik API server talks to: it exposes the usual four async
methods (core/vector_store/base_vector_store.py:7-65).  Ranks > 0 call ``worker_loop()`` and execute the same command
stream.  Documents are assigned to ranks whole (least-loaded rank at first sight), so the ``doc_ids`` filter and
``delete_chunks_by_document_id`` stay rank-local and no page data ever moves between GPUs after ingest.

Concurrent query_similar coroutines are coalesced on rank 0 (store.QueryCoalescer) into one command per GPU pass.
A query command is: rank 0 broadcasts (query rows, k, one (doc_ids, app_id) filter PER QUERY) -> every rank builds its own
page masks (one row per distinct filter), scans its shard and selects its top-k with GLOBAL ids ``(rank << 40) | local_page`` -> ONE all-gather of n_q*k*12 bytes per rank -> merge on
every rank (``ShardedMaxSim``) -> rank 0 turns ids into DocumentChunks.  Rank 0 mirrors every rank's catalogue (payloads
and metadata live only there); the catalogues evolve deterministically from the command stream, including compactions.

The index factory and the collective helpers are injectable so the host logic runs under ``gloo`` on CPU with
oracle-backed stand-ins (tests/test_sharded_store_gloo.py); the product wiring is MaxSimIndex + NCCL.
"""
from __future__ import annotations

import asyncio
import logging
import threading
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .catalog import PageCatalog, PageRecord
from .models import DocumentChunk
from .sharded import ShardedMaxSim
from .store import BaseVectorStore, QueryCoalescer, _QueryRequest, as_query_matrix, build_store_metrics

logger = logging.getLogger(__name__)
RANK_SHIFT = 40


def split_global_id(gid: int) -> Tuple[int, int]:
    return int(gid) >> RANK_SHIFT, int(gid) & ((1 << RANK_SHIFT) - 1)


class ShardedB200MultiVectorStore(QueryCoalescer, BaseVectorStore):
    def __init__(self, mode: str = "bf16", device: Optional[int] = None, group: Optional[dist.ProcessGroup] = None,
                 index_factory: Optional[Callable[[], Any]] = None, compact_dead_fraction: float = 0.3,
                 coalesce_queries: bool = True, max_coalesced_tokens: int = 1024, max_coalesced_queries: int = 64):
        if not dist.is_initialized():
            raise RuntimeError("ShardedB200MultiVectorStore needs an initialised torch.distributed process group (torchrun)")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.mode = mode
        self.uri = f"b200-sharded://{self.world}"
        self.storage = None
        self.compact_dead_fraction = float(compact_dead_fraction)
        if index_factory is None:
            from .index import MaxSimIndex  # CUDA product path; raises without a GPU

            dev = self.rank if device is None else int(device)
            index_factory = lambda: MaxSimIndex(device=dev, dtype=mode)  # noqa: E731
        self.index = index_factory()
        self.data_device = getattr(self.index, "device", torch.device("cpu"))
        # catalogue of THIS rank's pages (rank 0 additionally mirrors everybody's, with payloads)
        self.catalogs: Dict[int, PageCatalog] = {r: PageCatalog() for r in (range(self.world) if self.rank == 0 else [self.rank])}
        self.doc_rank: Dict[str, int] = {}
        self.rank_rows = [0] * self.world
        self._lock = threading.RLock()  # re-entrant: a query holds it across the command AND the catalogue look-ups
        self._sharded = ShardedMaxSim(self._local_search, self._merge, group)
        self._init_coalescer(coalesce_queries, max_coalesced_tokens, max_coalesced_queries)
        self._mask_matrix: Optional[torch.Tensor] = None  # [n_distinct_filters, words] int32, this rank's pages
        self._mask_index: Optional[torch.Tensor] = None   # int32 [n_q]: row of _mask_matrix per query, -1 = unfiltered
        self._skip_local = False
        self._pending_rows: Optional[np.ndarray] = None

    # ------------------------------------------------------------------ collective plumbing
    def _bcast_obj(self, obj):
        box = [obj]
        dist.broadcast_object_list(box, src=0, group=self.group)
        return box[0]

    def _bcast_rows(self, rows: Optional[np.ndarray], n_rows: int) -> torch.Tensor:
        t = torch.empty((n_rows, 128), dtype=torch.float32, device=self.data_device)
        if self.rank == 0:
            t.copy_(torch.from_numpy(np.ascontiguousarray(rows, dtype=np.float32)))
        if n_rows:
            dist.broadcast(t, src=0, group=self.group)
        return t

    def _local_search(self, q: torch.Tensor, q_lens: Sequence[int], k: int):
        if self._skip_local:  # empty shard or nothing authorised here: contribute an empty list to the collective
            return (torch.full((len(q_lens), k), float("-inf"), device=self.data_device),
                    torch.full((len(q_lens), k), -1, dtype=torch.int64, device=self.data_device))
        if self._mask_matrix is None:
            ts, ti, _ = self.index.search_device(q, list(q_lens), k, id_base=self.rank << RANK_SHIFT)
        else:
            ts, ti, _ = self.index.search_device(q, list(q_lens), k, allow_mask_dev=self._mask_matrix,
                                                 id_base=self.rank << RANK_SHIFT, mask_index_dev=self._mask_index)
        return ts, ti

    def _merge(self, cand_scores, cand_ids, k):
        return self.index.merge_topk(cand_scores, cand_ids, k)

    # ------------------------------------------------------------------ command execution (identical on every rank)
    def _execute(self, cmd: Tuple) -> Any:
        op = cmd[0]
        if op == "add":
            _, owner, recs, lens, n_rows = cmd
            rows = self._bcast_rows(self._pending_rows, n_rows)
            self._pending_rows = None
            for (doc, num, app) in recs:
                self.doc_rank[doc] = owner
            self.rank_rows[owner] += int(sum(lens))
            if owner == self.rank:
                off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
                self.index.add_pages([rows[off[i]:off[i + 1]] for i in range(len(lens))])
            if owner in self.catalogs and self.rank != 0:
                for (doc, num, app), n in zip(recs, lens):
                    self.catalogs[owner].add(PageRecord(doc, num, "", {}, app, n))
            return None
        if op == "query":
            _, q_lens, k, filters = cmd
            q = self._bcast_rows(self._pending_rows, int(sum(q_lens)))
            self._pending_rows = None
            cat = self.catalogs[self.rank]
            n = len(cat)
            rows, index, seen, any_visible = [], [], {}, False
            for doc_ids, app_id in filters:
                visible, words = cat.allow_words(doc_ids, app_id) if n else (False, None)
                if not visible:  # nothing authorised on this rank: an all-zero mask keeps the query out of the local top-k
                    words = np.zeros((n + 31) // 32, dtype=np.uint32)
                else:
                    any_visible = True
                if words is None:
                    index.append(-1)
                    continue
                key = words.tobytes()
                if key not in seen:
                    seen[key] = len(rows)
                    rows.append(words)
                index.append(seen[key])
            self._skip_local = n == 0 or not any_visible
            if rows and not self._skip_local:
                self._mask_matrix = torch.from_numpy(np.stack(rows).view(np.int32)).to(self.data_device)
                self._mask_index = torch.tensor(index, dtype=torch.int32, device=self.data_device)
            else:
                self._mask_matrix = self._mask_index = None
            kk = max(1, min(int(k), 4096))
            return self._sharded.search(q, q_lens, kk)
        if op == "delete":
            _, document_id = cmd
            owner = self.doc_rank.pop(document_id, None)
            if owner is None:
                return False
            cat = self.catalogs.get(owner)  # rank 0 mirrors every rank; rank r only knows its own shard
            if cat is not None:
                freed = sum(cat.records[p].n_rows for p in cat.pages_of(document_id))
                cat.delete_document(document_id)
                self.rank_rows[owner] -= freed
                if cat.dead_fraction > self.compact_dead_fraction:  # same decision on the owner and on rank 0's mirror
                    keep, _ = cat.compaction_plan()
                    if owner == self.rank:
                        self.index.compact(keep)
                    cat.apply_compaction(keep)
            return True
        if op == "stop":
            return "stop"
        raise ValueError(f"unknown command {op!r}")

    def worker_loop(self) -> None:
        """Ranks > 0: execute rank 0's command stream until close()."""
        assert self.rank != 0, "rank 0 drives the store through the BaseVectorStore methods"
        while True:
            cmd = self._bcast_obj(None)
            if self._execute(cmd) == "stop":
                return

    def _drive(self, cmd: Tuple, rows: Optional[np.ndarray] = None) -> Any:
        """Rank 0: publish a command (and its row payload) and execute it locally."""
        with self._lock:
            self._pending_rows = rows
            self._bcast_obj(cmd)
            return self._execute(cmd)

    # ------------------------------------------------------------------ BaseVectorStore surface (rank 0)
    def initialize(self) -> bool:
        return True

    def close(self) -> None:
        if self.rank == 0 and dist.is_initialized():
            self._drive(("stop",))
        if hasattr(self.index, "close"):
            self.index.close()

    async def store_embeddings(self, chunks: List[DocumentChunk], app_id: Optional[str] = None
                               ) -> Tuple[bool, List[str], Dict[str, Any]]:
        assert self.rank == 0
        by_doc: Dict[str, List[Tuple[DocumentChunk, np.ndarray]]] = {}
        for c in chunks:
            if getattr(c, "embedding", None) is None:
                logger.error("Missing embeddings for chunk %s-%s", c.document_id, c.chunk_number)
                continue
            emb = np.asarray(c.embedding.detach().float().cpu().numpy() if hasattr(c.embedding, "detach") else c.embedding,
                             dtype=np.float32).reshape(-1, 128)
            by_doc.setdefault(c.document_id, []).append((c, emb))
        ids: List[str] = []
        for doc, items in by_doc.items():
            owner = self.doc_rank.get(doc)
            if owner is None:
                owner = int(np.argmin(self.rank_rows))  # whole documents go to the least-loaded rank
            lens = [int(e.shape[0]) for _, e in items]
            recs = [(c.document_id, int(c.chunk_number), app_id) for c, _ in items]
            rows = np.concatenate([e for _, e in items]) if sum(lens) else np.zeros((0, 128), np.float32)
            for (c, e) in items:  # rank 0's mirror carries the payloads
                self.catalogs[owner].add(PageRecord(c.document_id, int(c.chunk_number), c.content, dict(c.metadata or {}), app_id,
                                                    int(e.shape[0])))
            await asyncio.to_thread(self._drive, ("add", owner, recs, lens, int(sum(lens))), rows)
            ids.extend(f"{c.document_id}-{c.chunk_number}" for c, _ in items)
        return True, ids, build_store_metrics(vector_store_rows=len(ids), vector_store_backend=f"b200-hbm x{self.world}")

    async def query_similar(self, query_embedding, k: int, doc_ids: Optional[List[str]] = None,
                            app_id: Optional[str] = None, skip_image_content: bool = False) -> List[DocumentChunk]:
        assert self.rank == 0
        q = as_query_matrix(query_embedding)
        if k <= 0 or all(len(c) == 0 for c in self.catalogs.values()):
            return []
        if self.coalesce_queries:
            return await self._enqueue_query(q, k, doc_ids, app_id)
        return (await asyncio.to_thread(self._search_coalesced_locked, [_QueryRequest(q, int(k), doc_ids, app_id, None)]))[0]

    def _search_coalesced_locked(self, batch) -> List[List[DocumentChunk]]:
        """One command (= one pass on every GPU) for a batch of independent requests."""
        q_lens = [int(r.q.shape[0]) for r in batch]
        kmax = max(r.k for r in batch)
        rows = np.concatenate([r.q for r in batch]) if len(batch) > 1 else batch[0].q
        out = []
        with self._lock:
            ts, ti, tc = self._drive(("query", q_lens, kmax, [(r.doc_ids, r.app_id) for r in batch]), rows)
            ts, ti, tc = ts.cpu().numpy(), ti.cpu().numpy(), tc.cpu().numpy()
            for i, r in enumerate(batch):
                hits = []
                for j in range(min(int(tc[i]), r.k)):
                    rk, local = split_global_id(ti[i, j])
                    rec = self.catalogs[rk].records[local]
                    hits.append(DocumentChunk(document_id=rec.document_id, chunk_number=rec.chunk_number, content=rec.content,
                                              embedding=[], metadata=dict(rec.metadata), score=float(ts[i, j])))
                out.append(hits)
        return out

    async def get_chunks_by_id(self, chunk_identifiers: List[Tuple[str, int]], app_id: Optional[str] = None,
                               skip_image_content: bool = False) -> List[DocumentChunk]:
        assert self.rank == 0
        out = []
        for doc_id, num in dict.fromkeys((d, int(n)) for d, n in chunk_identifiers):
            r = self.doc_rank.get(doc_id)
            pid = None if r is None else self.catalogs[r].lookup(doc_id, num)
            if pid is None:
                continue
            rec = self.catalogs[r].records[pid]
            if app_id is not None and rec.app_id is not None and rec.app_id != app_id:
                continue
            out.append(DocumentChunk(document_id=rec.document_id, chunk_number=rec.chunk_number, content=rec.content,
                                     embedding=[], metadata=dict(rec.metadata), score=0.0))
        return out

    async def delete_chunks_by_document_id(self, document_id: str, app_id: Optional[str] = None) -> bool:
        assert self.rank == 0
        try:
            await asyncio.to_thread(self._drive, ("delete", document_id))
            return True
        except Exception as e:  # noqa: BLE001
            logger.error("Error deleting chunks for document %s: %s", document_id, e)
            return False
