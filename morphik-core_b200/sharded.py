"""Document-sharded MaxSim search over the GPUs of one box.

Every page's score is independent of every other page, so the corpus partitions by document: rank r owns a contiguous
range of whole documents (balanced by patch rows), scans only its shard, and the ranks exchange nothing but their
per-query top-k lists: ONE NCCL all-gather of ``n_q * k * 12`` bytes per rank (int64 global page id + float32 score),
followed by a merge on every rank.  No page data ever crosses NVLink.  The reference has nothing comparable (it is a
single-process asyncio server); this is the multi-GPU extension the north star asks for.

One process per GPU (torchrun); ``torch.distributed`` is plumbing only.  The local search and the merge are injected so
that the host logic (shard plan, packing of the exchange buffer, gather, merge order) is testable on CPU with the
``gloo`` backend (tests/test_sharded_gloo.py); the product wiring uses ``MaxSimIndex.search_device`` and
``MaxSimIndex.merge_topk`` (CUDA).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def plan_document_shards(doc_rows: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Split documents 0..D-1 (doc_rows[d] = patch rows of document d) into `world` contiguous ranges with nearly equal
    row totals.  Returns [(doc_begin, doc_end)] per rank; ranges may be empty when D < world."""
    total = int(sum(int(x) for x in doc_rows))
    bounds = [0]
    acc = 0
    d = 0
    n = len(doc_rows)
    for r in range(1, world):
        target = total * r / world
        while d < n and acc + int(doc_rows[d]) / 2.0 <= target:
            acc += int(doc_rows[d])
            d += 1
        bounds.append(d)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def pack_exchange(ids: torch.Tensor, scores: torch.Tensor) -> torch.Tensor:
    """[n_q,k] int64 ids + [n_q,k] float32 scores -> one uint8 buffer (ids first), the unit of the all-gather."""
    n = ids.numel()
    buf = torch.empty(n * 12, dtype=torch.uint8, device=ids.device)
    buf[: n * 8].view(torch.int64).copy_(ids.reshape(-1))
    buf[n * 8:].view(torch.float32).copy_(scores.reshape(-1))
    return buf


def unpack_exchange(gathered: torch.Tensor, world: int, n_q: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """[world, n_q*k*12] uint8 -> candidate ids [n_q, world*k] int64 and scores [n_q, world*k] float32 (rank-major)."""
    n = n_q * k
    g = gathered.view(world, n * 12)
    ids = g[:, : n * 8].contiguous().view(torch.int64).view(world, n_q, k)
    sc = g[:, n * 8:].contiguous().view(torch.float32).view(world, n_q, k)
    return (ids.permute(1, 0, 2).reshape(n_q, world * k).contiguous(),
            sc.permute(1, 0, 2).reshape(n_q, world * k).contiguous())


LocalSearch = Callable[[torch.Tensor, Sequence[int], int], Tuple[torch.Tensor, torch.Tensor]]
Merge = Callable[[torch.Tensor, torch.Tensor, int], Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]


class ShardedMaxSim:
    """Per-rank object: local shard search + all-gather + merge.  ``search`` returns the same result on every rank."""

    def __init__(self, local_search: LocalSearch, merge: Merge, group: Optional[dist.ProcessGroup] = None):
        self.local_search = local_search
        self.merge = merge
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    @classmethod
    def from_index(cls, index, id_base: int, group: Optional[dist.ProcessGroup] = None) -> "ShardedMaxSim":
        """Product wiring: `index` is this rank's MaxSimIndex over its shard; id_base = global id of its first page."""

        def local_search(q_dev, q_lens, k):
            ts, ti, _ = index.search_device(q_dev, q_lens, k, id_base=id_base)
            return ts, ti

        return cls(local_search, index.merge_topk, group)

    def broadcast_queries(self, q: torch.Tensor, src: int = 0) -> torch.Tensor:
        if self.world > 1:
            dist.broadcast(q, src=src, group=self.group)
        return q

    def search(self, q: torch.Tensor, q_lens: Sequence[int], k: int):
        """q: [sum T,128] on this rank's device (identical on all ranks -- see broadcast_queries).
        Returns (scores [n_q,k], global page ids [n_q,k], counts [n_q])."""
        n_q = len(q_lens)
        ts, ti = self.local_search(q, q_lens, k)
        if self.world == 1:
            return self.merge(ts, ti, k)
        mine = pack_exchange(ti, ts)
        gathered = torch.empty(self.world * mine.numel(), dtype=torch.uint8, device=mine.device)
        dist.all_gather_into_tensor(gathered, mine, group=self.group)  # the one collective of the path
        cand_ids, cand_scores = unpack_exchange(gathered, self.world, n_q, k)
        return self.merge(cand_scores, cand_ids, k)
