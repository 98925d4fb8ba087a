"""World-size-2 test of the multi-GPU host logic on CPU (gloo): shard plan (equal and rate-weighted), the exchange layout
the local top-k is written into, the single all-gather, and the merge order.  Local search and merge are oracle-backed
test doubles (the product runs the same layout through libb200ms: b200ms_sharded_search_begin/_end with NCCL --
morphik-core_b200/sharded.py:ShardedMaxSim.from_index)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from morphik_core_b200.sharded import (ShardedMaxSim, exchange_bytes, exchange_views, gathered_candidates,
                                       plan_document_shards)
from oracle import maxsim_oracle as orc


def test_plan_document_shards_balanced_and_contiguous():
    rows = [1024] * 100
    plan = plan_document_shards(rows, 8)
    assert plan[0][0] == 0 and plan[-1][1] == 100 and all(a[1] == b[0] for a, b in zip(plan, plan[1:]))
    sizes = [e - b for b, e in plan]
    assert max(sizes) - min(sizes) <= 1
    rng = np.random.default_rng(0)
    rows = rng.integers(1, 5000, size=997).tolist()
    plan = plan_document_shards(rows, 4)
    tot = [sum(rows[b:e]) for b, e in plan]
    assert sum(tot) == sum(rows) and max(tot) - min(tot) <= 5000
    assert plan_document_shards([10, 10], 4)[-1][1] == 2  # fewer documents than ranks: some ranges are empty
    assert plan_document_shards([], 2) == [(0, 0), (0, 0)]


def test_plan_document_shards_weighted_by_rank_speed():
    rows = [1024] * 1000
    plan = plan_document_shards(rows, 4, weights=[1.0, 0.8, 1.0, 1.2])  # rank 1 scans 20 % slower, rank 3 20 % faster
    sizes = [e - b for b, e in plan]
    assert sum(sizes) == 1000 and sizes[1] < sizes[0] < sizes[3] and abs(sizes[1] - 200) <= 1 and abs(sizes[3] - 300) <= 1
    with pytest.raises(ValueError):
        plan_document_shards(rows, 4, weights=[1, 1, 0, 1])


def test_exchange_layout_views_and_gather():
    """The local top-k is written IN PLACE into [n_q*k int64 ids][n_q*k f32 scores]; the gathered buffer is world such blocks."""
    assert exchange_bytes(1, 7) == 96 and exchange_bytes(32, 10) == 3840  # padded to 16: every gathered block stays aligned
    n_q, k, world = 3, 4, 2
    ids = torch.arange(24, dtype=torch.int64).reshape(world, n_q, k)
    sc = torch.arange(24, dtype=torch.float32).reshape(world, n_q, k) * 0.5
    bufs = []
    for w in range(world):
        b = torch.zeros(exchange_bytes(n_q, k), dtype=torch.uint8)
        iv, sv = exchange_views(b, n_q, k)
        iv.copy_(ids[w])
        sv.copy_(sc[w])
        assert b[: n_q * k * 8].view(torch.int64).tolist() == ids[w].reshape(-1).tolist()  # ids first, then scores
        bufs.append(b)
    ci, cs = gathered_candidates(torch.cat(bufs), world, n_q, k)
    assert ci.shape == (3, 8) and ci[1].tolist() == ids[0, 1].tolist() + ids[1, 1].tolist()
    assert cs[2].tolist() == sc[0, 2].tolist() + sc[1, 2].tolist()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_merge(gathered, world, n_q, k):
    cand_ids, cand_scores = gathered_candidates(gathered, world, n_q, k)
    ts = torch.full((n_q, k), float("-inf"))
    ti = torch.full((n_q, k), -1, dtype=torch.int64)
    tc = torch.zeros(n_q, dtype=torch.int32)
    for q in range(n_q):
        ids, s = cand_ids[q].numpy(), cand_scores[q].numpy()
        valid = ids >= 0
        order = np.lexsort((ids[valid], -s[valid].astype(np.float64)))[:k]
        ts[q, :len(order)] = torch.from_numpy(s[valid][order])
        ti[q, :len(order)] = torch.from_numpy(ids[valid][order])
        tc[q] = len(order)
    return ts, ti, tc


def _worker(rank, world, port, k, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(77)  # same corpus description on every rank
        doc_pages = rng.integers(1, 6, size=23).tolist()  # pages per document
        page_lens = [int(x) for x in rng.integers(1, 70, size=sum(doc_pages))]
        pages = [rng.standard_normal((n, 128)).astype(np.float32) for n in page_lens]
        queries = [rng.standard_normal((t, 128)).astype(np.float32) for t in (32, 9, 40)]
        doc_first_page = np.concatenate([[0], np.cumsum(doc_pages)])
        doc_rows = [sum(page_lens[doc_first_page[d]:doc_first_page[d + 1]]) for d in range(len(doc_pages))]
        plan = plan_document_shards(doc_rows, world)
        p0, p1 = int(doc_first_page[plan[rank][0]]), int(doc_first_page[plan[rank][1]])
        my_pages, my_lens = pages[p0:p1], page_lens[p0:p1]

        def local_search(q, q_lens, kk, ti, ts):  # oracle-backed stand-in for the CUDA search writing the exchange views
            ts.fill_(float("-inf"))
            ti.fill_(-1)
            rows = np.concatenate(my_pages) if my_pages else np.zeros((0, 128), np.float32)
            off = orc.page_offsets(my_lens)
            qoff = np.concatenate([[0], np.cumsum(q_lens)])
            for qi in range(len(q_lens)):
                s = orc.float_maxsim_c(q[qoff[qi]:qoff[qi + 1]].numpy(), rows, off)
                a, b = orc.topk_np(s, kk)
                ts[qi, :len(a)] = torch.from_numpy(a.astype(np.float32))
                ti[qi, :len(b)] = torch.from_numpy(b + p0)

        sharded = ShardedMaxSim(local_search, _oracle_merge)
        q = torch.from_numpy(np.concatenate(queries))
        if rank != 0:
            q = torch.zeros_like(q)
        q = sharded.broadcast_queries(q, src=0)
        ts, ti, tc = sharded.search(q, [len(x) for x in queries], k)
        # every rank must hold the global answer
        rows = np.concatenate(pages)
        off = orc.page_offsets(page_lens)
        for qi, qq in enumerate(queries):
            want_s, want_i = orc.topk_np(orc.float_maxsim_c(qq, rows, off), k)
            n = len(want_i)  # k may exceed the number of pages: unused slots are -1 / -inf
            assert int(tc[qi]) == n and ti[qi][:n].tolist() == want_i.tolist(), (rank, qi)
            assert torch.all(ti[qi][n:] == -1)
            np.testing.assert_allclose(ts[qi][:n].numpy(), want_s, rtol=1e-6)
        open(os.path.join(result_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("k", [5, 200])
def test_sharded_search_world2_gloo(tmp_path, k):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), k, str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]
