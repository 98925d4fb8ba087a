"""World-size-2 CPU (gloo) test of ShardedB200MultiVectorStore: rank 0 drives the BaseVectorStore API, rank 1 runs the
worker loop; documents are placed whole on ranks, queries/filters/deletes give exactly what one oracle-scored store would.
The per-rank index is an oracle-backed stand-in with the MaxSimIndex methods the sharded store uses (the product
constructs the CUDA MaxSimIndex; tests/test_gpu_sharded_nccl.py covers that wiring on 2 GPUs)."""
import asyncio
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import maxsim_oracle as orc


class OracleShardIndex:
    """CPU stand-in for MaxSimIndex (float MaxSim, fp32): add_pages / compact / search_device / merge_topk."""

    device = torch.device("cpu")

    def __init__(self):
        self.pages = []

    def add_pages(self, pages):
        first = len(self.pages)
        new = [np.asarray(p, dtype=np.float32) for p in pages]
        if any(len(p) and float(p[0, 0]) == 12345.0 for p in new):  # test hook: "this rank cannot take these pages"
            raise MemoryError("simulated allocation failure")
        self.pages.extend(new)
        return first, len(pages)

    def compact(self, keep):
        self.pages = [self.pages[int(i)] for i in keep]

    def search_device(self, q, q_lens, k, allow_mask_dev=None, id_base=0, mask_index_dev=None):
        lens = [len(p) for p in self.pages]
        rows = np.concatenate(self.pages) if sum(lens) else np.zeros((0, 128), np.float32)
        off = orc.page_offsets(lens)

        def allow_of(i):  # one mask for everybody, or a row of the mask matrix per query (-1 = unfiltered)
            if allow_mask_dev is None:
                return None
            words = allow_mask_dev.numpy().view(np.uint32)
            if mask_index_dev is not None:
                mi = int(mask_index_dev[i])
                if mi < 0:
                    return None
                words = words[mi]
            return np.unpackbits(np.ascontiguousarray(words).view(np.uint8), bitorder="little")[: len(lens)].astype(bool)

        ts = torch.full((len(q_lens), k), float("-inf"))
        ti = torch.full((len(q_lens), k), -1, dtype=torch.int64)
        tc = torch.zeros(len(q_lens), dtype=torch.int32)
        qo = np.concatenate([[0], np.cumsum(q_lens)])
        for i in range(len(q_lens)):
            s = orc.float_maxsim_c(q[qo[i]:qo[i + 1]].numpy(), rows, off)
            a, b = orc.topk_np(s, k, allow_of(i))
            ts[i, :len(a)] = torch.from_numpy(a.astype(np.float32))
            ti[i, :len(b)] = torch.from_numpy(b + id_base)
            tc[i] = len(a)
        return ts, ti, tc

    def merge_topk(self, cand_scores, cand_ids, k):
        n_q = cand_scores.shape[0]
        ts = torch.full((n_q, k), float("-inf"))
        ti = torch.full((n_q, k), -1, dtype=torch.int64)
        tc = torch.zeros(n_q, dtype=torch.int32)
        for q in range(n_q):
            ids, s = cand_ids[q].numpy(), cand_scores[q].numpy()
            v = ids >= 0
            order = np.lexsort((ids[v], -s[v].astype(np.float64)))[:k]
            ts[q, :len(order)] = torch.from_numpy(s[v][order])
            ti[q, :len(order)] = torch.from_numpy(ids[v][order])
            tc[q] = len(order)
        return ts, ti, tc


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from morphik_core_b200.models import DocumentChunk
        from morphik_core_b200.sharded_store import ShardedB200MultiVectorStore, split_global_id

        store = ShardedB200MultiVectorStore(mode="bf16", index_factory=OracleShardIndex, compact_dead_fraction=0.3)
        if rank != 0:
            store.worker_loop()
            open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
            return
        run = asyncio.run
        rng = np.random.default_rng(5)
        docs = {f"doc{d}": [rng.standard_normal((int(rng.integers(3, 40)), 128)).astype(np.float32) for _ in range(int(rng.integers(1, 5)))]
                for d in range(12)}
        for d, pages in docs.items():
            ok, ids, _ = run(store.store_embeddings([DocumentChunk(document_id=d, content=f"{d}/{j}", embedding=p, chunk_number=j,
                                                                   metadata={"j": j}) for j, p in enumerate(pages)], app_id="app"))
            assert ok and ids == [f"{d}-{j}" for j in range(len(pages))]
        # whole documents per rank, both ranks used, load roughly balanced
        owners = {d: store.doc_rank[d] for d in docs}
        assert set(owners.values()) == {0, 1}
        assert abs(store.rank_rows[0] - store.rank_rows[1]) <= 160

        def oracle(q, k, allowed_docs=None):
            flat = [(d, j, p) for d, pages in docs.items() for j, p in enumerate(pages) if allowed_docs is None or d in allowed_docs]
            rows = np.concatenate([p for _, _, p in flat])
            s = orc.float_maxsim_c(q, rows, orc.page_offsets([len(p) for _, _, p in flat]))
            order = np.argsort(-s.astype(np.float64), kind="stable")[:k]
            return [(flat[i][0], flat[i][1]) for i in order], s[order]

        q = docs["doc7"][0] + 0.01 * rng.standard_normal(docs["doc7"][0].shape).astype(np.float32)
        res = run(store.query_similar(q, k=6))
        want, ws = oracle(q, 6)
        assert [(r.document_id, r.chunk_number) for r in res] == want and res[0].document_id == "doc7"
        np.testing.assert_allclose([r.score for r in res], ws, rtol=1e-6)
        assert res[0].content == "doc7/0" and res[0].metadata == {"j": 0} and res[0].embedding == []
        # doc_ids filter that lives entirely on ONE rank, and one that matches nothing
        some = [d for d in docs if owners[d] == 1][:2]
        res = run(store.query_similar(q, k=20, doc_ids=some))
        assert {r.document_id for r in res} <= set(some) and [(r.document_id, r.chunk_number) for r in res] == oracle(q, 20, set(some))[0]
        assert run(store.query_similar(q, k=5, doc_ids=["nope"])) == []
        assert run(store.query_similar(q, k=5, app_id="other")) == []
        # concurrent callers with different filters and k share ONE command (one pass per rank) and get their own answers
        on0 = [d for d in docs if owners[d] == 0][:2]
        reqs = [dict(query_embedding=q, k=6), dict(query_embedding=q, k=20, doc_ids=some), dict(query_embedding=q, k=3, doc_ids=on0),
                dict(query_embedding=docs["doc1"][0], k=4, doc_ids=some + on0), dict(query_embedding=q, k=5, doc_ids=["nope"]),
                dict(query_embedding=q, k=2, app_id="app")]
        lone = [run(store.query_similar(**r)) for r in reqs]

        async def many():
            return await asyncio.gather(*[store.query_similar(**r) for r in reqs])

        together = run(many())
        key = lambda rs: [(r.document_id, r.chunk_number, r.score) for r in rs]  # noqa: E731
        assert [key(r) for r in together] == [key(r) for r in lone] and together[4] == [] and store.last_coalesced_batch == len(reqs)
        assert [(r.document_id, r.chunk_number) for r in together[2]] == oracle(q, 3, set(on0))[0]
        # deletes (with compaction on the owning rank) keep results identical to the oracle over the survivors
        for d in ("doc7", "doc2", "doc3", "doc9", "doc11"):
            assert run(store.delete_chunks_by_document_id(d)) is True
            docs.pop(d)
        res = run(store.query_similar(q, k=8))
        assert [(r.document_id, r.chunk_number) for r in res] == oracle(q, 8)[0]
        assert run(store.get_chunks_by_id([("doc7", 0)])) == [] and len(run(store.get_chunks_by_id([("doc1", 0), ("doc1", 0)]))) == 1
        assert split_global_id((1 << 40) | 17) == (1, 17)
        # a k beyond the merge capacity (world * k <= 8192) is truncated, not fatal (ADVICE: large-k query killed the store)
        res = run(store.query_similar(q, k=5000))
        assert len(res) == sum(len(p) for p in docs.values())
        # a failure on ONE rank (here: the owner's add_pages) surfaces on rank 0 and the command stream stays in step
        from morphik_core_b200.sharded_store import ShardedStoreError

        bad = np.full((4, 128), 12345.0, dtype=np.float32)
        for attempt in range(2):  # whichever rank is least loaded owns it: both ranks' failures must be survivable
            with pytest.raises(ShardedStoreError):
                run(store.store_embeddings([DocumentChunk(document_id=f"bad{attempt}", content="x", embedding=bad, chunk_number=0)]))
        res = run(store.query_similar(q, k=8))
        assert [(r.document_id, r.chunk_number) for r in res] == oracle(q, 8)[0]
        assert "bad0" not in store.doc_rank and all(c.lookup("bad0", 0) is None for c in store.catalogs.values())
        # concurrent writers and readers (ADVICE high: mirror vs index order): catalogue order == index order afterwards
        new_docs = {f"new{d}": [rng.standard_normal((int(rng.integers(3, 30)), 128)).astype(np.float32) for _ in range(2)] for d in range(8)}

        async def storm():
            writes = [store.store_embeddings([DocumentChunk(document_id=d, content=f"{d}/{j}", embedding=p, chunk_number=j, metadata={"j": j})
                                              for j, p in enumerate(ps)], app_id="app") for d, ps in new_docs.items()]
            reads = [store.query_similar(q, k=5) for _ in range(6)]
            return await asyncio.gather(*writes, *reads)

        run(storm())
        docs.update(new_docs)
        for d, ps in new_docs.items():
            r = store.doc_rank[d]
            for j, p in enumerate(ps):
                pid = store.catalogs[r].lookup(d, j)
                assert pid is not None and store.catalogs[r].records[pid].n_rows == len(p)
        for name in ("new3", "new6"):  # documents that landed on either rank resolve to their own payloads
            q2 = new_docs[name][1]
            res = run(store.query_similar(q2, k=9))
            assert [(r.document_id, r.chunk_number) for r in res] == oracle(q2, 9)[0] and res[0].content == f"{name}/1"
        assert {store.doc_rank[d] for d in new_docs} == {0, 1}
        store.close()
        open(os.path.join(out_dir, "ok0"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_sharded_store_world2_gloo(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]
