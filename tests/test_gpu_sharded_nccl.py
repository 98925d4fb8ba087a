"""2-GPU (NCCL) test of the document-sharded search: every rank scans only its shard on its own B200, one all-gather of
the per-shard top-k, merge on the device -- and every rank must end up with the oracle's global top-k.
Skipped on boxes with fewer than 2 GPUs (the host logic is covered on CPU by tests/test_sharded_gloo.py)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("CUDA device required", allow_module_level=True)

import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from morphik_core_b200.index import MaxSimIndex
        from morphik_core_b200.sharded import ShardedMaxSim, plan_document_shards
        from oracle import maxsim_oracle as orc

        rng = np.random.default_rng(2024)  # identical corpus description on every rank
        doc_pages = rng.integers(1, 8, size=60).tolist()
        page_lens = [int(x) for x in rng.integers(1, 300, size=sum(doc_pages))]
        pages = []
        for n in page_lens:
            x = rng.standard_normal((n, 128)).astype(np.float32)
            pages.append(x / np.linalg.norm(x, axis=1, keepdims=True))
        queries = []
        for t in (32, 32, 12, 40):
            x = rng.standard_normal((t, 128)).astype(np.float32)
            queries.append(x / np.linalg.norm(x, axis=1, keepdims=True))
        first = np.concatenate([[0], np.cumsum(doc_pages)])
        doc_rows = [sum(page_lens[first[d]:first[d + 1]]) for d in range(len(doc_pages))]
        plan = plan_document_shards(doc_rows, world)
        p0, p1 = int(first[plan[rank][0]]), int(first[plan[rank][1]])

        idx = MaxSimIndex(device=rank, dtype=mode)
        idx.add_pages(pages[p0:p1])
        sharded = ShardedMaxSim.from_index(idx, id_base=p0)
        q = torch.from_numpy(np.concatenate(queries)).cuda()
        if rank != 0:
            q.zero_()
        q = sharded.broadcast_queries(q, src=0)
        k = 9
        ts, ti, tc = sharded.search(q, [len(x) for x in queries], k)
        torch.cuda.synchronize()
        ts, ti = ts.cpu().numpy(), ti.cpu().numpy()
        rows = np.concatenate(pages)
        off = orc.page_offsets(page_lens)
        for qi, qq in enumerate(queries):
            if mode == "bf16":
                want = orc.float_maxsim_c(orc.bf16_round_np(qq), orc.bf16_round_np(rows), off)
            else:
                want = orc.binary_maxsim_c(orc.sign_pack_c(qq), orc.sign_pack_c(rows), off)[0]
            ws, wi = orc.topk_np(want, k)
            assert ti[qi].tolist() == wi.tolist(), (rank, qi, ti[qi].tolist(), wi.tolist())
            np.testing.assert_allclose(ts[qi], ws, rtol=3e-5)
        open(os.path.join(result_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["bf16", "binary"])
def test_sharded_search_world2_nccl(tmp_path, mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mp.spawn(_worker, args=(2, _free_port(), mode, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]
