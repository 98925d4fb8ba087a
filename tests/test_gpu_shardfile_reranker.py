"""GPU tests of next rows f-2 (packed shard file round trip, importers of the reference's durable forms) and f-4 (MaxSim
reranker adapter)."""
import asyncio
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("CUDA device required", allow_module_level=True)

from morphik_core_b200 import shardfile  # noqa: E402
from morphik_core_b200.index import MaxSimIndex  # noqa: E402
from morphik_core_b200.models import DocumentChunk  # noqa: E402
from morphik_core_b200.reranker import B200MaxSimReranker  # noqa: E402
from oracle import maxsim_oracle as orc  # noqa: E402


def unit_rows(rng, n):
    x = rng.standard_normal((n, 128)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


@pytest.mark.parametrize("dtype", ["bf16", "int8", "binary"])
def test_shard_file_round_trip(tmp_path, dtype):
    rng = np.random.default_rng(0)
    lens = [int(x) for x in rng.integers(0, 300, size=80)] + [1030, 0, 1]
    pages = [unit_rows(rng, n) if n else np.zeros((0, 128), np.float32) for n in lens]
    queries = [unit_rows(rng, 32), unit_rows(rng, 11)]
    a = MaxSimIndex(dtype=dtype)
    a.add_pages(pages)
    path = str(tmp_path / f"shard_{dtype}.b2ms")
    nbytes = shardfile.save_index(a, path)
    assert nbytes == os.path.getsize(path)
    header, lens2, off = shardfile.read_header(path)
    assert header["dtype"] == dtype and lens2.tolist() == lens and off % 4096 == 0
    b = shardfile.load_index(path)
    assert b.n_pages == a.n_pages and torch.equal(a.packed_rows(), b.packed_rows())
    sa, sb = a.score_matrix(queries), b.score_matrix(queries)
    assert np.array_equal(sa, sb)
    ta, tb = a.search_host(queries, k=10), b.search_host(queries, k=10)
    assert np.array_equal(ta[1], tb[1]) and np.array_equal(ta[0], tb[0])
    with open(path, "r+b") as f:  # corruption is detected, not scanned
        f.truncate(os.path.getsize(path) - 100)
    with pytest.raises(ValueError, match="truncated"):
        shardfile.load_index(path)


def test_import_reference_durable_forms(tmp_path):
    rng = np.random.default_rng(1)
    pages = [unit_rows(rng, n) for n in (40, 64, 7)]
    # (a) per-page float32 .npy objects, as fast_multivector_store.py:673-707 writes them
    paths = []
    for i, p in enumerate(pages):
        path = tmp_path / f"multivector_doc_{i}.npy"
        np.save(path, p.astype(np.float32))
        paths.append(str(path))
    idx = MaxSimIndex(dtype="bf16")
    assert shardfile.import_npy_pages(idx, paths) == 3
    ref = MaxSimIndex(dtype="bf16")
    ref.add_pages(pages)
    assert torch.equal(idx.packed_rows(), ref.packed_rows())
    # (b) Postgres BIT(128)[] rows: packed bytes -> a binary corpus with exactly those bits
    bit_rows = [[bytes(r) for r in orc.sign_pack_c(p)] for p in pages]
    bidx = MaxSimIndex(dtype="binary")
    bidx.add_pages(shardfile.pages_from_bit_rows(bit_rows))
    bref = MaxSimIndex(dtype="binary")
    bref.add_pages(pages)
    assert torch.equal(bidx.packed_rows(), bref.packed_rows())


def test_maxsim_reranker_adapter():
    rng = np.random.default_rng(2)
    q = unit_rows(rng, 24)
    embs = [unit_rows(rng, n) for n in (50, 33, 64, 10)]
    embs[2][:24] = q  # chunk 2 contains the query tokens verbatim -> best
    chunks = [DocumentChunk(document_id=f"d{i}", content=f"c{i}", embedding=e, chunk_number=i) for i, e in enumerate(embs)]
    rr = B200MaxSimReranker(mode="bf16")
    out = asyncio.run(rr.rerank(q, chunks))
    want = orc.float_maxsim_c(orc.bf16_round_np(q), orc.bf16_round_np(np.concatenate(embs)), orc.page_offsets([50, 33, 64, 10]))
    assert [c.chunk_number for c in out] == np.argsort(-want, kind="stable").tolist() and out[0].chunk_number == 2
    np.testing.assert_allclose([c.score for c in out], np.sort(want)[::-1], rtol=3e-5)
    assert asyncio.run(rr.rerank(q, [])) == []
    kept = asyncio.run(rr.rerank(q, chunks, min_score=float(np.sort(want)[-2]) - 1e-3))  # threshold just under the 2nd best
    assert [c.chunk_number for c in kept] == [c.chunk_number for c in out[:2]]
    one = asyncio.run(rr.compute_score(q, embs[1]))
    many = asyncio.run(rr.compute_score(q, embs))
    assert abs(one - want[1]) < 1e-4 and np.allclose(many, want, rtol=3e-5)
