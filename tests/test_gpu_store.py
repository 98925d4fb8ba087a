"""The reference's own store tests (core/tests/unit/test_multivector.py), re-run against B200MultiVectorStore on a GPU.
Same scenarios and assertions; the Postgres fixture is replaced by the in-HBM store.  mode="binary" reproduces the Postgres
provider's scores exactly; mode="bf16" the float MaxSim of the "morphik" provider."""
import asyncio

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("CUDA device required", allow_module_level=True)

from morphik_core_b200.models import DocumentChunk  # noqa: E402
from morphik_core_b200.store import B200MultiVectorStore  # noqa: E402
from oracle import maxsim_oracle as orc  # noqa: E402


def run(coro):
    return asyncio.run(coro)


def get_sample_embeddings(num_vectors=3, dim=128, seed=0):  # test_multivector.py:23-34 (torch.rand * 2 - 1)
    g = torch.Generator().manual_seed(seed)
    return torch.rand((num_vectors, dim), generator=g) * 2 - 1


@pytest.fixture(params=["binary", "bf16", "int8", "fp8"])
def vector_store(request):
    store = B200MultiVectorStore(mode=request.param)
    yield store
    store.close()


def test_store_embeddings_empty(vector_store):  # test_multivector.py:206-211
    ok, ids, metrics = run(vector_store.store_embeddings([]))
    assert ok is True and ids == [] and "vector_store_rows" in metrics


def test_store_and_query_self_match(vector_store):  # test_multivector.py:152-181
    chunks = [DocumentChunk(document_id=f"doc_{i}", content=f"Test content {i}", embedding=get_sample_embeddings(seed=i),
                            chunk_number=i, metadata={"index": i}) for i in range(6)]
    ok, ids, _ = run(vector_store.store_embeddings(chunks))
    assert ok and ids == [f"doc_{i}-{i}" for i in range(6)]
    results = run(vector_store.query_similar(get_sample_embeddings(seed=2), k=3))
    assert len(results) == 3 and results[0].document_id == "doc_2"
    assert all(results[i].score >= results[i + 1].score for i in range(2))
    assert all(r.embedding == [] for r in results)
    if vector_store.mode == "binary":
        assert results[0].score == 3.0  # a 3-vector self query scores exactly T under SQL max_sim


def test_query_with_doc_ids(vector_store):  # test_multivector.py:184-202
    chunks = [DocumentChunk(document_id=f"doc_{i % 3}", content=f"c{i}", embedding=get_sample_embeddings(seed=10 + i),
                            chunk_number=i, metadata={}) for i in range(9)]
    run(vector_store.store_embeddings(chunks))
    results = run(vector_store.query_similar(get_sample_embeddings(seed=99), k=9, doc_ids=["doc_0", "doc_2"]))
    assert len(results) == 6 and {r.document_id for r in results} == {"doc_0", "doc_2"}
    assert run(vector_store.query_similar(get_sample_embeddings(seed=99), k=9, doc_ids=["unknown"])) == []


def test_multi_vector_similarity(vector_store):  # test_multivector.py:214-256
    e1 = np.ones((3, 128)); e1[:, 64:] = -1
    e2 = -np.ones((3, 128)); e2[:, 64:] = 1
    run(vector_store.store_embeddings([
        DocumentChunk(document_id="similarity_test_1", content="Similarity test content 1", embedding=e1, chunk_number=1),
        DocumentChunk(document_id="similarity_test_2", content="Similarity test content 2", embedding=e2, chunk_number=2)]))
    q = np.ones(128); q[64:] = -1
    results = run(vector_store.query_similar(np.array([q]), k=2))
    assert [r.document_id for r in results] == ["similarity_test_1", "similarity_test_2"]
    if vector_store.mode == "binary":
        assert [r.score for r in results] == [1.0, 0.0]
    elif vector_store.mode == "bf16":
        assert [r.score for r in results] == [128.0, -128.0]


def test_store_and_retrieve_metadata(vector_store):  # test_multivector.py:259-294
    meta = {"page": 3, "is_image": True, "nested": {"a": [1, 2]}}
    run(vector_store.store_embeddings([DocumentChunk(document_id="metadata_test", content="payload", chunk_number=7,
                                                     embedding=get_sample_embeddings(seed=5), metadata=meta)]))
    res = run(vector_store.query_similar(get_sample_embeddings(seed=5), k=1))
    assert res[0].metadata == meta and res[0].content == "payload" and res[0].chunk_number == 7
    got = run(vector_store.get_chunks_by_id([("metadata_test", 7), ("metadata_test", 8)]))
    assert len(got) == 1 and got[0].metadata == meta and got[0].score == 0.0


def test_delete_and_compaction_keep_scores_consistent(vector_store):
    rng = np.random.default_rng(0)
    pages = {f"d{i}": [rng.standard_normal((int(rng.integers(1, 90)), 128)).astype(np.float32) for _ in range(3)] for i in range(8)}
    for d, ps in pages.items():
        run(vector_store.store_embeddings([DocumentChunk(document_id=d, content=d, embedding=p, chunk_number=j)
                                           for j, p in enumerate(ps)], app_id="app"))
    q = rng.standard_normal((20, 128)).astype(np.float32)
    before = run(vector_store.query_similar(q, k=24))
    for d in ("d1", "d2", "d5"):  # > 30 % dead -> compaction on the device
        assert run(vector_store.delete_chunks_by_document_id(d)) is True
    after = run(vector_store.query_similar(q, k=24))
    expect = [(r.document_id, r.chunk_number, r.score) for r in before if r.document_id not in ("d1", "d2", "d5")]
    assert [(r.document_id, r.chunk_number, r.score) for r in after] == expect
    assert len(vector_store.catalog) == 15


def test_binary_scores_equal_sql_max_sim_oracle():
    store = B200MultiVectorStore(mode="binary")
    rng = np.random.default_rng(3)
    lens = [int(x) for x in rng.integers(1, 120, size=40)]
    pages = [rng.standard_normal((n, 128)).astype(np.float32) for n in lens]
    run(store.store_embeddings([DocumentChunk(document_id=f"doc{i // 4}", content="", embedding=p, chunk_number=i % 4)
                                for i, p in enumerate(pages)]))
    q = rng.standard_normal((32, 128)).astype(np.float32)
    res = run(store.query_similar(q, k=40))
    want, _ = orc.binary_maxsim_c(orc.sign_pack_c(q), orc.sign_pack_c(np.concatenate(pages)), orc.page_offsets(lens))
    ws, wi = orc.topk_np(want, 40)
    assert [(r.document_id, r.chunk_number) for r in res] == [(f"doc{i // 4}", i % 4) for i in wi]
    assert [r.score for r in res] == ws.tolist()
    store.close()


@pytest.mark.parametrize("mode", ["bf16", "binary", "int8"])
def test_device_tensor_ingest_fast_path_matches_host_ingest(mode):
    """CUDA bf16 tensors (what the embedding model holds) go straight into the pack kernel; the stored corpus
    and every score equal the reference-style host float32 ingest of the same bf16 values; min_score cuts the list."""
    g = torch.Generator().manual_seed(5)
    pages = [torch.nn.functional.normalize(torch.randn((n, 128), generator=g), dim=1).bfloat16() for n in (40, 1, 97, 1030, 33)]
    q = torch.nn.functional.normalize(torch.randn((32, 128), generator=g), dim=1).bfloat16()
    dev, host = B200MultiVectorStore(mode=mode), B200MultiVectorStore(mode=mode)
    mk = lambda conv: [DocumentChunk(document_id=f"d{i}", content="", embedding=conv(p), chunk_number=0)  # noqa: E731
                       for i, p in enumerate(pages)]
    ok, ids, _ = run(dev.store_embeddings(mk(lambda p: p.cuda())))
    assert ok and len(ids) == 5
    run(host.store_embeddings(mk(lambda p: p.float().numpy())))
    a, b = run(dev.query_similar(q.cuda(), k=5)), run(host.query_similar(q.float().numpy(), k=5))
    assert [(r.document_id, r.score) for r in a] == [(r.document_id, r.score) for r in b]
    cut = (a[1].score + a[2].score) / 2
    c = run(dev.query_similar_batch([q], k=5, min_score=cut))[0]
    assert [(r.document_id, r.score) for r in c] == [(r.document_id, r.score) for r in a[:2]]
    dev.close(); host.close()


@pytest.mark.parametrize("mode", ["bf16", "binary", "int8"])
def test_per_query_masks_equal_individual_searches(mode):
    """b200ms_search_host_masked: one pass, one allow-mask per query == the same queries searched one by one."""
    from morphik_core_b200.catalog import PageCatalog
    from morphik_core_b200.index import MaxSimIndex

    rng = np.random.default_rng(31)
    idx = MaxSimIndex(dtype=mode)
    pages = [rng.standard_normal((int(rng.integers(5, 90)), 128)).astype(np.float32) for _ in range(300)]
    idx.add_pages(pages)
    queries = [rng.standard_normal((int(rng.integers(4, 40)), 128)).astype(np.float32) for _ in range(11)]
    masks = []
    for i in range(11):
        if i % 4 == 0:
            masks.append(None)
        else:
            m = rng.random(300) < (0.02 if i % 4 == 1 else 0.5)
            masks.append(PageCatalog.mask_words(m))
    masks[5] = masks[3]  # shared filter: uploaded once
    ts, ti, tc = idx.search_host_masked(queries, 7, masks)
    for i, (q, m) in enumerate(zip(queries, masks)):
        s1, i1, c1 = idx.search_host([q], 7, allow_mask=m)
        assert tc[i] == c1[0] and np.array_equal(ti[i], i1[0]) and np.array_equal(ts[i], s1[0]), i
    with pytest.raises(ValueError):
        idx.search_host_masked(queries, 7, masks[:3])
    # the device form (what every rank of the sharded store calls): mask matrix + row index per query
    rows, index = [], []
    for m in masks:
        if m is None:
            index.append(-1)
        else:
            index.append(next((j for j, r in enumerate(rows) if np.array_equal(r, m)), len(rows)))
            if index[-1] == len(rows):
                rows.append(m)
    mm = torch.from_numpy(np.stack(rows).view(np.int32)).cuda()
    mi = torch.tensor(index, dtype=torch.int32).cuda()
    q_dev = torch.from_numpy(np.concatenate(queries)).cuda()
    ds, di, dc = idx.search_device(q_dev, [len(q) for q in queries], 7, allow_mask_dev=mm, mask_index_dev=mi)
    torch.cuda.synchronize()
    assert np.array_equal(ds.cpu().numpy(), ts) and np.array_equal(di.cpu().numpy(), ti) and np.array_equal(dc.cpu().numpy(), tc)


def test_coalesced_concurrent_queries_use_per_query_masks():
    """The store coalesces concurrent query_similar calls with different doc_ids into one masked GPU pass."""
    store = B200MultiVectorStore(mode="bf16")
    rng = np.random.default_rng(8)
    pages = [rng.standard_normal((int(rng.integers(10, 80)), 128)).astype(np.float32) for _ in range(120)]
    run(store.store_embeddings([DocumentChunk(document_id=f"d{i % 30}", content="", embedding=p, chunk_number=i // 30)
                                for i, p in enumerate(pages)]))
    reqs = [dict(query_embedding=pages[i][:12], k=2 + i % 4, doc_ids=None if i % 5 == 0 else [f"d{j}" for j in range(i % 30 + 1)])
            for i in range(20)]
    lone = [run(store.query_similar(**r)) for r in reqs]

    async def many():
        return await asyncio.gather(*[store.query_similar(**r) for r in reqs])

    together = run(many())
    key = lambda res: [(c.document_id, c.chunk_number, c.score) for c in res]  # noqa: E731
    assert [key(r) for r in together] == [key(r) for r in lone]
    assert store.last_coalesced_batch == 20
    store.close()


def test_save_load_round_trip(tmp_path):
    store = B200MultiVectorStore(mode="bf16")
    rng = np.random.default_rng(8)
    chunks = [DocumentChunk(document_id=f"d{i // 3}", content=f"content {i}", chunk_number=i % 3, metadata={"i": i},
                            embedding=rng.standard_normal((int(rng.integers(1, 80)), 128)).astype(np.float32)) for i in range(30)]
    run(store.store_embeddings(chunks, app_id="app"))
    run(store.delete_chunks_by_document_id("d4"))  # leaves tombstones (10 % dead: no automatic compaction)
    q = rng.standard_normal((32, 128)).astype(np.float32)
    before = run(store.query_similar(q, k=27))
    store.save(str(tmp_path))
    loaded = B200MultiVectorStore.load(str(tmp_path))
    after = run(loaded.query_similar(q, k=27))
    assert [(r.document_id, r.chunk_number, r.content, r.metadata, r.score) for r in after] == \
           [(r.document_id, r.chunk_number, r.content, r.metadata, r.score) for r in before]
    assert len(loaded.catalog) == 27 and run(loaded.get_chunks_by_id([("d4", 0)])) == []
    store.close(); loaded.close()


def test_two_stage_store_matches_exhaustive_when_candidates_cover_corpus():
    rng = np.random.default_rng(12)
    chunks = [DocumentChunk(document_id=f"d{i % 7}", content=str(i), chunk_number=i, metadata={},
                            embedding=rng.standard_normal((int(rng.integers(5, 60)), 128)).astype(np.float32)) for i in range(60)]
    ex = B200MultiVectorStore(mode="bf16")
    two = B200MultiVectorStore(mode="bf16", fde_candidates=64)  # >= corpus size: the rerank sees every page
    run(ex.store_embeddings(chunks)); run(two.store_embeddings(chunks))
    q = rng.standard_normal((20, 128)).astype(np.float32)
    for doc_ids in (None, ["d1", "d3"]):
        a = run(ex.query_similar(q, k=8, doc_ids=doc_ids))
        b = run(two.query_similar(q, k=8, doc_ids=doc_ids))
        assert [(r.document_id, r.chunk_number) for r in a] == [(r.document_id, r.chunk_number) for r in b]
        np.testing.assert_allclose([r.score for r in a], [r.score for r in b], rtol=1e-6)
    run(two.delete_chunks_by_document_id("d1")); run(two.delete_chunks_by_document_id("d2")); run(two.delete_chunks_by_document_id("d3"))
    run(ex.delete_chunks_by_document_id("d1")); run(ex.delete_chunks_by_document_id("d2")); run(ex.delete_chunks_by_document_id("d3"))
    a, b = run(ex.query_similar(q, k=8)), run(two.query_similar(q, k=8))
    assert [(r.document_id, r.chunk_number) for r in a] == [(r.document_id, r.chunk_number) for r in b]
    ex.close(); two.close()


def test_concurrent_queries_are_serialised_and_consistent():
    """Several query_similar coroutines in flight at once (the API server's situation, SURVEY 8b 'Threading'): the store's
    lock serialises the GPU calls and every coroutine gets the answer of its own query."""
    store = B200MultiVectorStore(mode="bf16")
    rng = np.random.default_rng(21)
    pages = [rng.standard_normal((int(rng.integers(10, 120)), 128)).astype(np.float32) for _ in range(200)]
    run(store.store_embeddings([DocumentChunk(document_id=f"d{i}", content="", embedding=p, chunk_number=0) for i, p in enumerate(pages)]))
    queries = [pages[i][:16] + 0.01 * rng.standard_normal((16, 128)).astype(np.float32) for i in range(24)]

    async def many():
        return await asyncio.gather(*[store.query_similar(q, k=3) for q in queries])

    results = run(many())
    assert [r[0].document_id for r in results] == [f"d{i}" for i in range(24)]
    sequential = [run(store.query_similar(q, k=3)) for q in queries]
    assert [[(c.document_id, c.score) for c in r] for r in results] == [[(c.document_id, c.score) for c in r] for r in sequential]
    store.close()


def _page_chunks(rng, doc, n_pages, lo=3, hi=90):
    out = []
    for j in range(n_pages):
        x = rng.standard_normal((int(rng.integers(lo, hi)), 128)).astype(np.float32)
        x = torch.from_numpy(x / np.linalg.norm(x, axis=1, keepdims=True)).bfloat16().float().numpy()  # bf16-valued like ColPali output
        out.append(DocumentChunk(document_id=doc, content=f"{doc}/{j}", embedding=x, chunk_number=j, metadata={"j": j}))
    return out


@pytest.mark.parametrize("mode,fde_candidates", [("bf16", None), ("binary", None), ("bf16", 64), ("int8", 64)])
def test_journaled_store_appends_and_replays(tmp_path, mode, fde_candidates):
    """every store_embeddings call becomes an append-only segment, every delete a tombstone line; reopening the
    directory replays them in order (incl. delete-then-reinsert of the same document) without rewriting anything."""
    import os

    rng = np.random.default_rng(17)
    d = str(tmp_path / "store")
    kw = dict(mode=mode, fde_candidates=fde_candidates, compact_dead_fraction=0.9)
    s = B200MultiVectorStore.open(d, **kw)
    docs = {f"doc{i}": _page_chunks(rng, f"doc{i}", int(rng.integers(1, 5))) for i in range(10)}
    for name, chunks in docs.items():
        run(s.store_embeddings(chunks, app_id="app"))
    sizes = {f: os.path.getsize(os.path.join(d, f)) for f in os.listdir(d) if f.endswith(".b2ms")}
    assert len(sizes) == 10
    assert run(s.delete_chunks_by_document_id("doc3")) and run(s.delete_chunks_by_document_id("doc7"))
    docs["doc3"] = _page_chunks(rng, "doc3", 2)  # re-insert after the tombstone: must survive the replay
    run(s.store_embeddings(docs["doc3"], app_id="app"))
    docs.pop("doc7")
    assert {f: os.path.getsize(os.path.join(d, f)) for f in sizes} == sizes  # nothing was rewritten
    q = docs["doc5"][0].embedding
    want = [(r.document_id, r.chunk_number, r.score) for r in run(s.query_similar(q, k=8, app_id="app"))]
    assert want[0][:2] == ("doc5", 0) and all(x[0] != "doc7" for x in want)
    s.close()
    s2 = B200MultiVectorStore.open(d, **kw)
    got = [(r.document_id, r.chunk_number, r.score) for r in run(s2.query_similar(q, k=8, app_id="app"))]
    assert got == want
    assert run(s2.get_chunks_by_id([("doc3", 1), ("doc7", 0)]))[0].content == "doc3/1"
    assert len(run(s2.get_chunks_by_id([("doc7", 0)]))) == 0
    # checkpoint: one base segment, empty tombstone log, same answers
    s2.save(d)
    assert sorted(f for f in os.listdir(d) if f.endswith(".b2ms")) == ["corpus.b2ms", "seg-000001.b2ms"]
    run(s2.store_embeddings(_page_chunks(rng, "late", 1), app_id="app"))
    s2.close()
    s3 = B200MultiVectorStore.open(d, **kw)
    assert [(r.document_id, r.chunk_number, r.score) for r in run(s3.query_similar(q, k=8, app_id="app", doc_ids=list(docs)))] == want
    assert run(s3.get_chunks_by_id([("late", 0)]))[0].content == "late/0"
    s3.close()


def test_load_with_fde_candidates_rebuilds_two_stage(tmp_path):
    rng = np.random.default_rng(23)
    s = B200MultiVectorStore(mode="bf16")
    for i in range(30):
        run(s.store_embeddings(_page_chunks(rng, f"d{i}", 3, lo=32, hi=120)))
    q = s.catalog.records[40]
    qe = _page_chunks(np.random.default_rng(1), "q", 1)[0].embedding[:20]
    want = [(r.document_id, r.chunk_number) for r in run(s.query_similar(qe, k=5))]
    s.save(str(tmp_path))
    two = B200MultiVectorStore.load(str(tmp_path), fde_candidates=90)  # 90 candidates = the whole corpus: exhaustive answer
    assert two._two_stage is not None and two._two_stage.n_pages == 90
    assert [(r.document_id, r.chunk_number) for r in run(two.query_similar(qe, k=5))] == want
    assert {"encode_query_ms", "ns_query_ms", "load_multivectors_ms", "rerank_scoring_ms", "build_chunks_ms", "total_ms"} <= set(
        two.last_query_timing) or two.coalesce_queries
    res = run(two.query_similar_batch([qe], k=5))
    assert {"encode_query_ms", "ns_query_ms", "rerank_scoring_ms", "build_chunks_ms", "total_ms"} <= set(two.last_query_timing)
    assert [(r.document_id, r.chunk_number) for r in res[0]] == want
    s.close(); two.close()


def test_store_zero_pad_compat_option():
    """B200MultiVectorStore(zero_pad_compat=128) scores like the reference's padded batches."""
    rng = np.random.default_rng(29)
    lens = [5, 64, 1, 33, 64, 17, 40, 2]
    pages = [np.abs(rng.standard_normal((n, 128))).astype(np.float32) for n in lens]
    pages = [p / np.linalg.norm(p, axis=1, keepdims=True) for p in pages]
    q = -np.abs(rng.standard_normal((8, 128))).astype(np.float32)
    s = B200MultiVectorStore(mode="bf16", zero_pad_compat=128)
    run(s.store_embeddings([DocumentChunk(document_id=f"d{i}", content="", embedding=p, chunk_number=0) for i, p in enumerate(pages)]))
    res = run(s.query_similar(q, k=8))
    want = orc.float_maxsim_c(orc.bf16_round_np(q), orc.bf16_round_np(np.concatenate(pages)), orc.page_offsets(lens),
                              zero_pad_compat=True, batch=128)
    got = {r.document_id: r.score for r in res}
    for i in range(8):
        assert abs(got[f"d{i}"] - want[i]) <= 3e-5 * max(1.0, abs(want[i]))
    assert all(got[f"d{i}"] == 0.0 for i in range(8) if lens[i] < 64) and all(got[f"d{i}"] < 0 for i in (1, 4))
    s.close()
