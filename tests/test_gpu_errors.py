"""Error behaviour and degenerate inputs through the C-ABI on a real GPU: calls fail loudly with a message (negative code ->
NativeError / ValueError), never crash, and empty corpora / empty queries give defined results."""
import asyncio
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("CUDA device required", allow_module_level=True)

from morphik_core_b200 import _native as nat  # noqa: E402
from morphik_core_b200.index import MaxSimIndex  # noqa: E402
from morphik_core_b200.models import DocumentChunk  # noqa: E402
from morphik_core_b200.store import B200MultiVectorStore  # noqa: E402


def test_call_order_and_argument_errors():
    h = nat.Handle(0)
    q = torch.zeros((128, 128), dtype=torch.bfloat16, device="cuda")
    s = torch.zeros((4, 32), dtype=torch.float32, device="cuda")
    rc = nat.lib.b200ms_score(h.ptr, ctypes.c_void_p(q.data_ptr()), 1, None, None, 0, ctypes.c_void_p(s.data_ptr()), 32, None)
    assert rc == -3 and b"no corpus attached" in nat.lib.b200ms_last_error(h.ptr)  # B200MS_ESTATE
    assert nat.lib.b200ms_set_option(h.ptr, b"no_such_option", 1) == -1
    assert nat.lib.b200ms_set_corpus(h.ptr, ctypes.c_void_p(q.data_ptr() + 16), nat.BF16, nat.i32_array([32]), 1) == -1  # misaligned
    assert b"1024-byte aligned" in nat.lib.b200ms_last_error(h.ptr)
    assert nat.lib.b200ms_fde_encode(h.ptr, None, nat.F32, None, 0, 0, None, None) == -3  # FDE not configured
    h.close()
    with pytest.raises(nat.NativeError, match="bad device"):
        nat.Handle(torch.cuda.device_count() + 3)


def test_index_input_validation_and_limits():
    idx = MaxSimIndex(dtype="bf16")
    with pytest.raises(ValueError, match=r"\[P,128\]"):
        idx.add_pages([np.zeros((4, 64), np.float32)])
    idx.add_pages([np.random.default_rng(0).standard_normal((40, 128)).astype(np.float32)])
    with pytest.raises(ValueError, match=r"\[T,128\]"):
        idx.search_host([np.zeros((3, 127), np.float32)], k=1)
    with pytest.raises(nat.NativeError, match="k <= 4096"):
        idx.search_host([np.zeros((3, 128), np.float32)], k=5000)
    with pytest.raises(ValueError):
        MaxSimIndex(dtype="fp64")


def test_empty_corpus_and_empty_query_are_defined():
    idx = MaxSimIndex(dtype="bf16")
    ts, ti, tc = idx.search_host([np.ones((5, 128), np.float32)], k=3)  # no pages at all
    assert tc[0] == 0 and np.all(ti == -1) and np.all(np.isinf(ts))
    idx.add_pages([np.ones((4, 128), np.float32), np.zeros((0, 128), np.float32), -np.ones((2, 128), np.float32)])
    ts, ti, tc = idx.search_host([np.zeros((0, 128), np.float32)], k=3)  # a query without tokens scores 0 everywhere
    assert tc[0] == 3 and ti[0].tolist() == [0, 1, 2] and ts[0].tolist() == [0.0, 0.0, 0.0]
    ts, ti, tc = idx.search_host([np.ones((2, 128), np.float32)], k=3)
    assert ti[0].tolist() == [0, 1, 2] and ts[0].tolist() == [256.0, 0.0, -256.0]  # the empty page scores 0 (COALESCE)
    store = B200MultiVectorStore(mode="binary")
    assert asyncio.run(store.query_similar(np.ones((3, 128)), k=5)) == []
    assert asyncio.run(store.query_similar(np.ones((3, 128)), k=0)) == []
    with pytest.raises(ValueError):
        asyncio.run(store.store_embeddings([DocumentChunk(document_id="d", content="", embedding=np.zeros((2, 64)), chunk_number=0)]))
    store.close()
