"""CPU tests of host-side pieces added with the "next" rows: FDE configuration/matrices (fde.py), shard-file header
handling (shardfile.py) -- no GPU needed; the device parts are covered by tests/test_gpu_fde.py and
tests/test_gpu_shardfile_reranker.py."""
import json

import numpy as np
import pytest

from morphik_core_b200 import fde, shardfile
from oracle import maxsim_oracle as orc


def test_fde_config_mirrors_reference_call_site():
    # core/vector_store/fast_multivector_store.py:325-331
    cfg = fde.FixedDimensionalEncodingConfig(dimension=128, num_repetitions=20, num_simhash_projections=5,
                                             projection_dimension=16, projection_type="AMS_SKETCH")
    assert cfg.fde_dimension == 10240 and cfg.num_partitions == 32 and abs(cfg.scale - 0.25) < 1e-12
    a = fde.fde_matrices(cfg)
    b = fde.fde_matrices(fde.FixedDimensionalEncodingConfig())
    assert all(np.array_equal(x, y) for x, y in zip(a, b))  # deterministic in the seed
    c = fde.fde_matrices(fde.FixedDimensionalEncodingConfig(seed=2))
    assert not np.array_equal(a[0], c[0])
    assert np.array_equal(a[0][1], c[0][0])  # repetition r of seed s uses stream s + r
    for bad in (dict(dimension=64), dict(projection_type="DEFAULT_IDENTITY"), dict(final_projection_dimension=100)):
        with pytest.raises(ValueError):
            fde.fde_matrices(fde.FixedDimensionalEncodingConfig(**bad))


def test_fde_oracle_semantics_sum_vs_average_and_empty_partitions():
    cfg = fde.FixedDimensionalEncodingConfig(num_repetitions=3)
    sh, ai, sg = fde.fde_matrices(cfg)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((10, 128)).astype(np.float32)
    q = orc.fde_encode_c_proj(x, sh, ai, sg, cfg.scale, False, 16).reshape(3, 32, 16)
    d = orc.fde_encode_c_proj(x, sh, ai, sg, cfg.scale, True, 16).reshape(3, 32, 16)
    # 10 points fall into at most 10 of the 32 partitions: the rest stay exactly zero (fill_empty_partitions=False)
    occupied = np.abs(q).sum(axis=2) > 0
    assert (occupied.sum(axis=1) <= 10).all() and np.array_equal(occupied, np.abs(d).sum(axis=2) > 0)
    # duplicating every point doubles the query FDE (SUM) and leaves the document FDE (AVERAGE) unchanged
    x2 = np.concatenate([x, x])
    q2 = orc.fde_encode_c_proj(x2, sh, ai, sg, cfg.scale, False, 16).reshape(3, 32, 16)
    d2 = orc.fde_encode_c_proj(x2, sh, ai, sg, cfg.scale, True, 16).reshape(3, 32, 16)
    np.testing.assert_allclose(q2, 2 * q, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(d2, d, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(orc.fde_encode_np(x, sh, ai, sg, cfg.scale, True, 16), d.reshape(-1), rtol=1e-5, atol=1e-6)


def test_shard_header_round_trip_and_rejection(tmp_path):
    lens = np.array([3, 0, 64, 1030], dtype=np.int32)
    header = {"magic": shardfile.MAGIC, "version": 1, "dtype": "bf16", "n_pages": 4, "n_rows_padded": 32 + 0 + 64 + 1056,
              "row_bytes": 256, "i8_scale": 127.0}
    path = tmp_path / "x.b2ms"
    with open(path, "wb") as f:
        f.write(json.dumps(header).encode().ljust(shardfile.HEADER_BYTES, b"\0"))
        f.write(lens.tobytes().ljust(4096, b"\0"))
        f.write(b"\0" * (header["n_rows_padded"] * 256))
    h, l, off = shardfile.read_header(str(path))
    assert h == header and l.tolist() == lens.tolist() and off == 8192
    bad = tmp_path / "bad.b2ms"
    bad.write_bytes(json.dumps({"magic": "nope"}).encode().ljust(4096, b"\0"))
    with pytest.raises(ValueError, match="not a B2MSHARD"):
        shardfile.read_header(str(bad))
    pages = shardfile.pages_from_bit_rows([[bytes([0b10100101] + [0] * 15)], []])
    assert pages[0].shape == (1, 128) and pages[0][0, :8].tolist() == [1, -1, 1, -1, -1, 1, -1, 1] and pages[1].shape == (0, 128)
    assert np.array_equal(orc.sign_pack_c(pages[0])[0], np.array([0b10100101] + [0] * 15, dtype=np.uint8))


def test_store_journal_log_survives_a_torn_last_line(tmp_path):
    """journal.log is append-only; a crash mid-append leaves a torn last line that replay must ignore (everything before it
    was fsynced); sequence numbers continue after the last complete segment."""
    d = tmp_path / "store"
    j = shardfile.StoreJournal(str(d))
    assert j.read_ops() == [] and j.seq == 0
    j._append({"op": "segment", "seq": 1, "pages": 3})
    j.log_delete("doc-a")
    j._append({"op": "segment", "seq": 2, "pages": 1})
    with open(d / "journal.log", "a") as f:
        f.write('{"op": "segment", "seq": 3, "pa')  # torn
    ops = shardfile.StoreJournal(str(d)).read_ops()
    assert [o["op"] for o in ops] == ["segment", "delete", "segment"] and ops[1]["document_id"] == "doc-a"
    assert shardfile.StoreJournal(str(d)).seq == 2
    assert j.segment_files(2)[0].endswith("seg-000002.b2ms")
    (d / "seg-000001.b2ms").write_bytes(b"x")
    j.reset()
    assert sorted(p.name for p in d.iterdir()) == [] and j.seq == 0
