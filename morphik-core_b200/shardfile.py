"""Packed shard files -- the durable form of a device corpus.

The reference keeps multivectors as one object per page: ``numpy.save`` of float32 ``[P,128]`` under
``multivector/{document_id}/{chunk_number}.npy`` (core/vector_store/fast_multivector_store.py:673-707, read back one by one
at :713-774 through a disk LRU cache), or as ``BIT(128)[]`` rows in Postgres (multi_vector_store.py:240-251).  Rebuilding a
GPU corpus from those costs one object fetch / row decode per page.  A shard file is the packed HBM layout itself
(DESIGN.md section 3), so loading is a sequence of large sequential reads + H2D copies and nothing else:

    offset 0     : 4096-byte header  = magic "B2MSHARD", version, dtype, n_pages, n_rows_padded, row_bytes, i8_scale (JSON, NUL padded)
    offset 4096  : page_lens  int32[n_pages]   (true lengths), zero padded to a 4096-byte boundary
    then         : rows       n_rows_padded * row_bytes bytes, exactly as they sit in HBM

Importers for the reference's durable forms: ``import_npy_pages`` (float32 .npy files) and ``pages_from_bit_rows``
(packed 16-byte sign-bit rows as pgvector's ``Bit`` stores them, multi_vector_store.py:339-345).
"""
from __future__ import annotations

import json
import os
from typing import Iterable, List, Sequence, Tuple

import numpy as np
import torch

from . import _native as nat
from .index import MaxSimIndex, _aligned_bytes

MAGIC = "B2MSHARD"
HEADER_BYTES = 4096
CHUNK_BYTES = 256 << 20  # staging granularity for file <-> pinned <-> device copies


def _round_up(x: int, a: int) -> int:
    return (x + a - 1) // a * a


def save_index(index: MaxSimIndex, path: str) -> int:
    """Write the packed corpus of `index` to `path`; returns bytes written."""
    lens = np.asarray(index.page_lens, dtype=np.int32)
    rows = index.packed_rows().reshape(-1)  # uint8 device view
    header = {"magic": MAGIC, "version": 1, "dtype": index.dtype_name, "n_pages": int(len(lens)),
              "n_rows_padded": int(index.n_rows_padded), "row_bytes": int(index.row_bytes), "i8_scale": float(index.i8_scale)}
    blob = json.dumps(header).encode()
    assert len(blob) < HEADER_BYTES
    lens_bytes = _round_up(lens.nbytes, 4096)
    with open(path, "wb") as f:
        f.write(blob.ljust(HEADER_BYTES, b"\0"))
        f.write(lens.tobytes().ljust(lens_bytes, b"\0"))
        total = rows.numel()
        stage = torch.empty(min(CHUNK_BYTES, max(total, 1)), dtype=torch.uint8).pin_memory()
        for o in range(0, total, CHUNK_BYTES):
            n = min(CHUNK_BYTES, total - o)
            stage[:n].copy_(rows[o:o + n])  # D2H into pinned memory (synchronous for a pinned destination)
            f.write(stage[:n].numpy().tobytes() if n < stage.numel() else stage.numpy().data)
    return HEADER_BYTES + lens_bytes + int(rows.numel())


def read_header(path: str) -> Tuple[dict, np.ndarray, int]:
    with open(path, "rb") as f:
        raw = f.read(HEADER_BYTES)
        header = json.loads(raw.rstrip(b"\0").decode())
        if header.get("magic") != MAGIC or header.get("version") != 1:
            raise ValueError(f"{path}: not a B2MSHARD v1 file")
        lens = np.frombuffer(f.read(4 * header["n_pages"]), dtype=np.int32).copy()
    payload_off = HEADER_BYTES + _round_up(4 * header["n_pages"], 4096)
    return header, lens, payload_off


def load_index(path: str, device: int = 0) -> MaxSimIndex:
    """Read a shard file straight into a 1024-byte-aligned device buffer and adopt it (no repacking)."""
    header, lens, payload_off = read_header(path)
    index = MaxSimIndex(device=device, dtype=header["dtype"], i8_scale=header["i8_scale"])
    total = header["n_rows_padded"] * header["row_bytes"]
    if int(nat.lib.b200ms_padded_rows(nat.i32_array(lens.tolist()), len(lens))) != header["n_rows_padded"]:
        raise ValueError(f"{path}: page lengths do not match the row count")
    if os.path.getsize(path) < payload_off + total:
        raise ValueError(f"{path}: truncated payload")
    buf = _aligned_bytes(total, index.device)
    stages = [torch.empty(min(CHUNK_BYTES, max(total, 1)), dtype=torch.uint8).pin_memory() for _ in range(2)]
    events = [torch.cuda.Event(), torch.cuda.Event()]
    with open(path, "rb") as f, torch.cuda.device(index.device):
        f.seek(payload_off)
        for i, o in enumerate(range(0, total, CHUNK_BYTES)):  # double-buffered: read chunk i+1 while chunk i is in flight
            n = min(CHUNK_BYTES, total - o)
            st = stages[i & 1]
            if i >= 2:
                events[i & 1].synchronize()
            got = f.readinto(memoryview(st.numpy())[:n])
            if got != n:
                raise ValueError(f"{path}: short read")
            buf[o:o + n].copy_(st[:n], non_blocking=True)
            events[i & 1].record()
        torch.cuda.current_stream().synchronize()
    index.adopt_packed(buf, lens.tolist())
    return index


def import_npy_pages(index: MaxSimIndex, paths: Sequence[str], batch: int = 256) -> int:
    """Append pages stored the reference's way (one float32 [P,128] ``.npy`` per page, fast_multivector_store.py:673-707)."""
    n = 0
    for i in range(0, len(paths), batch):
        pages = [np.load(p).astype(np.float32, copy=False).reshape(-1, nat.DIM) for p in paths[i:i + batch]]
        index.add_pages(pages)
        n += len(pages)
    return n


def pages_from_bit_rows(rows_per_page: Iterable[Sequence[bytes]]) -> List[np.ndarray]:
    """Postgres ``BIT(128)[]`` rows (each element the 16 packed bytes pgvector's ``Bit`` holds, MSB first) -> float32 +-1
    pages whose sign bits are exactly those bits -- feed to a ``dtype="binary"`` index (sign(x) > 0 recovers every bit; this
    is the migration path of scripts/migrate_postgres_to_turbopuffer.py:233-273 in the other direction)."""
    pages = []
    for rows in rows_per_page:
        arr = np.frombuffer(b"".join(rows), dtype=np.uint8).reshape(-1, 16) if len(rows) else np.zeros((0, 16), np.uint8)
        bits = np.unpackbits(arr, axis=1, bitorder="big").astype(np.float32)
        pages.append(bits * 2.0 - 1.0)
    return pages
