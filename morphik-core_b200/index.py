"""MaxSimIndex -- a device-resident, packed multi-vector corpus plus the search calls over it.

This is the piece that replaces the reference's exhaustive scorers:
  * Postgres evaluating SQL ``max_sim`` over every ``multi_vector_embeddings`` row
    (core/vector_store/multi_vector_store.py:287-311, 746-763)            -> dtype "binary"
  * ``ColQwen2_5_Processor.score_multi_vector`` + ``torch.topk``
    (core/vector_store/fast_multivector_store.py:553-557)                -> dtype "bf16"
  * (new, BASELINE config 3) integer MaxSim over int8 patches            -> dtype "int8"

PyTorch is used only to own device memory and to name the current stream; every computation goes through the
C-ABI of libb200ms.so (``_native``).  Nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _native as nat

ArrayLike = Union[np.ndarray, torch.Tensor]


def _vp(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _aligned_bytes(nbytes: int, device: torch.device, align: int = 1024) -> torch.Tensor:
    """uint8 device tensor whose data_ptr is `align`-byte aligned (set_corpus / TMA want 1024)."""
    raw = torch.empty(max(int(nbytes), 1) + align, dtype=torch.uint8, device=device)
    off = (-raw.data_ptr()) % align
    return raw[off:off + max(int(nbytes), 1)]


class MaxSimIndex:
    """Packed corpus of pages (each a [P_i, 128] matrix of patch embeddings) on one GPU.

    Layout in HBM: rows of ``row_bytes`` (bf16 256 B / int8 128 B / sign bits 16 B), every page padded to a
    multiple of 32 rows by repeating its last row, pages back to back in insertion order.  ``page id`` = insertion index.
    """

    def __init__(self, device: int = 0, dtype: str = "bf16", i8_scale: float = 127.0, capacity_rows: int = 0):
        if dtype not in nat.DTYPE_NAMES or nat.DTYPE_NAMES[dtype] == nat.F32:
            raise ValueError(f"dtype must be one of bf16/int8/fp8/binary, got {dtype!r}")
        if not torch.cuda.is_available():
            raise nat.NativeError("MaxSimIndex needs a CUDA device (B200); there is no CPU fallback")
        self.device = torch.device("cuda", int(device))
        self.dtype = nat.DTYPE_NAMES[dtype]
        self.dtype_name = dtype
        self.row_bytes = nat.ROW_BYTES[self.dtype]
        # corpus AND query quantisation scale: int8 rint(x*127) for unit-norm rows; fp8 e4m3(x*64) -- a power of two keeps the
        # scaling exact and puts |x| in [2^-12, 7] into e4m3's normal range
        self.i8_scale = 64.0 if (self.dtype == nat.F8 and float(i8_scale) == 127.0) else float(i8_scale)
        self.h = nat.Handle(int(device))
        self._buf: Optional[torch.Tensor] = None  # aligned uint8 storage of the packed rows
        self._cap_rows = 0
        self._rows = 0  # padded rows in use
        self._page_lens: List[int] = []
        self._attached = False
        if capacity_rows:
            self._grow(int(capacity_rows))

    # ------------------------------------------------------------------ properties
    @property
    def n_pages(self) -> int:
        return len(self._page_lens)

    @property
    def n_rows_padded(self) -> int:
        return self._rows

    @property
    def page_lens(self) -> np.ndarray:
        return np.asarray(self._page_lens, dtype=np.int32)

    @property
    def score_scale(self) -> float:
        """Factor that turns the kernels' raw sums into reference-scale scores."""
        if self.dtype == nat.B1:
            return 1.0 / 128.0  # sum_t max_r (128 - ham) / 128  == SQL max_sim
        if self.dtype in (nat.I8, nat.F8):
            return 1.0 / (self.i8_scale * self.i8_scale)
        return 1.0

    @property
    def float_scores(self) -> bool:
        """True when the kernels emit float32 group scores (bf16 / fp8 corpora), False for exact int32 (int8 / binary)."""
        return self.dtype in (nat.BF16, nat.F8)

    def launch_count(self) -> int:
        return int(nat.lib.b200ms_launch_count(self.h.ptr))

    def last_score_ms(self) -> float:
        ms = float(nat.lib.b200ms_last_score_ms(self.h.ptr))
        if ms < 0:
            self.h.check(int(ms), "b200ms_last_score_ms")
        return ms

    def score_times_ms(self, n: int) -> List[float]:
        """Device times (ms) of the scoring kernels of the last n score/search calls (CUDA events on the launch stream)."""
        buf = (ctypes.c_float * max(n, 1))()
        rc = nat.lib.b200ms_score_times_ms(self.h.ptr, buf, int(n))
        if rc < 0:
            self.h.check(rc, "b200ms_score_times_ms")
        return [float(buf[i]) for i in range(n)]

    def set_tuning(self, unit_rows: int = 0, max_ctas: int = 0):
        self.h.check(nat.lib.b200ms_set_tuning(self.h.ptr, int(unit_rows), int(max_ctas)), "b200ms_set_tuning")
        self._attached = False

    def set_option(self, name: str, value: int):
        self.h.check(nat.lib.b200ms_set_option(self.h.ptr, name.encode(), int(value)), f"b200ms_set_option({name})")
        if name == "unit_rows":
            self._attached = False

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ building the corpus
    def _grow(self, need_rows: int):
        if need_rows <= self._cap_rows:
            return
        new_cap = max(need_rows, int(self._cap_rows * 1.5), 4096)
        new_buf = _aligned_bytes(new_cap * self.row_bytes, self.device)
        if self._buf is not None and self._rows:
            new_buf[: self._rows * self.row_bytes].copy_(self._buf[: self._rows * self.row_bytes])
        self._buf = new_buf
        self._cap_rows = new_cap
        self._attached = False

    def add_pages(self, pages: Sequence[ArrayLike]) -> Tuple[int, int]:
        """Append pages ([P_i,128] float32/bfloat16; numpy on host or torch on either side).  Returns (first id, count).

        Mirrors the write half of store_embeddings (multi_vector_store.py:681-703 quantise+insert;
        fast_multivector_store.py:673-707 np.save of float32): one H2D copy of the ragged rows, one pack kernel.
        """
        first = self.n_pages
        if len(pages) == 0:
            return first, 0
        lens = [int(p.shape[0]) for p in pages]
        for p in pages:
            if p.ndim != 2 or p.shape[1] != nat.DIM:
                raise ValueError(f"page embeddings must be [P,{nat.DIM}], got {tuple(p.shape)}")
        src, src_dtype = self._stage_rows(pages)
        lens_c = nat.i32_array(lens)
        add_rows = int(nat.lib.b200ms_padded_rows(lens_c, len(lens)))
        self._grow(self._rows + add_rows)
        dst = self._buf[self._rows * self.row_bytes:]
        with torch.cuda.device(self.device):
            self.h.check(
                nat.lib.b200ms_pack_pages(self.h.ptr, _vp(src), src_dtype, lens_c, len(lens), _vp(dst), self.dtype,
                                          ctypes.c_float(self.i8_scale), self._stream()),
                "b200ms_pack_pages",
            )
        self._rows += add_rows
        self._page_lens.extend(lens)
        self._attached = False
        return first, len(lens)

    def _stage_rows(self, mats: Sequence[ArrayLike]) -> Tuple[torch.Tensor, int]:
        """Concatenate ragged [n_i,128] matrices into one device tensor (float32 or bfloat16)."""
        if all(isinstance(m, torch.Tensor) for m in mats):
            dt = torch.bfloat16 if all(m.dtype == torch.bfloat16 for m in mats) else torch.float32
            cat = torch.cat([m.to(device=self.device, dtype=dt) for m in mats], dim=0) if len(mats) > 1 else mats[0].to(
                device=self.device, dtype=dt)
            cat = cat.contiguous()
            if cat.shape[0] == 0:
                cat = torch.zeros((1, nat.DIM), dtype=dt, device=self.device)
            return cat, (nat.BF16 if dt == torch.bfloat16 else nat.F32)
        mats = [np.asarray(m.cpu() if isinstance(m, torch.Tensor) else m, dtype=np.float32) for m in mats]
        total = int(sum(m.shape[0] for m in mats))
        if total == 0:
            return torch.zeros((1, nat.DIM), dtype=torch.float32, device=self.device), nat.F32
        if total < 256:  # queries, single small pages: a pageable copy is cheaper than touching the staging buffer
            return torch.from_numpy(np.ascontiguousarray(np.concatenate(mats, axis=0))).to(self.device), nat.F32
        # ingest: gather the ragged pages straight into a reusable PINNED staging buffer and copy once (pageable H2D
        # measured 2.6 GB/s through store_embeddings; the reference hands pages over as host float32, :673-707)
        pin = getattr(self, "_pin", None)
        if pin is None or pin.shape[0] < total:
            pin = self._pin = torch.empty((max(total, 1 << 15), nat.DIM), dtype=torch.float32).pin_memory()
        view = pin[:total].numpy()
        o = 0
        for m in mats:
            view[o:o + m.shape[0]] = m
            o += m.shape[0]
        dev = torch.empty((total, nat.DIM), dtype=torch.float32, device=self.device)
        dev.copy_(pin[:total], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()  # the staging buffer is reused by the next call
        return dev, nat.F32

    def adopt_packed(self, rows: torch.Tensor, page_lens: Sequence[int]):
        """Use an already packed device buffer (uint8 view or typed tensor, 1024-byte aligned, layout as described in the
        class docstring) as the corpus without copying -- bench.py builds its synthetic shard this way."""
        buf = rows.view(torch.uint8).reshape(-1)
        if buf.data_ptr() % 1024:
            raise ValueError("packed corpus must be 1024-byte aligned")
        lens = [int(x) for x in page_lens]
        lens_c = nat.i32_array(lens)
        need = int(nat.lib.b200ms_padded_rows(lens_c, len(lens)))
        if buf.numel() < need * self.row_bytes:
            raise ValueError("packed buffer smaller than the page lengths imply")
        self._buf, self._cap_rows, self._rows = buf, buf.numel() // self.row_bytes, need
        self._page_lens = lens
        self._attached = False

    def append_packed(self, rows: torch.Tensor, page_lens: Sequence[int]) -> Tuple[int, int]:
        """Append already packed rows (uint8 view of the corpus dtype's layout, pages padded to 32 rows) -- a shard-file
        segment -- with one device copy.  Returns (first id, count)."""
        lens = [int(x) for x in page_lens]
        flat = rows.view(torch.uint8).reshape(-1)
        need = int(nat.lib.b200ms_padded_rows(nat.i32_array(lens), len(lens)))
        if flat.numel() < need * self.row_bytes:
            raise ValueError("packed buffer smaller than the page lengths imply")
        first = self.n_pages
        self._grow(self._rows + need)
        self._buf[self._rows * self.row_bytes:(self._rows + need) * self.row_bytes].copy_(flat[: need * self.row_bytes])
        self._rows += need
        self._page_lens.extend(lens)
        self._attached = False
        return first, len(lens)

    def page_row_range(self, first: int, n: int) -> Tuple[int, int]:
        """Padded row range [r0, r1) that pages [first, first+n) occupy in the packed buffer."""
        pad = [(x + 31) // 32 * 32 for x in self._page_lens]
        r0 = int(sum(pad[:first]))
        return r0, r0 + int(sum(pad[first:first + n]))

    def _attach(self):
        if self._attached:
            return
        lens_c = nat.i32_array(self._page_lens)
        with torch.cuda.device(self.device):
            torch.cuda.current_stream(self.device).synchronize()  # pack kernels ran on torch's stream
            self.h.check(nat.lib.b200ms_set_corpus(self.h.ptr, _vp(self._buf), self.dtype, lens_c, self.n_pages),
                         "b200ms_set_corpus")
        self._attached = True

    def clear(self):
        """Drop every page but keep the device buffer and the native handle (scratch indexes, e.g. the reranker's)."""
        self._rows = 0
        self._page_lens = []
        self._attached = False

    def compact(self, keep: Sequence[int]):
        """Keep only the pages in `keep` (ascending old ids), renumbering them 0..len(keep)-1: device-to-device copies of
        the surviving page runs into a fresh buffer (DELETE ... WHERE document_id, multi_vector_store.py:929-933)."""
        keep = [int(p) for p in keep]
        lens = self._page_lens
        pad = [(n + 31) // 32 * 32 for n in lens]
        starts = np.concatenate([[0], np.cumsum(pad)]).astype(np.int64)
        new_rows = int(sum(pad[p] for p in keep))
        new_buf = _aligned_bytes(max(new_rows, 4096) * self.row_bytes, self.device)
        dst = 0
        i = 0
        while i < len(keep):  # copy maximal runs of consecutive surviving pages
            j = i
            while j + 1 < len(keep) and keep[j + 1] == keep[j] + 1:
                j += 1
            a, b = int(starts[keep[i]]) * self.row_bytes, int(starts[keep[j] + 1]) * self.row_bytes
            new_buf[dst:dst + (b - a)].copy_(self._buf[a:b])
            dst += b - a
            i = j + 1
        self._buf, self._cap_rows, self._rows = new_buf, new_buf.numel() // self.row_bytes, new_rows
        self._page_lens = [lens[p] for p in keep]
        self._attached = False

    def packed_rows(self) -> torch.Tensor:
        """The packed corpus as a uint8 tensor [rows_padded, row_bytes] (a view)."""
        if self._buf is None:
            return torch.empty((0, self.row_bytes), dtype=torch.uint8, device=self.device)
        return self._buf[: self._rows * self.row_bytes].view(self._rows, self.row_bytes)

    # ------------------------------------------------------------------ queries
    @staticmethod
    def _q_lens(queries: Sequence[ArrayLike]) -> List[int]:
        for q in queries:
            if q.ndim != 2 or q.shape[1] != nat.DIM:
                raise ValueError(f"query embeddings must be [T,{nat.DIM}], got {tuple(q.shape)}")
        return [int(q.shape[0]) for q in queries]

    def mask_from_pages(self, allowed: np.ndarray) -> np.ndarray:
        """bool[n_pages] -> uint32 bitmask words (bit p&31 of word p>>5), the device filter format."""
        allowed = np.asarray(allowed, dtype=bool)
        if allowed.shape[0] != self.n_pages:
            raise ValueError("mask length must equal n_pages")
        pad = (-len(allowed)) % 32
        bits = np.concatenate([allowed, np.zeros(pad, dtype=bool)]) if pad else allowed
        return np.packbits(bits.reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).reshape(-1).copy()

    def _check_mask(self, allow_mask) -> np.ndarray:
        """The library reads ceil(n_pages/32) words: a mask built for an older (shorter) corpus must never reach it."""
        m = np.ascontiguousarray(allow_mask, dtype=np.uint32)
        words = (self.n_pages + 31) // 32
        if m.shape != (words,):
            raise ValueError(f"allow-mask must have {words} uint32 words for {self.n_pages} pages, got shape {m.shape}")
        return m

    def search_host(self, queries: Sequence[np.ndarray], k: int, allow_mask: Optional[np.ndarray] = None,
                    id_base: int = 0, out: Optional[Tuple[np.ndarray, np.ndarray, np.ndarray]] = None):
        """End-to-end query step from HOST buffers (float32 [T_i,128] each): H2D, pack, scan, top-k, D2H; synchronous.

        Returns (scores float32 [n_q,k], page ids int64 [n_q,k], counts int32 [n_q]); unused slots are -inf / -1.
        Replaces the scoring half of query_similar (multi_vector_store.py:721-763 / fast_multivector_store.py:553-557).
        """
        self._attach()
        lens = self._q_lens(queries)
        n_q = len(lens)
        if n_q == 1:
            q_host = np.ascontiguousarray(queries[0], dtype=np.float32)
        else:
            q_host = np.ascontiguousarray(np.concatenate([np.asarray(q, dtype=np.float32) for q in queries], axis=0))
        if out is None:
            out = (np.empty((n_q, k), np.float32), np.empty((n_q, k), np.int64), np.empty((n_q,), np.int32))
        ts, ti, tc = out
        mask_p = None
        if allow_mask is not None:
            allow_mask = self._check_mask(allow_mask)
            mask_p = allow_mask.ctypes.data_as(ctypes.c_void_p)
        self.h.check(
            nat.lib.b200ms_search_host(self.h.ptr, q_host.ctypes.data_as(ctypes.c_void_p), nat.i32_array(lens), n_q, int(k),
                                       mask_p, ctypes.c_float(self.i8_scale), ctypes.c_float(self.score_scale),
                                       int(id_base), ts.ctypes.data_as(ctypes.c_void_p),
                                       ti.ctypes.data_as(ctypes.c_void_p), tc.ctypes.data_as(ctypes.c_void_p)),
            "b200ms_search_host",
        )
        return ts, ti, tc

    def search_host_masked(self, queries: Sequence[np.ndarray], k: int, allow_masks: Sequence[Optional[np.ndarray]],
                           id_base: int = 0):
        """search_host with one allow-mask PER QUERY (``allow_masks[i]``: uint32 words as for ``allow_mask``, or None =
        unfiltered): queries of different users -- different authorised ``doc_ids`` (document_service.py:408-417) -- share one
        pass over the corpus.  Identical mask objects/contents are uploaded once."""
        self._attach()
        lens = self._q_lens(queries)
        n_q = len(lens)
        if len(allow_masks) != n_q:
            raise ValueError("one allow-mask (or None) per query")
        q_host = np.ascontiguousarray(np.concatenate([np.asarray(q, dtype=np.float32) for q in queries], axis=0))
        words = (self.n_pages + 31) // 32
        rows, index, seen = [], [], {}
        for m in allow_masks:
            if m is None:
                index.append(-1)
                continue
            m = np.ascontiguousarray(m, dtype=np.uint32)
            if m.shape != (words,):
                raise ValueError(f"allow-mask must have {words} uint32 words, got {m.shape}")
            key = m.tobytes()
            if key not in seen:
                seen[key] = len(rows)
                rows.append(m)
            index.append(seen[key])
        ts, ti, tc = np.empty((n_q, k), np.float32), np.empty((n_q, k), np.int64), np.empty((n_q,), np.int32)
        mat = np.stack(rows) if rows else None
        self.h.check(
            nat.lib.b200ms_search_host_masked(
                self.h.ptr, q_host.ctypes.data_as(ctypes.c_void_p), nat.i32_array(lens), n_q, int(k),
                None if mat is None else mat.ctypes.data_as(ctypes.c_void_p), len(rows), nat.i32_array(index),
                ctypes.c_float(self.i8_scale), ctypes.c_float(self.score_scale), int(id_base),
                ts.ctypes.data_as(ctypes.c_void_p), ti.ctypes.data_as(ctypes.c_void_p), tc.ctypes.data_as(ctypes.c_void_p)),
            "b200ms_search_host_masked",
        )
        return ts, ti, tc

    def search_host_flat(self, q_host: Union[np.ndarray, torch.Tensor], q_lens: Sequence[int], k: int,
                         out_scores: torch.Tensor, out_ids: torch.Tensor, out_counts: torch.Tensor,
                         allow_mask: Optional[np.ndarray] = None, id_base: int = 0):
        """Same as search_host for callers that keep (pinned) host buffers: q_host is [sum T,128] float32.
        The ctypes argument objects of the previous call are reused when nothing changed (interactive callers repeat the
        same shapes; building them costs as much as the whole C call for a 100-page corpus)."""
        if not self._attached:
            self._attach()
        qp = q_host.data_ptr() if isinstance(q_host, torch.Tensor) else q_host.ctypes.data
        mask_p = None
        if allow_mask is not None:
            allow_mask = self._check_mask(allow_mask)
            mask_p = allow_mask.ctypes.data_as(ctypes.c_void_p)
        key = (tuple(q_lens), int(k), int(id_base), out_scores.data_ptr(), out_ids.data_ptr(), out_counts.data_ptr())
        cached = getattr(self, "_flat_args", None)
        if cached is None or cached[0] != key:
            cached = (key, nat.i32_array(q_lens), ctypes.c_float(self.i8_scale), ctypes.c_float(self.score_scale),
                      ctypes.c_void_p(out_scores.data_ptr()), ctypes.c_void_p(out_ids.data_ptr()),
                      ctypes.c_void_p(out_counts.data_ptr()))
            self._flat_args = cached
        _, lens_c, c_scale, c_sscale, p_s, p_i, p_c = cached
        rc = nat.lib.b200ms_search_host(self.h.ptr, qp, lens_c, len(q_lens), key[1], mask_p, c_scale, c_sscale, key[2], p_s, p_i, p_c)
        if rc != 0:
            self.h.check(rc, "b200ms_search_host")

    def search_device(self, q_dev: torch.Tensor, q_lens: Sequence[int], k: int, allow_mask_dev: Optional[torch.Tensor] = None,
                      id_base: int = 0, out: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = None,
                      mask_index_dev: Optional[torch.Tensor] = None):
        """Asynchronous search with device-resident queries ([sum T,128] float32|bfloat16) on torch's current stream.

        ``allow_mask_dev``: one mask (int32/uint32 words) for every query, or -- with ``mask_index_dev`` (int32 [n_q], -1 =
        unfiltered) -- a [n_masks, words] matrix with one row per distinct filter (b200ms_search_device_masked)."""
        self._attach()
        n_q = len(q_lens)
        if q_dev.dtype not in (torch.float32, torch.bfloat16) or not q_dev.is_contiguous():
            raise ValueError("q_dev must be a contiguous float32/bfloat16 tensor")
        if out is None:
            out = (torch.empty((n_q, k), dtype=torch.float32, device=self.device),
                   torch.empty((n_q, k), dtype=torch.int64, device=self.device),
                   torch.empty((n_q,), dtype=torch.int32, device=self.device))
        ts, ti, tc = out
        if mask_index_dev is not None:
            if (allow_mask_dev is None or allow_mask_dev.ndim != 2 or allow_mask_dev.shape[1] != (self.n_pages + 31) // 32
                    or not allow_mask_dev.is_contiguous()):
                raise ValueError("per-query masks: allow_mask_dev must be a contiguous [n_masks, ceil(n_pages/32)] word matrix")
            if mask_index_dev.dtype != torch.int32 or mask_index_dev.numel() != n_q:
                raise ValueError("mask_index_dev must be int32 [n_q]")
            with torch.cuda.device(self.device):
                self.h.check(
                    nat.lib.b200ms_search_device_masked(
                        self.h.ptr, _vp(q_dev), nat.BF16 if q_dev.dtype == torch.bfloat16 else nat.F32, nat.i32_array(q_lens),
                        n_q, int(k), _vp(allow_mask_dev), int(allow_mask_dev.shape[0]), _vp(mask_index_dev),
                        ctypes.c_float(self.i8_scale), ctypes.c_float(self.score_scale), int(id_base), _vp(ts), _vp(ti),
                        _vp(tc), self._stream()),
                    "b200ms_search_device_masked",
                )
            return ts, ti, tc
        with torch.cuda.device(self.device):
            self.h.check(
                nat.lib.b200ms_search_device(self.h.ptr, _vp(q_dev), nat.BF16 if q_dev.dtype == torch.bfloat16 else nat.F32,
                                             nat.i32_array(q_lens), n_q, int(k), _vp(allow_mask_dev),
                                             ctypes.c_float(self.i8_scale), ctypes.c_float(self.score_scale), int(id_base),
                                             _vp(ts), _vp(ti), _vp(tc), self._stream()),
                "b200ms_search_device",
            )
        return ts, ti, tc

    def score_groups(self, queries: Sequence[ArrayLike]):
        """Raw kernel output for tests and diagnostics: (group_scores tensor [n_groups_padded, ld] float32|int32,
        group_offsets int32 [n_q+1], n_pages).  Calls pack_queries + score through the C-ABI."""
        self._attach()
        lens = self._q_lens(queries)
        n_q = len(lens)
        src, src_dtype = self._stage_rows(queries)
        lens_c = nat.i32_array(lens)
        groups = int(nat.lib.b200ms_query_groups(lens_c, n_q))
        gp = max((groups + 3) // 4 * 4, 4)
        q_packed = _aligned_bytes(gp * 32 * self.row_bytes, self.device)
        goff = (ctypes.c_int32 * (n_q + 1))()
        ng = ctypes.c_int(0)
        ld = (self.n_pages + 31) // 32 * 32
        sdt = torch.float32 if self.float_scores else torch.int32
        scores = torch.zeros((gp, max(ld, 32)), dtype=sdt, device=self.device)
        with torch.cuda.device(self.device):
            self.h.check(
                nat.lib.b200ms_pack_queries(self.h.ptr, _vp(src), src_dtype, lens_c, n_q, _vp(q_packed), self.dtype,
                                            ctypes.c_float(self.i8_scale), goff, ctypes.byref(ng), self._stream()),
                "b200ms_pack_queries",
            )
            self.h.check(
                nat.lib.b200ms_score(self.h.ptr, _vp(q_packed), ng.value, lens_c, goff, n_q, _vp(scores), max(ld, 32),
                                     self._stream()),
                "b200ms_score",
            )
        return scores, np.asarray(list(goff), dtype=np.int32), self.n_pages, q_packed

    def score_matrix(self, queries: Sequence[ArrayLike]) -> np.ndarray:
        """[n_q, n_pages] scores on the reference's scale (float64 on host; the per-query sum over 32-token groups is done
        here in float64 for diagnostics only -- the product path sums inside the top-k kernel)."""
        scores, goff, n_pages, _ = self.score_groups(queries)
        torch.cuda.synchronize(self.device)
        s = scores[:, :n_pages].cpu().numpy().astype(np.float64)
        out = np.zeros((len(goff) - 1, n_pages), dtype=np.float64)
        for q in range(len(goff) - 1):
            out[q] = s[goff[q]:goff[q + 1]].sum(axis=0) * self.score_scale
        return out

    def sign_pack(self, x: ArrayLike) -> torch.Tensor:
        """[n,128] float32/bfloat16 -> uint8 [n,16] MSB-first sign bits on the device (fast_ops.binary_quantize_packed)."""
        src, src_dtype = self._stage_rows([x])
        n = int(x.shape[0])
        out = torch.empty((max(n, 1), 16), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            self.h.check(nat.lib.b200ms_sign_pack(self.h.ptr, _vp(src), src_dtype, n, _vp(out), self._stream()),
                         "b200ms_sign_pack")
        return out[:n]

    def hamming_distance_batch(self, query_bits, candidate_bits) -> torch.Tensor:
        """fast_ops.hamming_distance_batch on the device: packed 16-byte query (bytes / uint8[16]) vs packed candidates
        (list of bytes / uint8 [n,16]) -> int32 tensor [n] of Hamming distances."""
        q = np.frombuffer(bytes(query_bits), dtype=np.uint8) if isinstance(query_bits, (bytes, bytearray)) else np.asarray(query_bits, np.uint8)
        if isinstance(candidate_bits, (list, tuple)):
            lens = {len(c) for c in candidate_bits}
            if lens - {16}:
                raise ValueError(f"Vector length mismatch: candidates must be 16 bytes like the query, got {sorted(lens)}")
            c = np.frombuffer(b"".join(bytes(x) for x in candidate_bits), dtype=np.uint8).reshape(-1, 16)
        else:
            c = candidate_bits
        if q.size != 16:
            raise ValueError(f"Vector length mismatch: {q.size} vs 16")  # ValueError like the Rust implementation
        qd = torch.from_numpy(q.reshape(16).copy()).to(self.device)
        cd = c if isinstance(c, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(c, dtype=np.uint8)).to(self.device)
        cd = cd.contiguous().view(-1, 16)
        qa = _aligned_bytes(16, self.device, 16); qa.copy_(qd)
        ca = _aligned_bytes(max(cd.numel(), 16), self.device, 16); ca[: cd.numel()].copy_(cd.reshape(-1))
        out = torch.empty((cd.shape[0],), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            self.h.check(nat.lib.b200ms_hamming_batch(self.h.ptr, _vp(qa), _vp(ca), cd.shape[0], _vp(out), self._stream()),
                         "b200ms_hamming_batch")
        return out

    def merge_topk(self, cand_scores: torch.Tensor, cand_ids: torch.Tensor, k: int):
        """Merge candidate lists [n_q, m] (ids < 0 ignored) -> top-k by (score desc, id asc), on the device."""
        n_q, m = cand_scores.shape
        ts = torch.empty((n_q, k), dtype=torch.float32, device=self.device)
        ti = torch.empty((n_q, k), dtype=torch.int64, device=self.device)
        tc = torch.empty((n_q,), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            self.h.check(
                nat.lib.b200ms_merge_topk(self.h.ptr, _vp(cand_scores.contiguous()), _vp(cand_ids.contiguous()), n_q, m, int(k),
                                          _vp(ts), _vp(ti), _vp(tc), self._stream()),
                "b200ms_merge_topk",
            )
        return ts, ti, tc

    # ------------------------------------------------------------------ candidate rerank (two-stage search)
    def rerank_batch(self, q_dev: torch.Tensor, q_lens: Sequence[int], cand_ids: torch.Tensor, k: int,
                     out: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = None):
        """Exact MaxSim of every query over ITS OWN candidate list: cand_ids int64 [n_q, n_cand] on the device (-1 = unused
        slot) -> (scores [n_q,k], page ids [n_q,k], counts [n_q]) on the device; one call for the whole batch
        (b200ms_rerank_batch_device; fast_multivector_store.py:545-557 does this one query at a time on <= 75 pages)."""
        self._attach()
        n_q = len(q_lens)
        if cand_ids.dtype != torch.int64 or cand_ids.ndim != 2 or cand_ids.shape[0] != n_q or not cand_ids.is_contiguous():
            raise ValueError("cand_ids must be a contiguous int64 [n_q, n_cand] device tensor")
        n_cand = int(cand_ids.shape[1])
        kk = int(min(k, n_cand))
        if out is None:
            out = (torch.empty((n_q, kk), dtype=torch.float32, device=self.device),
                   torch.empty((n_q, kk), dtype=torch.int64, device=self.device),
                   torch.empty((n_q,), dtype=torch.int32, device=self.device))
        ts, ti, tc = out
        with torch.cuda.device(self.device):
            self.h.check(
                nat.lib.b200ms_rerank_batch_device(self.h.ptr, _vp(q_dev), nat.BF16 if q_dev.dtype == torch.bfloat16 else nat.F32,
                                                   nat.i32_array(q_lens), n_q, _vp(cand_ids), n_cand, kk,
                                                   ctypes.c_float(self.i8_scale), ctypes.c_float(self.score_scale), _vp(ts),
                                                   _vp(ti), _vp(tc), self._stream()),
                "b200ms_rerank_batch_device")
        return ts, ti, tc

    # ------------------------------------------------------------------ multi-GPU (document shards, NCCL inside libb200ms)
    @staticmethod
    def comm_unique_id() -> bytes:
        """128-byte NCCL unique id (rank 0 creates it; ship it to every rank by any channel)."""
        buf = (ctypes.c_uint8 * 128)()
        rc = nat.lib.b200ms_comm_unique_id(buf)
        if rc != 0:
            msg = nat.lib.b200ms_last_error(None)
            raise nat.NativeError(f"b200ms_comm_unique_id failed ({rc}): {msg.decode() if msg else ''}")
        return bytes(buf)

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        """Collective: create this handle's NCCL communicator (libb200ms dlopens libnccl; torch is not involved)."""
        if len(unique_id) != 128:
            raise ValueError("unique_id must be the 128 bytes of comm_unique_id()")
        buf = (ctypes.c_uint8 * 128).from_buffer_copy(unique_id)
        with torch.cuda.device(self.device):
            self.h.check(nat.lib.b200ms_comm_init(self.h.ptr, buf, int(rank), int(world)), "b200ms_comm_init")

    def bcast(self, t: torch.Tensor, root: int = 0):
        """ncclBroadcast of a contiguous device tensor in place (e.g. the query rows of a batch) on torch's current stream."""
        with torch.cuda.device(self.device):
            self.h.check(nat.lib.b200ms_bcast_device(self.h.ptr, _vp(t), t.numel() * t.element_size(), int(root), self._stream()),
                         "b200ms_bcast_device")
        return t

    def send(self, t: torch.Tensor, peer: int):
        """ncclSend of a contiguous device tensor to rank `peer` over the handle's communicator (torch's current stream)."""
        with torch.cuda.device(self.device):
            self.h.check(nat.lib.b200ms_send_device(self.h.ptr, _vp(t), t.numel() * t.element_size(), int(peer), self._stream()),
                         "b200ms_send_device")

    def recv(self, t: torch.Tensor, peer: int):
        with torch.cuda.device(self.device):
            self.h.check(nat.lib.b200ms_recv_device(self.h.ptr, _vp(t), t.numel() * t.element_size(), int(peer), self._stream()),
                         "b200ms_recv_device")
        return t

    def allgather_topk(self, xchg: torch.Tensor, n_q: int, k: int, out=None):
        """The one collective of the path for a caller-made local list: xchg = uint8 device buffer in the exchange layout
        ([n_q*k int64 global ids][n_q*k float32 scores]) -> merged (scores, ids, counts), identical on every rank."""
        if out is None:
            out = (torch.empty((n_q, k), dtype=torch.float32, device=self.device),
                   torch.empty((n_q, k), dtype=torch.int64, device=self.device),
                   torch.empty((n_q,), dtype=torch.int32, device=self.device))
        ts, ti, tc = out
        with torch.cuda.device(self.device):
            self.h.check(nat.lib.b200ms_allgather_topk(self.h.ptr, None, _vp(xchg), int(n_q), int(k), _vp(ts), _vp(ti), _vp(tc),
                                                       self._stream()), "b200ms_allgather_topk")
        return ts, ti, tc

    def sharded_search_begin(self, q_dev: torch.Tensor, q_lens: Sequence[int], k: int, id_base: int,
                             out: Tuple[torch.Tensor, torch.Tensor, torch.Tensor],
                             allow_masks_dev: Optional[torch.Tensor] = None, mask_index_dev: Optional[torch.Tensor] = None) -> int:
        """Enqueue local scan + top-k on torch's current stream and the all-gather + merge behind it on the handle's
        communication stream; returns a ticket.  The current stream does not wait for the other ranks."""
        self._attach()
        ts, ti, tc = out
        n_masks = 0 if allow_masks_dev is None else int(allow_masks_dev.shape[0])
        with torch.cuda.device(self.device):
            t = int(nat.lib.b200ms_sharded_search_begin(
                self.h.ptr, _vp(q_dev), nat.BF16 if q_dev.dtype == torch.bfloat16 else nat.F32, nat.i32_array(q_lens), len(q_lens),
                int(k), _vp(allow_masks_dev), n_masks, _vp(mask_index_dev), ctypes.c_float(self.i8_scale),
                ctypes.c_float(self.score_scale), int(id_base), _vp(ts), _vp(ti), _vp(tc), self._stream()))
        if t < 0:
            self.h.check(t, "b200ms_sharded_search_begin")
        return t

    def sharded_search_end(self, ticket: int):
        """Make torch's current stream wait for that ticket's merged result (in the buffers given to begin)."""
        with torch.cuda.device(self.device):
            self.h.check(nat.lib.b200ms_sharded_search_end(self.h.ptr, int(ticket), self._stream()), "b200ms_sharded_search_end")

    def sharded_search_host_begin(self, q_host: torch.Tensor, q_lens: Sequence[int], k: int, id_base: int) -> int:
        """Host-buffer form: q_host float32 [sum T,128] (pinned for an asynchronous copy); returns a ticket."""
        self._attach()
        t = int(nat.lib.b200ms_sharded_search_host_begin(
            self.h.ptr, ctypes.c_void_p(q_host.data_ptr()), nat.i32_array(q_lens), len(q_lens), int(k), None, 0, None,
            ctypes.c_float(self.i8_scale), ctypes.c_float(self.score_scale), int(id_base)))
        if t < 0:
            self.h.check(t, "b200ms_sharded_search_host_begin")
        return t

    def sharded_search_host_end(self, ticket: int, out_scores: torch.Tensor, out_ids: torch.Tensor, out_counts: torch.Tensor):
        """Blocks until the ticket's merged top-k is on the host; fills the three (host) tensors."""
        self.h.check(nat.lib.b200ms_sharded_search_host_end(self.h.ptr, int(ticket), ctypes.c_void_p(out_scores.data_ptr()),
                                                            ctypes.c_void_p(out_ids.data_ptr()),
                                                            ctypes.c_void_p(out_counts.data_ptr())),
                     "b200ms_sharded_search_host_end")

    def close(self):
        self.h.close()
        self._buf = None
