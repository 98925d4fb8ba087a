"""Fixed-dimensional encodings (MUVERA FDE) + two-stage search: FDE candidate scan -> exact MaxSim rerank.

API mirror of the reference's ``fixed_dimensional_encoding`` extension module as Morphik uses it
(core/vector_store/fast_multivector_store.py:325-331, 447-449, 521):

    cfg = FixedDimensionalEncodingConfig(dimension=128, num_repetitions=20, num_simhash_projections=5,
                                         projection_dimension=16, projection_type="AMS_SKETCH")
    q_fde = generate_query_encoding(q, cfg)        # np.ndarray [n,128] -> np.float32 [10240]
    d_fde = generate_document_encoding(p, cfg)

and ``TwoStageIndex``, which replaces stages 1-4 of FastMultiVectorStore.query_similar (:521-557): the query FDE, the
Turbopuffer ANN over document FDEs (cosine distance, ``top_k = min(10*k, 75)`` there; any candidate count here), the
gather of candidate multivectors and the MaxSim rerank -- all on the GPU, with the multivectors already resident.

PARITY NOTE: the extension's C++ sources are not in the reference snapshot, so the exact random streams of
the upstream implementation (std::mt19937 based) cannot be reproduced or checked here.  The construction (SimHash
partitions with Gray-code index, AMS sketch, SUM for queries / AVERAGE for documents, empty partitions zero or -- with
``fill_empty_partitions`` -- the projection of the nearest point, optional final count-sketch projection) follows the
published algorithm; by default the matrices come from ``numpy.random.default_rng(seed + repetition)``, so FDEs produced
here are self-consistent (query and document sides share the matrices) but NOT interchangeable with vectors already
stored in a Turbopuffer namespace by the reference.

LOADING UPSTREAM MATRICES: the random matrices are INPUTS of the C-ABI (``b200ms_fde_configure_ex``).  A deployment that
must stay interchangeable with vectors written by the reference dumps the extension's matrices once (SimHash Gaussian
[R,128,K], AMS bucket index / sign [R,128], and for ``final_projection_dimension`` the count-sketch index / sign
[R*2^K*P]) and passes them through ``FdeMatrices`` / ``matrices_from_npz`` -- every consumer below takes an optional
``matrices`` argument.  The encoder then reproduces upstream bit for bit up to fp32 summation order.
"""
from __future__ import annotations

import ctypes
import threading
from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as nat
from .index import MaxSimIndex, _vp


@dataclass(frozen=True)
class FixedDimensionalEncodingConfig:
    """Field names follow the upstream config message the reference fills at fast_multivector_store.py:325-331 (it sets
    only the first five; ``fill_empty_partitions`` and ``final_projection_dimension`` stay at their defaults there)."""
    dimension: int = 128
    num_repetitions: int = 20
    num_simhash_projections: int = 5
    projection_dimension: int = 16
    projection_type: str = "AMS_SKETCH"
    seed: int = 1
    fill_empty_partitions: bool = False
    final_projection_dimension: int = 0  # > 0: count-sketch the whole encoding down to this many dims

    @property
    def num_partitions(self) -> int:
        return 1 << self.num_simhash_projections

    @property
    def inner_dimension(self) -> int:
        return self.num_repetitions * self.num_partitions * self.projection_dimension

    @property
    def fde_dimension(self) -> int:
        return self.final_projection_dimension if self.final_projection_dimension > 0 else self.inner_dimension

    @property
    def scale(self) -> float:
        return 1.0 / float(np.sqrt(self.projection_dimension))

    def validate(self) -> None:
        if self.dimension != nat.DIM:
            raise ValueError(f"dimension must be {nat.DIM}")
        if self.projection_type != "AMS_SKETCH":
            raise ValueError("only projection_type='AMS_SKETCH' (the reference's setting) is implemented")
        if self.final_projection_dimension < 0 or self.final_projection_dimension % 8:
            raise ValueError("final_projection_dimension must be 0 or a positive multiple of 8")


@dataclass(frozen=True)
class FdeMatrices:
    """The random matrices of one configuration -- generated here (``fde_matrices_full``) or dumped from the upstream
    extension (``matrices_from_npz``)."""
    simhash: np.ndarray      # float32 [R,128,K] Gaussian SimHash projections
    ams_index: np.ndarray    # int32   [R,128]   AMS sketch: bucket of every input dim, in [0, projection_dimension)
    ams_sign: np.ndarray     # float32 [R,128]   +-1
    final_index: Optional[np.ndarray] = None  # int32   [inner_dimension] final count sketch: output dim of every entry
    final_sign: Optional[np.ndarray] = None   # float32 [inner_dimension] +-1


def fde_matrices_full(cfg: FixedDimensionalEncodingConfig) -> FdeMatrices:
    cfg.validate()
    R, K, P = cfg.num_repetitions, cfg.num_simhash_projections, cfg.projection_dimension
    simhash = np.empty((R, cfg.dimension, K), dtype=np.float32)
    ams_index = np.empty((R, cfg.dimension), dtype=np.int32)
    ams_sign = np.empty((R, cfg.dimension), dtype=np.float32)
    for r in range(R):
        rng = np.random.default_rng(cfg.seed + r)
        simhash[r] = rng.standard_normal((cfg.dimension, K)).astype(np.float32)
        ams_index[r] = rng.integers(0, P, size=cfg.dimension).astype(np.int32)
        ams_sign[r] = np.where(rng.integers(0, 2, size=cfg.dimension) == 1, 1.0, -1.0).astype(np.float32)
    fi = fs = None
    if cfg.final_projection_dimension > 0:
        rng = np.random.default_rng(cfg.seed + 1_000_003)
        fi = rng.integers(0, cfg.final_projection_dimension, size=cfg.inner_dimension).astype(np.int32)
        fs = np.where(rng.integers(0, 2, size=cfg.inner_dimension) == 1, 1.0, -1.0).astype(np.float32)
    return FdeMatrices(simhash, ams_index, ams_sign, fi, fs)


def fde_matrices(cfg: FixedDimensionalEncodingConfig) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(simhash [R,128,K] float32 Gaussian, ams_index [R,128] int32 in [0,proj), ams_sign [R,128] float32 +-1)."""
    m = fde_matrices_full(cfg)
    return m.simhash, m.ams_index, m.ams_sign


def matrices_from_npz(path: str) -> FdeMatrices:
    """Load matrices dumped from the upstream ``fixed_dimensional_encoding`` extension (arrays ``simhash``, ``ams_index``,
    ``ams_sign`` and optionally ``final_index`` / ``final_sign``) so that encodings match vectors the reference stored."""
    z = np.load(path)
    return FdeMatrices(np.ascontiguousarray(z["simhash"], np.float32), np.ascontiguousarray(z["ams_index"], np.int32),
                       np.ascontiguousarray(z["ams_sign"], np.float32),
                       np.ascontiguousarray(z["final_index"], np.int32) if "final_index" in z else None,
                       np.ascontiguousarray(z["final_sign"], np.float32) if "final_sign" in z else None)


def configure_handle(index: MaxSimIndex, cfg: FixedDimensionalEncodingConfig, matrices: Optional[FdeMatrices] = None) -> None:
    cfg.validate()
    m = matrices or fde_matrices_full(cfg)
    R, K, P = cfg.num_repetitions, cfg.num_simhash_projections, cfg.projection_dimension
    if m.simhash.shape != (R, cfg.dimension, K) or m.ams_index.shape != (R, cfg.dimension) or m.ams_sign.shape != (R, cfg.dimension):
        raise ValueError("FDE matrices do not match the configuration")
    fd = cfg.final_projection_dimension
    if fd > 0 and (m.final_index is None or m.final_sign is None or m.final_index.shape != (cfg.inner_dimension,)):
        raise ValueError("final_projection_dimension needs final_index / final_sign of inner_dimension entries")
    vp = lambda a: None if a is None else np.ascontiguousarray(a).ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    sh, ai, sg = (np.ascontiguousarray(m.simhash, np.float32), np.ascontiguousarray(m.ams_index, np.int32),
                  np.ascontiguousarray(m.ams_sign, np.float32))
    fi = None if fd == 0 else np.ascontiguousarray(m.final_index, np.int32)
    fs = None if fd == 0 else np.ascontiguousarray(m.final_sign, np.float32)
    index.h.check(
        nat.lib.b200ms_fde_configure_ex(index.h.ptr, R, K, P, ctypes.c_float(cfg.scale), vp(sh), vp(ai), vp(sg),
                                        int(cfg.fill_empty_partitions), int(fd), vp(fi), vp(fs)),
        "b200ms_fde_configure_ex")


def encode_items(index: MaxSimIndex, items: Sequence, is_document: bool, fde_dim: int) -> torch.Tensor:
    """FDEs of ragged [n_i,128] items -> float32 device tensor [len(items), fde_dim] (one kernel launch per 65535 items)."""
    out = torch.empty((len(items), fde_dim), dtype=torch.float32, device=index.device)
    step = 65535
    for i0 in range(0, len(items), step):
        part = items[i0:i0 + step]
        lens = [int(x.shape[0]) for x in part]
        src, src_dtype = index._stage_rows(part)
        with torch.cuda.device(index.device):
            index.h.check(
                nat.lib.b200ms_fde_encode(index.h.ptr, _vp(src), src_dtype, nat.i32_array(lens), len(lens), int(is_document),
                                          _vp(out[i0:i0 + len(part)]), index._stream()),
                "b200ms_fde_encode")
    return out


# ------------------------------------------------------------------ module-level API (drop-in for the extension module)
_default_lock = threading.Lock()
_default_encoders: Dict[Tuple[int, FixedDimensionalEncodingConfig], MaxSimIndex] = {}


def _encoder(cfg: FixedDimensionalEncodingConfig, device: int = 0) -> MaxSimIndex:
    with _default_lock:
        key = (device, cfg)
        if key not in _default_encoders:
            idx = MaxSimIndex(device=device, dtype="bf16")  # only its handle/stream are used
            configure_handle(idx, cfg)
            _default_encoders[key] = idx
        return _default_encoders[key]


def generate_query_encoding(point_cloud, config: FixedDimensionalEncodingConfig, device: int = 0) -> np.ndarray:
    """fde.generate_query_encoding(np.ndarray [n,128], cfg) -> np.float32 [fde_dimension] (fast_multivector_store.py:521)."""
    x = np.ascontiguousarray(np.asarray(point_cloud, dtype=np.float32).reshape(-1, nat.DIM))
    idx = _encoder(config, device)
    return encode_items(idx, [x], False, config.fde_dimension)[0].cpu().numpy()


def generate_document_encoding(point_cloud, config: FixedDimensionalEncodingConfig, device: int = 0) -> np.ndarray:
    """fde.generate_document_encoding(np.ndarray [n,128], cfg) (fast_multivector_store.py:447-449)."""
    x = np.ascontiguousarray(np.asarray(point_cloud, dtype=np.float32).reshape(-1, nat.DIM))
    idx = _encoder(config, device)
    return encode_items(idx, [x], True, config.fde_dimension)[0].cpu().numpy()


# ------------------------------------------------------------------ two-stage index
class TwoStageIndex:
    """MaxSimIndex + a dense FDE matrix [N, fde_dim] bf16 with per-row inverse norms.

    search(): query FDE (device) -> exhaustive cosine scan of the FDE matrix (HBM-bound, 2*fde_dim bytes per page) -> top
    ``n_candidates`` per query -> exact MaxSim over those pages only (b200ms_rerank_device) -> top-k.
    """

    def __init__(self, device: int = 0, dtype: str = "bf16", config: Optional[FixedDimensionalEncodingConfig] = None,
                 matrices: Optional[FdeMatrices] = None, index: Optional[MaxSimIndex] = None):
        self.cfg = config or FixedDimensionalEncodingConfig()
        self.index = index if index is not None else MaxSimIndex(device=device, dtype=dtype)
        configure_handle(self.index, self.cfg, matrices)
        self.device = self.index.device
        self.fde_dim = self.cfg.fde_dimension
        self._F: Optional[torch.Tensor] = None  # bf16 [cap, fde_dim]
        self._inv: Optional[torch.Tensor] = None  # float32 [cap]
        self._n = 0
        self.last_timing_ms: Dict[str, float] = {}

    @property
    def n_pages(self) -> int:
        return self._n

    def _grow(self, need: int) -> None:
        cap = 0 if self._F is None else self._F.shape[0]
        if need <= cap:
            return
        new_cap = max(need, int(cap * 1.5), 256)
        F = torch.empty((new_cap, self.fde_dim), dtype=torch.bfloat16, device=self.device)
        inv = torch.empty((new_cap,), dtype=torch.float32, device=self.device)
        if self._n:
            F[: self._n].copy_(self._F[: self._n])
            inv[: self._n].copy_(self._inv[: self._n])
        self._F, self._inv = F, inv

    def add_pages(self, pages: Sequence) -> Tuple[int, int]:
        first, n = self.index.add_pages(pages)
        if n == 0:
            return first, 0
        fde = encode_items(self.index, list(pages), True, self.fde_dim)
        self._grow(self._n + n)
        with torch.cuda.device(self.device):
            self.index.h.check(
                nat.lib.b200ms_fde_finalize(self.index.h.ptr, _vp(fde), n, _vp(self._F[self._n:]), _vp(self._inv[self._n:]),
                                            self.index._stream()),
                "b200ms_fde_finalize")
        self._n += n
        return first, n

    def adopt_packed(self, rows: torch.Tensor, page_lens: Sequence[int], batch_pages: int = 16384) -> None:
        """Adopt an already packed bf16 device corpus (see MaxSimIndex.adopt_packed) and build its FDE matrix from the packed
        rows directly.  Requires every page length to be a multiple of 32 (no padding rows inside the packed buffer)."""
        lens = [int(x) for x in page_lens]
        if self.index.dtype != nat.BF16 or any(n % 32 for n in lens):
            raise ValueError("adopt_packed needs a bf16 corpus whose page lengths are multiples of 32")
        self.index.adopt_packed(rows, lens)
        n = len(lens)
        self._grow(n)
        flat = rows.view(torch.uint8).reshape(-1)
        tmp = torch.empty((min(batch_pages, max(n, 1)), self.fde_dim), dtype=torch.float32, device=self.device)
        row0 = 0
        with torch.cuda.device(self.device):
            for i0 in range(0, n, batch_pages):
                part = lens[i0:i0 + batch_pages]
                src = flat[row0 * 256:]
                self.index.h.check(
                    nat.lib.b200ms_fde_encode(self.index.h.ptr, _vp(src), nat.BF16, nat.i32_array(part), len(part), 1, _vp(tmp),
                                              self.index._stream()), "b200ms_fde_encode")
                self.index.h.check(
                    nat.lib.b200ms_fde_finalize(self.index.h.ptr, _vp(tmp), len(part), _vp(self._F[i0:]), _vp(self._inv[i0:]),
                                                self.index._stream()), "b200ms_fde_finalize")
                row0 += sum(part)
        self._n = n

    def rebuild_from_index(self, batch_pages: int = 16384) -> None:
        """(Re)build the FDE matrix for every page of the attached bf16 corpus from its packed rows
        (b200ms_fde_encode_corpus) -- the path of ``store.load(..., fde_candidates=N)`` and of a journal replay."""
        if self.index.dtype != nat.BF16:
            raise ValueError("rebuilding FDEs from packed rows needs a bf16 corpus (other dtypes persist their FDE rows)")
        self.index._attach()
        n = self.index.n_pages
        self._n = 0
        self._grow(n)
        tmp = torch.empty((min(batch_pages, max(n, 1)), self.fde_dim), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            for i0 in range(0, n, batch_pages):
                m = min(batch_pages, n - i0)
                self.index.h.check(nat.lib.b200ms_fde_encode_corpus(self.index.h.ptr, i0, m, _vp(tmp), self.index._stream()),
                                   "b200ms_fde_encode_corpus")
                self.index.h.check(nat.lib.b200ms_fde_finalize(self.index.h.ptr, _vp(tmp), m, _vp(self._F[i0:]),
                                                               _vp(self._inv[i0:]), self.index._stream()), "b200ms_fde_finalize")
        self._n = n

    def fde_rows(self, first: int = 0, n: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """(bf16 [n, fde_dim], inv_norm float32 [n]) views of pages [first, first+n) -- what a journal segment persists."""
        n = self._n - first if n is None else n
        return self._F[first:first + n], self._inv[first:first + n]

    def append_fde_rows(self, rows: torch.Tensor, inv: torch.Tensor) -> None:
        """Adopt persisted FDE rows for pages already appended to the MaxSim corpus (journal replay)."""
        n = int(rows.shape[0])
        self._grow(self._n + n)
        self._F[self._n:self._n + n].copy_(rows.to(self.device))
        self._inv[self._n:self._n + n].copy_(inv.to(self.device))
        self._n += n

    def fde_scores(self, q_fde: torch.Tensor) -> torch.Tensor:
        """[n_q, fde_dim] float32 device -> cosine-ranking scores [n_q, ld] (ld >= n_pages)."""
        n_q = q_fde.shape[0]
        ld = max((self._n + 31) // 32 * 32, 32)
        scores = torch.empty((n_q, ld), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self.index.h.check(
                nat.lib.b200ms_fde_scan(self.index.h.ptr, _vp(self._F), _vp(self._inv), self._n, _vp(q_fde.contiguous()), n_q,
                                        _vp(scores), ld, self.index._stream()),
                "b200ms_fde_scan")
        return scores

    def compact(self, keep: Sequence[int]) -> None:
        """Keep only pages `keep` (ascending old ids) in both the patch corpus and the FDE matrix."""
        keep_t = torch.as_tensor([int(p) for p in keep], dtype=torch.int64, device=self.device)
        self.index.compact(keep)
        if self._F is not None and len(keep):
            self._F = self._F[: self._n].index_select(0, keep_t).contiguous()
            self._inv = self._inv[: self._n].index_select(0, keep_t).contiguous()
        self._n = len(keep)

    def candidates(self, queries: Sequence, n_candidates: int, allow_mask_dev: Optional[torch.Tensor] = None):
        """First stage only: (candidate page ids int64 [n_q, n_candidates] (-1 padded), FDE scores, counts) on the device.
        allow_mask_dev: optional uint32 page bitmask (the doc_ids filter, applied BEFORE candidate selection like the
        reference's Turbopuffer filter, fast_multivector_store.py:527)."""
        n_q = len(queries)
        q_fde = encode_items(self.index, list(queries), False, self.fde_dim)
        scores = self.fde_scores(q_fde)
        kk = int(min(n_candidates, nat.MAX_K))
        ts = torch.empty((n_q, kk), dtype=torch.float32, device=self.device)
        ti = torch.empty((n_q, kk), dtype=torch.int64, device=self.device)
        tc = torch.empty((n_q,), dtype=torch.int32, device=self.device)
        goff = nat.i32_array(list(range(n_q + 1)))
        with torch.cuda.device(self.device):
            self.index.h.check(
                nat.lib.b200ms_topk(self.index.h.ptr, _vp(scores), nat.F32, self._n, scores.shape[1], goff, n_q,
                                    _vp(allow_mask_dev), kk,
                                    ctypes.c_float(1.0), 0, _vp(ts), _vp(ti), _vp(tc), self.index._stream()),
                "b200ms_topk(fde)")
        return ti, ts, tc

    def rerank(self, query, cand_ids: torch.Tensor, k: int):
        """Exact MaxSim of ONE query over the candidate pages (device int64 [n_cand], -1 = unused) -> top-k."""
        self.index._attach()
        src, src_dtype = self.index._stage_rows([query])
        n_cand = int(cand_ids.numel())
        kk = int(min(k, n_cand))
        ts = torch.empty((1, kk), dtype=torch.float32, device=self.device)
        ti = torch.empty((1, kk), dtype=torch.int64, device=self.device)
        tc = torch.empty((1,), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            self.index.h.check(
                nat.lib.b200ms_rerank_device(self.index.h.ptr, _vp(src), src_dtype, nat.i32_array([int(query.shape[0])]), 1,
                                             _vp(cand_ids.contiguous()), n_cand, kk, ctypes.c_float(self.index.i8_scale),
                                             ctypes.c_float(self.index.score_scale), _vp(ts), _vp(ti), _vp(tc),
                                             self.index._stream()),
                "b200ms_rerank_device")
        return ts, ti, tc

    def search_device(self, queries: Sequence, k: int, n_candidates: int = 1000, allow_mask_dev: Optional[torch.Tensor] = None,
                      id_base: int = 0):
        """Two-stage search, everything on the device and asynchronous: query FDEs -> ONE scan of the FDE matrix for the
        whole batch (tcgen05) -> top-n_candidates per query -> ONE batched rerank call (each query against its own list)
        -> top-k.  Returns device tensors (scores [n_q,kk], page ids + id_base [n_q,kk], counts [n_q]) and the CUDA events
        that bracket the stages (for last_timing_ms)."""
        n_q = len(queries)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        q_fde = encode_items(self.index, list(queries), False, self.fde_dim)
        ev[1].record()
        scores = self.fde_scores(q_fde)
        kc = int(min(max(n_candidates, k), nat.MAX_K, max(self._n, 1)))
        cs = torch.empty((n_q, kc), dtype=torch.float32, device=self.device)
        ci = torch.empty((n_q, kc), dtype=torch.int64, device=self.device)
        cc = torch.empty((n_q,), dtype=torch.int32, device=self.device)
        goff = nat.i32_array(list(range(n_q + 1)))
        with torch.cuda.device(self.device):
            self.index.h.check(
                nat.lib.b200ms_topk(self.index.h.ptr, _vp(scores), nat.F32, self._n, scores.shape[1], goff, n_q,
                                    _vp(allow_mask_dev), kc, ctypes.c_float(1.0), 0, _vp(cs), _vp(ci), _vp(cc),
                                    self.index._stream()),
                "b200ms_topk(fde)")
        ev[2].record()
        src, _ = self.index._stage_rows(list(queries))
        ts, ti, tc = self.index.rerank_batch(src, [int(q.shape[0]) for q in queries], ci, k)
        if id_base:
            ti = torch.where(ti >= 0, ti + int(id_base), ti)
        ev[3].record()
        return ts, ti, tc, ev

    def search(self, queries: Sequence, k: int, n_candidates: int = 1000, allow_mask: Optional[np.ndarray] = None):
        """Two-stage search; returns host arrays (scores [n_q,k], page ids [n_q,k], counts [n_q]) and fills last_timing_ms
        with the reference's stage names (fast_multivector_store.py:523,534,550,574): encode_query / ns.query (here: FDE scan
        + candidate top-k on the GPU) / load_multivectors (0: resident in HBM) / rerank_scoring.
        allow_mask: optional uint32 words (page bitmask) restricting BOTH stages (the reference filters inside the ANN query)."""
        n_q = len(queries)
        mask_dev = None if allow_mask is None else torch.from_numpy(
            self.index._check_mask(allow_mask).view(np.int32)).to(self.device)
        ts, ti, tc, ev = self.search_device(queries, k, n_candidates, mask_dev)
        torch.cuda.synchronize(self.device)
        kk = ts.shape[1]
        out_s = np.full((n_q, k), -np.inf, dtype=np.float32)
        out_i = np.full((n_q, k), -1, dtype=np.int64)
        out_s[:, :kk], out_i[:, :kk] = ts.cpu().numpy(), ti.cpu().numpy()
        out_c = tc.cpu().numpy().astype(np.int32)
        self.last_timing_ms = {"encode_query_ms": ev[0].elapsed_time(ev[1]), "ns_query_ms": ev[1].elapsed_time(ev[2]),
                               "load_multivectors_ms": 0.0, "rerank_scoring_ms": ev[2].elapsed_time(ev[3]),
                               "fde_candidates_ms": ev[0].elapsed_time(ev[2]), "maxsim_rerank_ms": ev[2].elapsed_time(ev[3])}
        return out_s, out_i, out_c
