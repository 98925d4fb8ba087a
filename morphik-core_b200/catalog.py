"""PageCatalog -- host-side bookkeeping next to the device corpus (pure Python/numpy; no CUDA).

The device only knows integer page ids.  Everything the reference keeps in relational columns next to the embeddings
(``document_id``, ``chunk_number``, ``content``, ``chunk_metadata`` of ``multi_vector_embeddings``,
core/vector_store/multi_vector_store.py:240-251) lives here, together with:
  * the ``WHERE document_id IN (...)`` filter (multi_vector_store.py:752-757;
    ``filters=("document_id","In",doc_ids)`` fast_multivector_store.py:527) turned into a per-page allow mask,
  * ``DELETE FROM multi_vector_embeddings WHERE document_id = ...`` (multi_vector_store.py:929-933) as tombstones
    plus a compaction plan.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np


@dataclass
class PageRecord:
    document_id: str
    chunk_number: int
    content: str
    metadata: Dict[str, Any] = field(default_factory=dict)
    app_id: Optional[str] = None
    n_rows: int = 0


class PageCatalog:
    def __init__(self) -> None:
        self.records: List[PageRecord] = []
        self.alive = np.zeros(0, dtype=bool)
        self._n = 0
        self._by_doc: Dict[str, List[int]] = {}
        self._by_key: Dict[Tuple[str, int], int] = {}

    # ------------------------------------------------------------------ size
    def __len__(self) -> int:
        return self._n

    @property
    def n_live(self) -> int:
        return int(self.alive[: self._n].sum())

    @property
    def dead_fraction(self) -> float:
        return 0.0 if self._n == 0 else 1.0 - self.n_live / self._n

    # ------------------------------------------------------------------ writes
    def add(self, rec: PageRecord) -> int:
        pid = self._n
        if pid >= self.alive.shape[0]:
            grown = np.zeros(max(1024, 2 * self.alive.shape[0]), dtype=bool)
            grown[: self._n] = self.alive[: self._n]
            self.alive = grown
        self.records.append(rec)
        self.alive[pid] = True
        self._n += 1
        self._by_doc.setdefault(rec.document_id, []).append(pid)
        self._by_key[(rec.document_id, int(rec.chunk_number))] = pid  # a re-insert of the same key shadows the old page
        return pid

    def delete_document(self, document_id: str) -> List[int]:
        pids = self._by_doc.pop(document_id, [])
        for pid in pids:
            self.alive[pid] = False
            rec = self.records[pid]
            if self._by_key.get((rec.document_id, int(rec.chunk_number))) == pid:
                del self._by_key[(rec.document_id, int(rec.chunk_number))]
        return pids

    # ------------------------------------------------------------------ reads
    def lookup(self, document_id: str, chunk_number: int) -> Optional[int]:
        pid = self._by_key.get((document_id, int(chunk_number)))
        return pid if pid is not None and self.alive[pid] else None

    def pages_of(self, document_id: str) -> List[int]:
        return [p for p in self._by_doc.get(document_id, []) if self.alive[p]]

    def allow_mask(self, doc_ids: Optional[Sequence[str]] = None, app_id: Optional[str] = None) -> Optional[np.ndarray]:
        """bool[n_pages] of pages a query may return, or None when every page qualifies (no mask upload needed).

        doc_ids=None means "no document filter" (multi_vector_store.py:752); an EMPTY list selects nothing.
        A page stored with an app_id is only visible to queries with the same app_id (the fast store's namespace,
        fast_multivector_store.py:311,526); pages stored without one are visible to everybody.
        """
        n = self._n
        mask = self.alive[:n].copy()
        if doc_ids is not None:
            sel = np.zeros(n, dtype=bool)
            for d in dict.fromkeys(doc_ids):
                ids = self._by_doc.get(d)
                if ids:
                    sel[ids] = True
            mask &= sel
        if app_id is not None:
            foreign = np.fromiter((r.app_id is not None and r.app_id != app_id for r in self.records), dtype=bool, count=n)
            mask &= ~foreign
        return None if mask.all() else mask

    @staticmethod
    def mask_words(mask: np.ndarray) -> np.ndarray:
        """bool[n] -> uint32 words, bit p&31 of word p>>5 (the device filter format of b200ms_topk)."""
        pad = (-len(mask)) % 32
        bits = np.concatenate([mask, np.zeros(pad, dtype=bool)]) if pad else mask
        if len(bits) == 0:
            return np.zeros(0, dtype=np.uint32)
        return np.packbits(bits.reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).reshape(-1).copy()

    # ------------------------------------------------------------------ compaction
    def compaction_plan(self) -> Tuple[np.ndarray, np.ndarray]:
        """(kept old page ids in order, old->new id map with -1 for dropped pages)."""
        keep = np.nonzero(self.alive[: self._n])[0]
        remap = np.full(self._n, -1, dtype=np.int64)
        remap[keep] = np.arange(len(keep))
        return keep, remap

    def apply_compaction(self, keep: Iterable[int]) -> None:
        old = self.records
        self.records, self._by_doc, self._by_key = [], {}, {}
        self.alive = np.zeros(0, dtype=bool)
        self._n = 0
        for pid in keep:
            self.add(old[int(pid)])
