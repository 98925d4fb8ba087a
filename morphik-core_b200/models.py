"""Record types of the plugin surface -- same fields as the reference's core/models/chunk.py:9-38.

When the package runs inside a Morphik checkout the reference's own classes are used (so isinstance checks in
callers keep working); standalone, structurally identical pydantic models are defined here.
"""
from __future__ import annotations

from typing import Any, Dict, List, Union

import numpy as np

try:  # inside a Morphik process: reuse the host application's model classes
    from core.models.chunk import Chunk, DocumentChunk  # type: ignore  # noqa: F401
except Exception:  # standalone
    from pydantic import BaseModel, Field

    Embedding = Union[List[float], List[List[float]], np.ndarray]

    class DocumentChunk(BaseModel):
        """A chunk stored in (or returned by) a vector store: core/models/chunk.py:9-20."""

        document_id: str
        content: str
        embedding: Any  # List[float] | List[List[float]] | np.ndarray ([P,128] for ColPali pages)
        chunk_number: int
        metadata: Dict[str, Any] = Field(default_factory=dict)
        score: float = 0.0

        model_config = {"arbitrary_types_allowed": True}

    class Chunk(BaseModel):
        """core/models/chunk.py:23-38."""

        content: str
        metadata: Dict[str, Any] = Field(default_factory=dict)

        model_config = {"arbitrary_types_allowed": True}

        def to_document_chunk(self, document_id: str, chunk_number: int, embedding: Any) -> DocumentChunk:
            return DocumentChunk(document_id=document_id, content=self.content, embedding=embedding,
                                 chunk_number=chunk_number, metadata=self.metadata)
