"""Chunk ids: the deterministic string ids of the reference (rag/utils/ids.py:17-29) and
their integer image shared by the stores (the kernels fuse and sort on int64 ids)."""
from __future__ import annotations

from hashlib import blake2b
from pathlib import Path
from typing import Dict, List, Optional, Union


def stable_chunk_id(*, source_path: Union[str, Path], page: int, chunk_index: int, course: Optional[str] = None,
                    unit: Optional[str] = None, prefix: str = "cm_") -> str:
    """prefix + blake2b-128 hex of "resolved path|page|chunk|course|unit" (empty string for
    missing course / unit), so that re-ingesting a file is idempotent."""
    resolved = str(Path(source_path).resolve())
    key = "|".join((resolved, str(page), str(chunk_index), course or "", unit or ""))
    return prefix + blake2b(key.encode("utf-8"), digest_size=16).hexdigest()


class IdRegistry:
    """str id <-> dense int64, process wide: a chunk has the same integer in the vector
    store and in the lexical store, whatever their insertion orders."""

    def __init__(self):
        self._num: Dict[str, int] = {}
        self._name: List[str] = []

    def intern(self, chunk_id: str) -> int:
        n = self._num.get(chunk_id)
        if n is None:
            n = self._num[chunk_id] = len(self._name)
            self._name.append(chunk_id)
        return n

    def name(self, num: int) -> str:
        return self._name[num]

    def __len__(self) -> int:
        return len(self._name)


REGISTRY = IdRegistry()
