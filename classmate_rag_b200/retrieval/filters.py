"""Metadata filters of the two stores, evaluated on the device.

* ``build_where_filter`` turns CLI-style filters into the vector store's ``where`` dict
  (reference rag/retrieval/vector_chroma.py:22-78).
* ``chroma_clauses`` / ``bm25_clauses`` flatten a ``where`` into a conjunction of
  (column key, wanted value) pairs with the exact semantics of the two reference
  evaluators: Chroma-style equality (key must exist and be equal) and
  ``_matches_filter`` (rag/retrieval/bm25.py:79-107: ``meta.get(f) != where[f]`` on six
  fields, so a None-valued filter key only matches documents LACKING the field; a
  ``tags: {"$contains": ...}`` clause is evaluated alone).
* ``MetaColumns`` keeps one dictionary-coded int32 column per referenced key on the device
  and calls ``cmr_filter_mask``.
"""
from __future__ import annotations

import re
from typing import Any, Dict, Hashable, List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import ops

SIMPLE_FIELDS = ("course", "unit", "language", "doc_type", "author", "semester")


def _slug_tag(t: str) -> str:
    return re.sub(r"[^a-z0-9]+", "_", (t or "").lower().strip()).strip("_")


def _parse_tags(obj) -> List[str]:
    if not obj:
        return []
    vals = [str(x) for x in obj] if isinstance(obj, (list, tuple)) else str(obj).split(",")
    return [v.strip() for v in vals if v.strip()]


def build_where_filter(meta_like: Mapping[str, Any]) -> Optional[Dict[str, Any]]:
    """Equality on the simple fields (blank strings and the placeholder doc_type "other"
    are ignored), tags as boolean flags ``tag_<slug>: True``; one clause is returned bare,
    several under ``$and``; nothing to filter on gives None."""
    if not meta_like:
        return None
    clauses: List[Dict[str, Any]] = []
    for f in SIMPLE_FIELDS:
        v = meta_like.get(f)
        if v is None:
            continue
        if isinstance(v, str):
            v = v.strip()
            if not v or (f == "doc_type" and v.lower() == "other"):
                continue
        clauses.append({f: v})
    for t in _parse_tags(meta_like.get("tags")):
        slug = _slug_tag(t)
        if slug:
            clauses.append({f"tag_{slug}": True})
    if not clauses:
        return None
    return clauses[0] if len(clauses) == 1 else {"$and": clauses}


Clause = Tuple[Hashable, Any]   # (column key, wanted value); value None = "field absent"


class UnsupportedWhere(ValueError):
    pass


def chroma_clauses(where: Optional[Mapping[str, Any]]) -> List[Clause]:
    """Chroma ``where`` -> conjunction.  Supported: ``{key: value}``, ``{key: {"$eq": value}}``
    and ``{"$and": [...]}`` (everything build_where_filter emits)."""
    if not where:
        return []
    out: List[Clause] = []
    for key, val in where.items():
        if key == "$and":
            for c in val:
                out.extend(chroma_clauses(c))
        elif isinstance(key, str) and key.startswith("$"):
            raise UnsupportedWhere(f"where operator {key!r} is not supported")
        elif isinstance(val, Mapping):
            if set(val) != {"$eq"}:
                raise UnsupportedWhere(f"where operator {sorted(val)} on {key!r} is not supported")
            out.append((("meta", key), _Present(val["$eq"])))
        else:
            out.append((("meta", key), _Present(val)))
    return out


def bm25_clauses(where: Optional[Mapping[str, Any]]) -> List[Clause]:
    """BM25Store._matches_filter as a conjunction (same evaluation order and short cuts)."""
    if not where:
        return []
    if "$and" in where:
        out: List[Clause] = []
        for c in where["$and"]:
            out.extend(bm25_clauses(c))
        return out
    tg = where.get("tags") if "tags" in where else None
    if isinstance(tg, Mapping) and "$contains" in tg:
        want = tg["$contains"]
        if not want:
            return []
        want_set = {want} if isinstance(want, str) else set(want)
        return [(("tag", t), True) for t in sorted(want_set, key=repr)]
    return [(("get", f), where[f]) for f in SIMPLE_FIELDS if f in where]


class _Present:
    """Wanted value of a Chroma-style clause: the key must exist AND compare equal."""
    __slots__ = ("value",)

    def __init__(self, value):
        self.value = value


class MetaColumns:
    """Dictionary-coded metadata columns on the device, built lazily per referenced key.

    column key ("meta", k): code of metas[i][k], -1 when the key is missing (or None)
    column key ("get",  f): same coding; a wanted value of None asks for code -1
    column key ("tag",  t): 1 when t is in metas[i].get("tags"), else 0
    """

    def __init__(self, device):
        self.device = device
        self._metas: Sequence[Mapping[str, Any]] = ()
        self._keys: List[Hashable] = []
        self._dicts: Dict[Hashable, Dict[Any, int]] = {}
        self._host_cols: Dict[Hashable, np.ndarray] = {}
        self._dev: Optional[torch.Tensor] = None

    def reset(self, metas: Sequence[Mapping[str, Any]]) -> None:
        """Point at a new (or mutated) metadata list; columns are rebuilt on demand."""
        self._metas = metas
        self._keys, self._dicts, self._host_cols, self._dev = [], {}, {}, None

    def _code_rows(self, key: Hashable, lo: int, hi: int, codes: Dict[Any, int]) -> np.ndarray:
        """Codes of rows [lo, hi) for one column (grows ``codes`` with the values it meets)."""
        kind, name = key
        col = np.full(hi - lo, -1, dtype=np.int32)
        metas = self._metas
        if kind == "tag":
            for i in range(lo, hi):
                col[i - lo] = 1 if name in (metas[i].get("tags") or []) else 0
            codes.setdefault(True, 1)
        else:
            for i in range(lo, hi):
                v = metas[i].get(name)
                if v is None:
                    continue
                try:
                    c = codes.get(v)
                    if c is None:
                        c = codes[v] = len(codes)
                except TypeError:  # unhashable metadata value: can never equal a filter scalar
                    c = -3
                col[i - lo] = c
        return col

    def _ensure(self, key: Hashable) -> int:
        if key in self._host_cols:
            return self._keys.index(key)
        codes: Dict[Any, int] = {}
        col = self._code_rows(key, 0, len(self._metas), codes)
        self._keys.append(key)
        self._dicts[key] = codes
        self._host_cols[key] = col
        self._dev = None
        return len(self._keys) - 1

    def extend(self, metas: Sequence[Mapping[str, Any]], n_old: int) -> None:
        """Rows were APPENDED to the metadata list (rows [0, n_old) unchanged): every column built
        so far grows by the codes of the new rows only, so a store that is filled batch by batch
        never re-codes what it already has."""
        self._metas = metas
        n = len(metas)
        if n == n_old:
            return
        for key in self._keys:
            self._host_cols[key] = np.concatenate([self._host_cols[key][:n_old],
                                                   self._code_rows(key, n_old, n, self._dicts[key])])
        self._dev = None

    def prepare(self, keys: Sequence[Hashable]) -> None:
        """Build the columns of these keys now (index build time) instead of at their first use."""
        for key in keys:
            self._ensure(key)

    def mask(self, clauses: Sequence[Clause], alive: Optional[torch.Tensor] = None) -> torch.Tensor:
        """uint8 [n] device mask of the rows satisfying every clause (and alive)."""
        fields, codes = [], []
        for key, want in clauses:
            f = self._ensure(key)
            if isinstance(want, _Present):
                want = want.value
                code = -2 if want is None else self._dicts[key].get(want, -2)
            elif want is None:
                code = -1
            else:
                try:
                    code = self._dicts[key].get(want, -2)
                except TypeError:
                    code = -2
            fields.append(f)
            codes.append(code)
        n = len(self._metas)
        if self._dev is None or self._dev.shape[0] != len(self._keys):
            host = (np.stack([self._host_cols[k] for k in self._keys]) if self._keys
                    else np.zeros((1, n), dtype=np.int32))
            self._dev = torch.from_numpy(np.ascontiguousarray(host)).to(self.device)
        cf = torch.tensor(fields, dtype=torch.int32, device=self.device)
        cc = torch.tensor(codes, dtype=torch.int32, device=self.device)
        return ops.filter_mask(self._dev, cf, cc, alive=alive)
