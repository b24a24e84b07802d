"""Exact dense vector store on one B200: the drop-in for the reference's Chroma wrapper
(rag/retrieval/vector_chroma.py:81-278).

Same class name, constructor, method names, keyword-only arguments, result dict shapes
and error behaviour; the HNSW graph walk behind ``collection.query`` is replaced by the
exact bf16 scan / tcgen05 GEMM of libcmrag (``cmr_dense_topk``), so results are the true
top-k, not an approximation.  Rows live in HBM as one bf16 ``[rows, dim]`` matrix in
insertion order; ids, documents and metadata stay on the host.

distance = 1 - q.c ("cosine" space on unit vectors: the reference always stores and
queries L2-normalised E5 output, rag/embeddings/__init__.py:85-105).  Ties are broken by
insertion order.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Dict, List, Mapping, Optional, Sequence

import numpy as np
import torch

from .. import ops
from .filters import MetaColumns, chroma_clauses
from .ids import REGISTRY

_COLLECTIONS: Dict[tuple, "_Collection"] = {}   # (persist_dir, name) -> live collection of this process


class _Collection:
    """Rows of one collection: host-side records + the device matrix."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.ids: List[str] = []
        self.documents: List[Optional[str]] = []
        self.metadatas: List[Dict[str, Any]] = []
        self.row_of: Dict[str, int] = {}
        self.dim = 0
        self.emb: Optional[torch.Tensor] = None       # bf16 [capacity, dim]
        self.alive: Optional[torch.Tensor] = None     # uint8 [capacity]
        self.gids: Optional[torch.Tensor] = None      # int64 [capacity] registry number of every row
        self.n_rows = 0
        self.n_dead = 0
        self.columns = MetaColumns(self.device)
        self.version = 0
        self._ws: Dict[tuple, ops.DenseWorkspace] = {}

    def workspace(self, n_queries: int, k: int) -> "ops.DenseWorkspace":
        """Reusable scratch + output buffers of cmr_dense_topk for this shape."""
        key = (self.n_rows, self.dim, n_queries, k)
        ws = self._ws.get(key)
        if ws is None:
            if len(self._ws) > 16:
                self._ws.clear()
            ws = self._ws[key] = ops.DenseWorkspace(self.n_rows, self.dim, n_queries, k, self.device)
        return ws

    def _reserve(self, extra: int, dim: int) -> None:
        if self.emb is None:
            self.dim = dim
            cap = max(1024, extra)
            self.emb = torch.zeros((cap, dim), dtype=torch.bfloat16, device=self.device)
            self.alive = torch.zeros((cap,), dtype=torch.uint8, device=self.device)
            self.gids = torch.full((cap,), -1, dtype=torch.int64, device=self.device)
            return
        if dim != self.dim:
            raise ValueError(f"embedding dimension {dim} does not match the collection's {self.dim}")
        need = self.n_rows + extra
        cap = self.emb.shape[0]
        if need > cap:
            cap = max(need, 2 * cap)
            for name, fill in (("emb", 0), ("alive", 0), ("gids", -1)):
                old = getattr(self, name)
                new = torch.full((cap, *old.shape[1:]), fill, dtype=old.dtype, device=self.device)
                new[: self.n_rows] = old[: self.n_rows]
                setattr(self, name, new)

    def delete(self, ids: Sequence[str]) -> int:
        rows = [self.row_of.pop(i) for i in ids if i in self.row_of]
        if rows:
            self.alive[torch.tensor(rows, dtype=torch.int64, device=self.device)] = 0
            self.n_dead += len(rows)
            self.version += 1
        return len(rows)

    def add(self, ids, documents, metadatas, emb_f32: np.ndarray) -> None:
        n = len(ids)
        if n == 0:
            return
        self._reserve(n, int(emb_f32.shape[1]))
        lo = self.n_rows
        x = torch.from_numpy(np.ascontiguousarray(emb_f32, dtype=np.float32)).to(self.device)
        self.emb[lo:lo + n] = ops.f32_to_bf16(x)
        self.alive[lo:lo + n] = 1
        self.gids[lo:lo + n] = torch.tensor([REGISTRY.intern(i) for i in ids], dtype=torch.int64, device=self.device)
        for j, cid in enumerate(ids):
            self.row_of[cid] = lo + j
        self.ids.extend(ids)
        self.documents.extend(documents)
        self.metadatas.extend(dict(m or {}) for m in metadatas)
        self.n_rows += n
        self.columns.reset(self.metadatas)
        self.version += 1

    def compact(self) -> None:
        """Drop tombstoned rows (keeps insertion order)."""
        if self.n_dead == 0:
            return
        keep = torch.nonzero(self.alive[: self.n_rows]).flatten()
        keep_h = keep.cpu().tolist()
        self.emb[: len(keep_h)] = self.emb[keep]
        self.gids[: len(keep_h)] = self.gids[keep]
        self.alive[: len(keep_h)] = 1
        self.alive[len(keep_h): self.n_rows] = 0
        self.ids = [self.ids[r] for r in keep_h]
        self.documents = [self.documents[r] for r in keep_h]
        self.metadatas = [self.metadatas[r] for r in keep_h]
        self.row_of = {cid: r for r, cid in enumerate(self.ids)}
        self.n_rows, self.n_dead = len(keep_h), 0
        self.columns.reset(self.metadatas)
        self.version += 1

    def matrix(self) -> torch.Tensor:
        return self.emb[: self.n_rows] if self.emb is not None else torch.zeros(
            (0, max(self.dim, 8)), dtype=torch.bfloat16, device=self.device)

    def mask(self, where: Optional[Mapping[str, Any]]) -> Optional[torch.Tensor]:
        """Device row mask of a query: None when every row qualifies (no filter, no
        tombstones)."""
        clauses = chroma_clauses(where)
        if not clauses and self.n_dead == 0:
            return None
        return self.columns.mask(clauses, alive=self.alive[: self.n_rows])

    def algo_for(self, mask: Optional[torch.Tensor]) -> str:
        """A very selective filter is served best by the exhaustive float64 scan, which only
        reads the allowed rows; otherwise the library picks scan / tcgen05 by batch size."""
        if mask is None:
            return "auto"
        allowed = int(mask.sum())
        return "exact" if allowed * 16 <= self.n_rows else "auto"


@dataclass
class ChromaVectorStore:
    persist_dir: Path
    collection_name: str = "classmate_rag"
    distance: str = "cosine"
    device: str = "cuda"

    _collection: Optional[_Collection] = field(default=None, repr=False)

    # ---- collection life cycle ----------------------------------------------------------
    def _key(self):
        return (str(Path(self.persist_dir)), self.collection_name)

    def _ensure_collection(self) -> _Collection:
        if self._collection is None:
            if self.distance != "cosine":
                raise ValueError("only the 'cosine' space of the reference is implemented")
            if not torch.cuda.is_available():
                raise RuntimeError("ChromaVectorStore needs a CUDA device: classmate_rag_b200 has no CPU path")
            col = _COLLECTIONS.get(self._key())
            if col is None:
                col = _COLLECTIONS[self._key()] = _Collection(self.device)
                snap = Path(self.persist_dir) / f"{self.collection_name}.cmrag"
                if snap.exists():
                    _load_snapshot(col, snap)
            self._collection = col
        return self._collection

    def reset_collection(self) -> None:
        _COLLECTIONS.pop(self._key(), None)
        self._collection = None
        snap = Path(self.persist_dir) / f"{self.collection_name}.cmrag"
        if snap.exists():
            for f in snap.iterdir():
                f.unlink()
            snap.rmdir()
        self._ensure_collection()

    @classmethod
    def from_config(cls) -> "ChromaVectorStore":
        """Same environment variables and defaults as the reference's load_config()."""
        return cls(persist_dir=Path(os.getenv("CHROMA_PERSIST_DIRECTORY") or "./indexes/chroma"),
                   collection_name=os.getenv("CHROMA_COLLECTION_NAME") or "classmate_rag", distance="cosine")

    # ---- upsert -----------------------------------------------------------------------------
    def upsert(self, *, ids: Sequence[str], documents: Sequence[str], metadatas: Sequence[Mapping[str, Any]],
               embeddings: np.ndarray, batch_size: int = 512) -> None:
        """Delete-then-add, as the reference does: a re-upserted id moves to the end of the
        insertion order.  ``batch_size`` is accepted for compatibility (one device copy)."""
        if len(ids) != len(documents) or len(ids) != len(metadatas) or len(ids) != len(embeddings):
            raise ValueError("Lengths of ids, documents, metadatas, and embeddings must match.")
        col = self._ensure_collection()
        emb = np.asarray(embeddings, dtype=np.float32)
        if emb.ndim != 2 and len(ids):
            raise ValueError("embeddings must be a [n, dim] array")
        if len(ids) and emb.shape[1] % 8 != 0:
            raise ValueError("embedding dimension must be a multiple of 8")
        col.delete(list(ids))
        # the last occurrence of a repeated id wins, at its position
        last = {cid: j for j, cid in enumerate(ids)}
        sel = sorted(last.values())
        if len(sel) != len(ids):
            ids = [ids[j] for j in sel]
            documents = [documents[j] for j in sel]
            metadatas = [metadatas[j] for j in sel]
            emb = emb[sel]
        col.add(list(ids), list(documents), list(metadatas), emb)

    def delete(self, ids: Sequence[str]) -> int:
        """Missing from the reference wrapper but attempted by its callers
        (rag/admin/manage.py:190); returns the number of rows removed."""
        return self._ensure_collection().delete(list(ids))

    def compact(self) -> None:
        self._ensure_collection().compact()

    def dedupe(self, threshold: float = 0.95) -> List[str]:
        """Near-duplicate filter over the stored embeddings (the `rag rebuild` hook,
        rag/admin/backup.py:226-233; keep-first rule of rag/utils/dedup.py:40-55 on cosine):
        a row is dropped iff an earlier KEPT row has q.c >= threshold.  Runs on the tensor
        cores (classmate_rag_b200.neardup).  Returns the ids removed, in insertion order."""
        from .. import neardup
        col = self._ensure_collection()
        col.compact()
        if col.n_rows < 2:
            return []
        keep = neardup.neardup_keep_mask(col.matrix().contiguous(), threshold).bool().cpu().numpy()
        dropped = [cid for cid, k in zip(col.ids, keep) if not k]
        col.delete(dropped)
        col.compact()
        return dropped

    def count(self) -> int:
        col = self._ensure_collection()
        return col.n_rows - col.n_dead

    # ---- query ----------------------------------------------------------------------------
    def _search(self, q_f32: np.ndarray, where, top_k: int):
        col = self._ensure_collection()
        if col.n_rows == 0:
            return col, None
        if q_f32.shape[1] != col.dim:
            raise ValueError(f"query dimension {q_f32.shape[1]} does not match the collection's {col.dim}")
        q = ops.f32_to_bf16(torch.from_numpy(np.ascontiguousarray(q_f32)).to(col.device))
        mask = col.mask(where)
        k = min(int(top_k), ops_max_k())
        if k <= 0:
            raise ValueError("top_k must be positive")
        # flagged (uncertifiable) queries are re-run on the exhaustive float64 scan
        out = ops.dense_topk_certified(col.matrix(), q, k, row_mask=mask, workspace=col.workspace(q.shape[0], k),
                                       algo=col.algo_for(mask))
        return col, out

    def query(self, *, query_embeddings: np.ndarray, where: Optional[Dict[str, Any]] = None, top_k: int = 8,
              include_documents: bool = True, include_embeddings: bool = False) -> List[Dict[str, Any]]:
        """Hits of the FIRST query only, like the reference (vector_chroma.py:236-240)."""
        q = np.asarray(query_embeddings).astype("float32")
        if q.ndim == 1:
            q = q[None, :]
        res = self.query_batch(query_embeddings=q[:1], where=where, top_k=top_k,
                               include_documents=include_documents, include_embeddings=include_embeddings)
        return res[0] if res else []

    def query_batch(self, *, query_embeddings: np.ndarray, where: Optional[Dict[str, Any]] = None, top_k: int = 8,
                    include_documents: bool = True, include_embeddings: bool = False) -> List[List[Dict[str, Any]]]:
        """Extension: every query of a [B, dim] batch in one pass over the matrix."""
        q = np.asarray(query_embeddings).astype("float32")
        if q.ndim == 1:
            q = q[None, :]
        col, out = self._search(q, where, top_k)
        if out is None:
            return [[] for _ in range(q.shape[0])]
        scores, rows, counts, _flags = out
        rows_h, scores_h, counts_h = rows.cpu().numpy(), scores.cpu().numpy(), counts.cpu().numpy()
        embs = None
        if include_embeddings:
            embs = ops.gather_rows(col.matrix(), rows).float().cpu().numpy()
        result = []
        for b in range(q.shape[0]):
            items = []
            for j in range(int(counts_h[b])):
                r = int(rows_h[b, j])
                item = {"id": col.ids[r], "document": col.documents[r] if include_documents else None,
                        "metadata": col.metadatas[r], "distance": 1.0 - float(scores_h[b, j])}
                if embs is not None:
                    item["embedding"] = embs[b, j].copy()
                items.append(item)
            result.append(items)
        return result

    # ---- persistence (the Chroma directory of the reference becomes a binary snapshot) -----
    def persist(self) -> Path:
        col = self._ensure_collection()
        col.compact()
        snap = Path(self.persist_dir) / f"{self.collection_name}.cmrag"
        snap.mkdir(parents=True, exist_ok=True)
        bits = col.matrix().contiguous().view(torch.int16).cpu().numpy().view(np.uint16)
        np.save(snap / "embeddings_bf16.npy", bits)
        with (snap / "records.jsonl").open("w", encoding="utf-8") as f:
            for cid, doc, meta in zip(col.ids, col.documents, col.metadatas):
                f.write(json.dumps({"id": cid, "document": doc, "metadata": meta}, ensure_ascii=False) + "\n")
        return snap


def ops_max_k() -> int:
    from .. import _lib
    return _lib.CMR_MAX_K


def _load_snapshot(col: _Collection, snap: Path) -> None:
    bits = np.load(snap / "embeddings_bf16.npy")
    ids, docs, metas = [], [], []
    with (snap / "records.jsonl").open("r", encoding="utf-8") as f:
        for line in f:
            if line.strip():
                rec = json.loads(line)
                ids.append(rec["id"])
                docs.append(rec.get("document"))
                metas.append(rec.get("metadata") or {})
    if len(ids) != bits.shape[0]:
        raise RuntimeError(f"corrupt snapshot {snap}: {len(ids)} records, {bits.shape[0]} rows")
    if not ids:
        return
    col._reserve(len(ids), int(bits.shape[1]))
    col.emb[: len(ids)] = torch.from_numpy(bits.view(np.int16)).to(col.device).view(torch.bfloat16)
    col.alive[: len(ids)] = 1
    col.gids[: len(ids)] = torch.tensor([REGISTRY.intern(i) for i in ids], dtype=torch.int64, device=col.device)
    col.ids, col.documents, col.metadatas = ids, docs, metas
    col.row_of = {cid: r for r, cid in enumerate(ids)}
    col.n_rows = len(ids)
    col.columns.reset(col.metadatas)
    col.version += 1
