"""Exact dense vector store on one B200: the drop-in for the reference's Chroma wrapper
(rag/retrieval/vector_chroma.py:81-278).

Same class name, constructor, method names, keyword-only arguments, result dict shapes
and error behaviour; the HNSW graph walk behind ``collection.query`` is replaced by the
exact bf16 scan / tcgen05 GEMM of libcmrag (``cmr_dense_topk``), so results are the true
top-k, not an approximation.  Rows live in HBM as one bf16 ``[rows, dim]`` matrix in
insertion order; ids, documents and metadata stay on the host.

distance = 1 - q.c ("cosine" space on unit vectors: the reference always stores and
queries L2-normalised E5 output, rag/embeddings/__init__.py:85-105).  Like hnswlib's cosine
space, rows and queries that are NOT unit length (norm off by more than 1e-3, e.g. an embedder
run with normalize=False) are L2-normalised on the way in; unit input is left bit for bit as
it is, so the bf16 matrix equals the oracle's.  Ties are broken by insertion order.

Durability follows the reference's ``chromadb.PersistentClient``: every upsert / delete is
appended to ``<persist_dir>/<collection>.cmrag/`` (raw bf16 rows + one JSON line per
operation) before the call returns, so a later process finds what an earlier one ingested
without anyone calling ``persist()`` (rag/pipeline/rag.py:411-413 never does).  ``persist()``
/ ``compact()`` rewrite the directory without tombstones, atomically (temp dir + rename).
"""
from __future__ import annotations

import json
import os
import shutil
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Dict, List, Mapping, Optional, Sequence

import numpy as np
import torch

from .. import ops
from .filters import SIMPLE_FIELDS, MetaColumns, chroma_clauses
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
        self.max_row_norm = 1.0       # largest |row| of the bf16 matrix (certificate bound)
        self.log_dir: Optional[Path] = None   # append-only operation log (None: in-memory collection)
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

    def delete(self, ids: Sequence[str], log: bool = True) -> int:
        gone = [i for i in ids if i in self.row_of]
        rows = [self.row_of.pop(i) for i in gone]
        if rows:
            self.alive[torch.tensor(rows, dtype=torch.int64, device=self.device)] = 0
            self.n_dead += len(rows)
            self.version += 1
            if log and self.log_dir is not None:
                _log_append(self.log_dir, self.dim, None, [{"delete": gone}])
        return len(rows)

    def add(self, ids, documents, metadatas, emb_f32: np.ndarray, log: bool = True, bits: Optional[torch.Tensor] = None) -> None:
        n = len(ids)
        if n == 0:
            return
        lo = self.n_rows
        if bits is None:
            self._reserve(n, int(emb_f32.shape[1]))
            x = unit_rows(torch.from_numpy(np.ascontiguousarray(emb_f32, dtype=np.float32)).to(self.device))
            bits = ops.f32_to_bf16(x)
        else:
            self._reserve(n, int(bits.shape[1]))
        self.emb[lo:lo + n] = bits
        self.max_row_norm = max(self.max_row_norm, float(bits.float().norm(dim=1).max()))
        if log and self.log_dir is not None:
            _log_append(self.log_dir, self.dim, bits,
                        [{"id": i, "document": d, "metadata": dict(m or {})} for i, d, m in zip(ids, documents, metadatas)])
        self.alive[lo:lo + n] = 1
        self.gids[lo:lo + n] = torch.tensor([REGISTRY.intern(i) for i in ids], dtype=torch.int64, device=self.device)
        for j, cid in enumerate(ids):
            self.row_of[cid] = lo + j
        self.ids.extend(ids)
        self.documents.extend(documents)
        n_meta_old = len(self.metadatas)
        self.metadatas.extend(dict(m or {}) for m in metadatas)
        self.n_rows += n
        if n_meta_old == lo and self.columns._metas is self.metadatas:
            self.columns.extend(self.metadatas, n_meta_old)      # appended rows only
        else:
            self.columns.reset(self.metadatas)
        self.columns.prepare([("meta", f) for f in SIMPLE_FIELDS])   # the reference's filterable fields are coded
        # at upsert time (Chroma indexes metadata at upsert too): the first filtered query pays no O(rows) loop
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

    def cert_eps(self, q_bf16: torch.Tensor) -> float:
        """Bound on |fp32 tensor-pipe score - exact| for these queries against this matrix."""
        return ops.dense_cert_eps(self.dim, float(q_bf16.float().norm(dim=-1).max()), self.max_row_norm)

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
                snap = self._snap_dir()
                if snap.exists():
                    _load_snapshot(col, snap)
                col.log_dir = snap
            self._collection = col
        return self._collection

    def _snap_dir(self) -> Path:
        return Path(self.persist_dir) / f"{self.collection_name}.cmrag"

    def reset_collection(self) -> None:
        _COLLECTIONS.pop(self._key(), None)
        self._collection = None
        shutil.rmtree(self._snap_dir(), ignore_errors=True)
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
        """Drop tombstoned rows from the matrix and from the directory on disk."""
        self.persist()

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
        q = ops.f32_to_bf16(unit_rows(torch.from_numpy(np.ascontiguousarray(q_f32)).to(col.device)))
        mask = col.mask(where)
        k = min(int(top_k), ops_max_k())
        if k <= 0:
            raise ValueError("top_k must be positive")
        # flagged (uncertifiable) queries are re-run on the exhaustive float64 scan
        out = ops.dense_topk_certified(col.matrix(), q, k, row_mask=mask, workspace=col.workspace(q.shape[0], k),
                                       algo=col.algo_for(mask), cert_eps=col.cert_eps(q))
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

    # ---- persistence (the Chroma directory of the reference becomes a binary operation log) -----
    def persist(self) -> Path:
        """Rewrite the collection's directory without tombstones: a temp directory is written in
        full and then renamed over the old one, so a crash leaves either the old or the new state."""
        col = self._ensure_collection()
        col.compact()
        snap = self._snap_dir()
        tmp = snap.with_name(snap.name + ".tmp")
        old = snap.with_name(snap.name + ".old")
        shutil.rmtree(tmp, ignore_errors=True)
        shutil.rmtree(old, ignore_errors=True)
        if col.n_rows:
            _log_append(tmp, col.dim, col.matrix().contiguous(),
                        [{"id": cid, "document": doc, "metadata": meta}
                         for cid, doc, meta in zip(col.ids, col.documents, col.metadatas)])
        else:
            tmp.mkdir(parents=True, exist_ok=True)
        if snap.exists():
            os.replace(snap, old)
        os.replace(tmp, snap)
        shutil.rmtree(old, ignore_errors=True)
        return snap


def unit_rows(x: torch.Tensor) -> torch.Tensor:
    """hnswlib's cosine space normalises what it is given.  Rows whose norm is off by more than
    1e-3 are divided by it; unit rows (the E5 contract) pass through untouched, zero rows stay."""
    norms = x.norm(dim=1, keepdim=True)
    off = ((norms - 1.0).abs() > 1e-3) & (norms > 0)
    if bool(off.any()):
        x = torch.where(off, x / norms.clamp(min=1e-30), x)
    return x


def _log_append(snap: Path, dim: int, bits: Optional[torch.Tensor], records: List[Dict[str, Any]]) -> None:
    """Append rows (raw little-endian bf16, [n, dim]) and their records to the operation log;
    the rows are written first, so a record line always has its row."""
    snap.mkdir(parents=True, exist_ok=True)
    meta = snap / "meta.json"
    if not meta.exists():
        meta.write_text(json.dumps({"version": 2, "dim": int(dim)}))
    if bits is not None:
        with (snap / "rows.bf16").open("ab") as f:
            f.write(bits.contiguous().view(torch.int16).cpu().numpy().tobytes())
    with (snap / "records.jsonl").open("a", encoding="utf-8") as f:
        for rec in records:
            f.write(json.dumps(rec, ensure_ascii=False) + "\n")


def ops_max_k() -> int:
    from .. import _lib
    return _lib.CMR_MAX_K


def _load_snapshot(col: _Collection, snap: Path) -> None:
    """Replay the operation log (version 2) or read a round-1 snapshot (embeddings_bf16.npy)."""
    ops_log: List[Dict[str, Any]] = []
    rec_path = snap / "records.jsonl"
    if rec_path.exists():
        with rec_path.open("r", encoding="utf-8") as f:
            for line in f:
                if line.strip():
                    try:
                        ops_log.append(json.loads(line))
                    except json.JSONDecodeError:
                        break       # a torn last line: everything before it is intact
    if (snap / "meta.json").exists():
        dim = int(json.loads((snap / "meta.json").read_text())["dim"])
        raw = np.fromfile(snap / "rows.bf16", dtype=np.uint16) if (snap / "rows.bf16").exists() else np.zeros(0, np.uint16)
        bits = raw[: raw.size // max(dim, 1) * max(dim, 1)].reshape(-1, max(dim, 1))
    elif (snap / "embeddings_bf16.npy").exists():
        bits = np.load(snap / "embeddings_bf16.npy")
    else:
        return
    n_adds = sum(1 for r in ops_log if "delete" not in r)
    if n_adds > bits.shape[0]:
        raise RuntimeError(f"corrupt snapshot {snap}: {n_adds} records, {bits.shape[0]} rows")
    row, i = 0, 0
    while i < len(ops_log):   # runs of adds go to the device in one copy
        if "delete" in ops_log[i]:
            col.delete(list(ops_log[i]["delete"]), log=False)
            i += 1
            continue
        j = i
        while j < len(ops_log) and "delete" not in ops_log[j]:
            j += 1
        run = ops_log[i:j]
        ids = [r["id"] for r in run]
        col.delete(ids, log=False)   # delete-then-add, as upsert does
        chunk = torch.from_numpy(bits[row:row + len(run)].view(np.int16).copy()).to(col.device).view(torch.bfloat16)
        col.add(ids, [r.get("document") for r in run], [r.get("metadata") or {} for r in run], None, log=False, bits=chunk)
        row += len(run)
        i = j
