"""Hybrid retrieval: Reciprocal Rank Fusion + optional MMR, the drop-in for the reference's
rag/retrieval/fusion.py (rrf_fuse :17-36, _mmr_order :39-61, HybridRetriever :64-167).

Every numeric step runs in libcmrag on the device: dense top-k pool (``cmr_dense_topk``),
MMR re-ordering (``cmr_gather_rows`` + ``cmr_mmr_select``), BM25 (``cmr_bm25_topk``), RRF +
per-id merge + the final stable sort (``cmr_hybrid_fuse``).  The host only translates
between string ids and their integer image and assembles the result dicts.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, Dict, List, Mapping, Optional, Sequence

import numpy as np
import torch

from .. import ops
from .bm25_store import BM25Store
from .filters import build_where_filter
from .ids import REGISTRY
from .vector_store import ChromaVectorStore


def rrf_fuse(*, rank_lists: Sequence[Sequence[str]], weights: Optional[Sequence[float]] = None,
             rrf_k: int = 60) -> Dict[str, float]:
    """score[id] = sum over lists of w * 1/(rrf_k + rank), rank 1-based, float64, keyed in
    first-appearance order (``cmr_rrf_fuse``)."""
    if not rank_lists:
        return {}
    n = len(rank_lists)
    if weights is None:
        weights = [1.0] * n
    elif len(weights) != n:
        raise ValueError("weights length must match rank_lists length")
    max_len = max((len(l) for l in rank_lists), default=0)
    if max_len == 0:
        return {}
    if not torch.cuda.is_available():
        raise RuntimeError("rrf_fuse runs on the device: classmate_rag_b200 has no CPU path")
    local: Dict[Any, int] = {}
    names: List[Any] = []
    ids = np.full((n, max_len), -1, dtype=np.int64)
    counts = np.zeros(n, dtype=np.int32)
    for li, lst in enumerate(rank_lists):
        counts[li] = len(lst)
        for r, _id in enumerate(lst):
            num = local.get(_id)
            if num is None:
                num = local[_id] = len(names)
                names.append(_id)
            ids[li, r] = num
    dev = torch.device("cuda", torch.cuda.current_device())
    out_ids, out_scores, out_count = ops.rrf_fuse_lists(
        torch.from_numpy(ids).to(dev), torch.from_numpy(counts).to(dev),
        torch.tensor([float(w) for w in weights], dtype=torch.float64, device=dev), rrf_k)
    c = int(out_count.item())
    oi, osc = out_ids[:c].cpu().numpy(), out_scores[:c].cpu().numpy()
    return {names[int(i)]: float(s) for i, s in zip(oi, osc)}


@dataclass
class HybridRetriever:
    vector_store: ChromaVectorStore
    bm25_store: BM25Store
    embedder: Any   # .encode_queries(list[str]) -> float32 [B, dim], unit rows (E5 contract)

    k_vector: int = 8
    k_bm25: int = 8
    rrf_k: int = 60
    weight_vector: float = 1.0
    weight_bm25: float = 1.0

    use_mmr: bool = True
    mmr_lambda: float = 0.5
    mmr_max_pool: int = 24

    # -- the two searches, device level ---------------------------------------------------
    def _stage_query(self, q_vec: np.ndarray, dev) -> torch.Tensor:
        """fp32 query -> device through a pinned staging buffer kept on the retriever (no pageable copy)."""
        q = np.ascontiguousarray(q_vec, dtype=np.float32).reshape(1, -1)
        pin = getattr(self, "_pin_q", None)
        if pin is None or pin.shape != q.shape:
            pin = self._pin_q = torch.empty(q.shape, dtype=torch.float32).pin_memory()
        pin.copy_(torch.from_numpy(q))
        return pin.to(dev, non_blocking=True)

    def _vector_search_device(self, q_vec: np.ndarray, where, k: int, certified: bool = True):
        """(gids i64 [1,k'], sims f64 [1,k'], counts i32 [1], flags i32 [1]) in final (post-MMR)
        order.  ``certified``: re-run an uncertified query on the exhaustive scan right here (one
        host synchronisation); retrieve() defers that check to its single read-back instead."""
        from .vector_store import unit_rows
        col = self.vector_store._ensure_collection()
        pool = max(k, self.mmr_max_pool) if self.use_mmr else k
        pool = min(pool, 64)
        dev = col.device
        if col.n_rows == 0:
            return (torch.full((1, max(k, 1)), -1, dtype=torch.int64, device=dev),
                    torch.zeros((1, max(k, 1)), dtype=torch.float64, device=dev),
                    torch.zeros((1,), dtype=torch.int32, device=dev), torch.zeros((1,), dtype=torch.int32, device=dev))
        q = ops.f32_to_bf16(unit_rows(self._stage_query(q_vec, dev)))
        mask = col.mask(where)
        fn = ops.dense_topk_certified if certified else ops.dense_topk
        scores, rows, counts, flags = fn(col.matrix(), q, pool, row_mask=mask, workspace=col.workspace(1, pool),
                                         algo=col.algo_for(mask), cert_eps=ops.dense_cert_eps(col.dim, 1.0005, col.max_row_norm))
        if self.use_mmr:
            cand = ops.gather_rows(col.matrix(), rows)
            rows, scores, counts = ops.mmr_select(cand, scores, rows, counts, min(k, pool), self.mmr_lambda)
        else:
            rows, scores = rows[:, :k], scores[:, :k]
            counts = torch.clamp(counts, max=k)
        gids = torch.where(rows >= 0, col.gids[rows.clamp(min=0)], rows)
        return gids.contiguous(), scores.contiguous(), counts.contiguous(), flags

    def _bm25_search_device(self, query: str, where, k: int):
        if not query.strip() or self.bm25_store.count() == 0:
            return None
        got = self.bm25_store.search_device([query], where, k)
        if got is None:
            return None
        ix, sc, docs, cnt = got
        store_rows = docs.clamp(min=0) if ix.rows_dev is None else ix.rows_dev[docs.clamp(min=0)]
        gids = torch.where(docs >= 0, self.bm25_store._gids[store_rows], docs)
        return gids.contiguous(), sc, cnt

    # -- reference-shaped helpers (kept for callers that use them directly) -------------------
    def _vector_search(self, *, query: str, where: Optional[Mapping[str, object]], k: int) -> List[Mapping[str, object]]:
        q_vec = np.asarray(self.embedder.encode_queries([query])[0], dtype=np.float32)
        gids, sims, cnt = [t.cpu().numpy() for t in self._vector_search_device(q_vec, where, k)[:3]]
        col = self.vector_store._ensure_collection()
        out = []
        for j in range(int(cnt[0])):
            cid = REGISTRY.name(int(gids[0, j]))
            r = col.row_of[cid]
            out.append({"id": cid, "document": col.documents[r], "metadata": col.metadatas[r],
                        "distance": 1.0 - float(sims[0, j])})
        return out

    def _bm25_search(self, *, query: str, where: Optional[Mapping[str, object]], k: int) -> List[Mapping[str, object]]:
        return self.bm25_store.search(query=query, where=where, top_k=k)

    def _read_back(self, tensors):
        """Device results -> pinned host buffers (cached per shape), asynchronous copies, one wait."""
        cache = getattr(self, "_pin_out", None)
        key = tuple((tuple(t.shape), t.dtype) for t in tensors)
        if cache is None or cache[0] != key:
            cache = self._pin_out = (key, [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in tensors])
        for h, t in zip(cache[1], tensors):
            h.copy_(t, non_blocking=True)
        torch.cuda.current_stream(tensors[0].device).synchronize()
        return [h.numpy() for h in cache[1]]

    # -- the entry point -----------------------------------------------------------------------------
    def retrieve(self, *, question: str, filters: Optional[Mapping[str, object]] = None, top_k: int = 8,
                 hybrid: bool = True) -> List[Dict[str, object]]:
        raw_filters = filters or {}
        chroma_where = build_where_filter(raw_filters) if raw_filters else None
        bm_where = raw_filters or None

        q_vec = np.asarray(self.embedder.encode_queries([question])[0], dtype=np.float32)
        k_vec = self.k_vector if hybrid else max(top_k, self.k_vector)
        # every kernel of the call is enqueued without a host synchronisation; results and the dense
        # certificate flag come back in ONE read-back into pinned buffers kept on the retriever
        *vec, flags = self._vector_search_device(q_vec, chroma_where, k_vec, certified=False)
        bm = self._bm25_search_device(question, bm_where, self.k_bm25) if hybrid else None
        dev_out = ops.hybrid_fuse(tuple(vec), bm, top_k=top_k, rrf_k=self.rrf_k,
                                  w_vec=self.weight_vector if hybrid else 1.0, w_bm=self.weight_bm25)
        host = self._read_back((*dev_out, flags))
        if int(host[5][0]) != 0:
            # thousands of exact duplicates around rank k: the fp32 pass could not certify its pool;
            # the exhaustive float64 scan serves this query (rare; one more round trip)
            *vec, _ = self._vector_search_device(q_vec, chroma_where, k_vec, certified=True)
            dev_out = ops.hybrid_fuse(tuple(vec), bm, top_k=top_k, rrf_k=self.rrf_k,
                                      w_vec=self.weight_vector if hybrid else 1.0, w_bm=self.weight_bm25)
            host = self._read_back((*dev_out, flags))
        ids, fused, vdist, bmsc, cnt = host[:5]

        col = self.vector_store._ensure_collection()
        entries = self.bm25_store._entries
        out: List[Dict[str, object]] = []
        for j in range(int(cnt[0])):
            cid = REGISTRY.name(int(ids[0, j]))
            vd = None if math.isnan(float(vdist[0, j])) else float(vdist[0, j])
            bs = None if math.isnan(float(bmsc[0, j])) else float(bmsc[0, j])
            document, metadata = None, {}
            if vd is not None:       # came from the vector list: its record wins when non-empty
                r = col.row_of[cid]
                document = None or col.documents[r]
                metadata = {} or col.metadatas[r] or {}
            if bs is not None:
                e = entries[cid]
                if not document and e.text:
                    document = e.text
                if not metadata and e.metadata:
                    metadata = e.metadata or {}
            out.append({"id": cid, "document": document, "metadata": metadata,
                        "scores": {"vector_distance": vd, "bm25_score": bs, "fused": float(fused[0, j])}})
        return out
