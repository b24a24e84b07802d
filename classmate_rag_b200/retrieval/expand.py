"""Neighbor expansion and the per-document cap: the drop-in for the reference's
rag/retrieval/expand.py:98-153 (driven by rag/pipeline/rag.py:429-455).

Integer / hash / string work on at most top_k * (1 + 2 * radius) records, so it stays on
the host.  The reference re-parses the whole BM25 JSONL catalog on every call; here the
catalog can also be handed over once (``use_catalog``) by the store that already holds it.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Callable, Dict, List, Mapping, Optional, Sequence, Tuple

from .ids import stable_chunk_id

_BM25_JSONL = Path("./indexes/bm25/bm25_index.jsonl")
_catalog_provider: Optional[Callable[[], Mapping[str, Tuple[str, Dict[str, object]]]]] = None


def use_catalog(provider) -> None:
    """``provider``: None (read the JSONL file, as the reference does), a mapping
    id -> (text, metadata), a BM25Store, or a callable returning such a mapping."""
    global _catalog_provider
    if provider is None:
        _catalog_provider = None
    elif hasattr(provider, "catalog"):
        _catalog_provider = provider.catalog
    elif callable(provider):
        _catalog_provider = provider
    else:
        _catalog_provider = lambda: provider  # noqa: E731


def _load_bm25_catalog() -> Mapping[str, Tuple[str, Dict[str, object]]]:
    """id -> (text, metadata); an unreadable line is skipped, a missing file is empty."""
    if _catalog_provider is not None:
        return _catalog_provider()
    out: Dict[str, Tuple[str, Dict[str, object]]] = {}
    if not _BM25_JSONL.exists():
        return out
    with _BM25_JSONL.open("r", encoding="utf-8", errors="ignore") as f:
        for raw in f:
            raw = raw.strip()
            if not raw:
                continue
            try:
                rec = json.loads(raw)
                cid = str(rec.get("id") or "")
                if cid:
                    out[cid] = (str(rec.get("text") or ""), dict(rec.get("metadata") or {}))
            except Exception:
                continue
    return out


def _neighbor_ids(meta: Mapping[str, object], *, radius: int) -> List[str]:
    """Ids of chunks chunk_id-radius .. chunk_id+radius of the same file, hashed with the
    SEED's page (chunk ids are per file, so neighbors across a page break never resolve --
    the reference behaves the same way)."""
    sp, page, cid = meta.get("source_path"), meta.get("page"), meta.get("chunk_id")
    if sp is None or page is None or cid is None:
        return []
    try:
        page_i, cid_i = int(page), int(cid)
    except Exception:
        return []
    course, unit = meta.get("course") or None, meta.get("unit") or None
    return [stable_chunk_id(source_path=Path(str(sp)), page=page_i, chunk_index=cid_i + d, course=course, unit=unit)
            for d in range(-radius, radius + 1) if d != 0]


def expand_with_neighbors(results: Sequence[Dict[str, object]], *, radius: int = 1,
                          max_per_doc: Optional[int] = None, neighbor_penalty: float = 0.001) -> List[Dict[str, object]]:
    """Each hit, in order, followed by its catalogued non-blank neighbors (score just below
    the seed's), deduplicated by id; then at most ``max_per_doc`` records per source_path.
    Records are {id, document, score, metadata}; the seed score is read from key "score"
    (HybridRetriever.retrieve does not set it, so it is 0.0, as in the reference)."""
    catalog = _load_bm25_catalog()
    seen = set()
    expanded: List[Dict[str, object]] = []
    for r in results:
        rid = str(r.get("id") or "")
        if not rid or rid in seen:
            continue
        seen.add(rid)
        score = float(r.get("score") or 0.0)
        meta = dict(r.get("metadata") or {})
        expanded.append({"id": rid, "document": str(r.get("document") or ""), "score": score, "metadata": meta})
        if radius > 0:
            for nid in _neighbor_ids(meta, radius=radius):
                if nid in seen or nid not in catalog:
                    continue
                ntext, nmeta = catalog[nid]
                if not (ntext or "").strip():
                    continue
                expanded.append({"id": nid, "document": ntext, "score": score - neighbor_penalty, "metadata": nmeta})
                seen.add(nid)
    if max_per_doc and max_per_doc > 0:
        per_doc: Dict[str, int] = {}
        kept = []
        for it in expanded:
            sp = str(it["metadata"].get("source_path") or "")
            if per_doc.get(sp, 0) < max_per_doc:
                kept.append(it)
                per_doc[sp] = per_doc.get(sp, 0) + 1
        expanded = kept
    return expanded
