"""NumPy / pure-Python CPU oracle for CLASSMATE-RAG's hybrid retrieval path.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Every function cites the
reference lines (relative to /root/reference) whose behaviour it restates.

Numerical contracts (shared with the CUDA path, stated in DESIGN.md):

* Dense score.  The corpus matrix and the query are bf16 (round-to-nearest-even
  of the fp32 E5 output).  ``exact score = dot(q, c)`` evaluated in float64 in a
  FIXED order: the row is cut into 16-byte vectors of 8 elements; 32 partial
  sums, partial ``l`` owns vectors ``l, l+32, l+64, ...`` and adds the (exact)
  products of their elements sequentially (vector by vector, element by
  element); then the 32 partials are combined by the halving tree
  ``a[l] += a[l+off]`` for off = 16, 8, 4, 2, 1.  (dim must be a multiple of 8.)
  A bf16*bf16 product is exact in float64, so the only roundings are the adds,
  and their order is pinned -> the CUDA rescoring kernel reproduces the value
  bit for bit.  ``distance = 1.0 - score`` (reference: hnswlib cosine space on
  unit vectors, rag/retrieval/vector_chroma.py:149-164,204-253).
* Ranking order everywhere: score descending, then row index ascending
  (reference: stable ``sorted(..., reverse=True)`` over insertion order,
  rag/retrieval/bm25.py:199).
* BM25: float64, operation for operation as rank_bm25.BM25Okapi.get_scores.
* RRF: float64, ``w * (1.0 / (rrf_k + rank))`` (rag/retrieval/fusion.py:17-36).
* MMR: the reference multiplies fp32 matrices through BLAS
  (rag/retrieval/fusion.py:39-61), whose summation order is unspecified.  The
  oracle pins it: similarities are the exact float64 dots defined above on the
  bf16 embeddings, the greedy combine is float64.
"""
from __future__ import annotations

import math
import re
from hashlib import blake2b
from pathlib import Path
from typing import Any, Dict, List, Mapping, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------
# bf16 helpers
# --------------------------------------------------------------------------


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16, returned as uint16 bit patterns."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    bias = ((u >> 16) & 1) + 0x7FFF
    return ((u + bias) >> 16).astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


def bf16_bits_to_f64(b: np.ndarray) -> np.ndarray:
    return bf16_bits_to_f32(b).astype(np.float64)


# --------------------------------------------------------------------------
# A1: exact dense top-k (replaces hnswlib behind ChromaVectorStore.query,
#     rag/retrieval/vector_chroma.py:204-253)
# --------------------------------------------------------------------------

_LANES = 32


def _tree32(acc: np.ndarray) -> np.ndarray:
    """acc[..., 32] -> [...]: the pinned halving tree."""
    for off in (16, 8, 4, 2, 1):
        acc = acc[..., :off] + acc[..., off:2 * off]
    return acc[..., 0]


def exact_dots(q_bits: np.ndarray, c_bits: np.ndarray, chunk: int = 8192) -> np.ndarray:
    """float64 dot of one bf16 query against bf16 rows in the pinned order."""
    q_bits = np.asarray(q_bits, dtype=np.uint16).reshape(-1)
    c_bits = np.asarray(c_bits, dtype=np.uint16)
    if c_bits.ndim == 1:
        c_bits = c_bits[None, :]
    n, d = c_bits.shape
    assert q_bits.shape[0] == d and d % 8 == 0
    nvec = d // 8
    per_lane = (nvec + _LANES - 1) // _LANES
    dp = per_lane * _LANES * 8
    q = np.zeros(dp, dtype=np.float64)
    q[:d] = bf16_bits_to_f64(q_bits)
    out = np.empty(n, dtype=np.float64)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        c = np.zeros((hi - lo, dp), dtype=np.float64)
        c[:, :d] = bf16_bits_to_f64(c_bits[lo:hi])
        prod = (c * q).reshape(hi - lo, per_lane, _LANES, 8)
        acc = np.zeros((hi - lo, _LANES), dtype=np.float64)
        for j in range(per_lane):      # lane l: vectors l, l+32, ... in order
            for e in range(8):         # elements of a vector in order
                acc = acc + prod[:, j, :, e]
        out[lo:hi] = _tree32(acc)
    return out


def exact_dot_pair(a_bits: np.ndarray, b_bits: np.ndarray) -> float:
    return float(exact_dots(a_bits, np.asarray(b_bits, dtype=np.uint16)[None, :])[0])


def order_desc_then_index(scores: np.ndarray, ids: Optional[np.ndarray] = None) -> np.ndarray:
    """Permutation sorting by (score desc, id asc)."""
    scores = np.asarray(scores, dtype=np.float64)
    if ids is None:
        ids = np.arange(scores.shape[0], dtype=np.int64)
    return np.lexsort((ids, -scores))


def dense_topk(q_bits: np.ndarray, c_bits: np.ndarray, k: int,
               mask: Optional[np.ndarray] = None,
               row_offset: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Exact brute-force top-k.  Returns (global ids int64[k'], scores f64[k'])."""
    scores = exact_dots(q_bits, c_bits)
    ids = np.arange(scores.shape[0], dtype=np.int64)
    if mask is not None:
        keep = np.asarray(mask).astype(bool)
        scores, ids = scores[keep], ids[keep]
    order = order_desc_then_index(scores, ids)[:k]
    return ids[order] + row_offset, scores[order]


# --------------------------------------------------------------------------
# A2: BM25 (rank_bm25.BM25Okapi restated; NOT vendored in the reference --
#     requirements.txt:4; call sites rag/retrieval/bm25.py:25,145,191,197).
#     parity unpinned for this class: no install, no reference test.
# --------------------------------------------------------------------------

BM25_K1 = 1.5
BM25_B = 0.75
BM25_EPS = 0.25


class BM25Okapi:
    """Okapi BM25 as published in rank_bm25 0.2.x (defaults k1=1.5, b=0.75,
    epsilon=0.25; the reference passes only the corpus, bm25.py:145,191)."""

    def __init__(self, corpus: Sequence[Sequence[str]], tokenizer=None,
                 k1: float = BM25_K1, b: float = BM25_B, epsilon: float = BM25_EPS):
        self.k1, self.b, self.epsilon = k1, b, epsilon
        self.corpus_size = 0
        self.doc_len: List[int] = []
        self.doc_freqs: List[Dict[str, int]] = []
        self.idf: Dict[str, float] = {}
        if tokenizer is not None:
            corpus = [tokenizer(doc) for doc in corpus]
        containing: Dict[str, int] = {}  # insertion order = first appearance
        total = 0
        for doc in corpus:
            self.doc_len.append(len(doc))
            total += len(doc)
            tf: Dict[str, int] = {}
            for w in doc:
                tf[w] = tf.get(w, 0) + 1
            self.doc_freqs.append(tf)
            for w in tf:
                containing[w] = containing.get(w, 0) + 1
            self.corpus_size += 1
        self.avgdl = total / self.corpus_size
        self.nd = containing
        # idf with the epsilon floor; the running sum is sequential in
        # first-appearance order (it fixes the last bits of average_idf)
        idf_sum = 0.0
        negative: List[str] = []
        for w, n_w in containing.items():
            v = math.log(self.corpus_size - n_w + 0.5) - math.log(n_w + 0.5)
            self.idf[w] = v
            idf_sum += v
            if v < 0:
                negative.append(w)
        self.average_idf = idf_sum / len(self.idf)
        floor = self.epsilon * self.average_idf
        for w in negative:
            self.idf[w] = floor

    def get_scores(self, query: Sequence[str]) -> np.ndarray:
        score = np.zeros(self.corpus_size)
        doc_len = np.array(self.doc_len)
        for tok in query:  # in order, duplicates repeated
            q_freq = np.array([(d.get(tok) or 0) for d in self.doc_freqs])
            score += (self.idf.get(tok) or 0) * (
                q_freq * (self.k1 + 1)
                / (q_freq + self.k1 * (1 - self.b + self.b * doc_len / self.avgdl)))
        return score


def bm25_idf_table(df: np.ndarray, n_docs: int, vocab_order: Optional[np.ndarray] = None,
                   epsilon: float = BM25_EPS) -> Tuple[np.ndarray, float]:
    """idf[V] float64 with the epsilon floor, from document frequencies.

    ``vocab_order`` lists the term ids with df>0 in first-appearance order (the
    dict order rank_bm25 sums in); default ascending id.  Terms with df==0 get
    idf 0.0 (rank_bm25: ``idf.get(q) or 0``)."""
    df = np.asarray(df, dtype=np.int64)
    if vocab_order is None:
        vocab_order = np.nonzero(df > 0)[0]
    idf = np.zeros(df.shape[0], dtype=np.float64)
    idf_sum = 0.0
    neg = []
    for t in vocab_order.tolist():
        n_t = int(df[t])
        v = math.log(n_docs - n_t + 0.5) - math.log(n_t + 0.5)
        idf[t] = v
        idf_sum += v
        if v < 0:
            neg.append(t)
    average_idf = idf_sum / max(1, len(vocab_order))
    if neg:
        idf[np.asarray(neg, dtype=np.int64)] = epsilon * average_idf
    return idf, average_idf


def bm25_scores_csr(term_ptr: np.ndarray, post_doc: np.ndarray, post_tf: np.ndarray,
                    doc_len: np.ndarray, idf: np.ndarray, avgdl: float,
                    query_terms: Sequence[int], k1: float = BM25_K1, b: float = BM25_B) -> np.ndarray:
    """get_scores over a CSR index, float64, operation for operation.

    Skipping docs that do not contain the term is exact: their contribution in
    rank_bm25 is ``idf * (0 * 2.5 / (0 + ...)) = +-0.0``."""
    n = doc_len.shape[0]
    score = np.zeros(n, dtype=np.float64)
    dl = np.asarray(doc_len)
    for t in query_terms:
        if t < 0 or t >= idf.shape[0]:
            continue
        lo, hi = int(term_ptr[t]), int(term_ptr[t + 1])
        if hi <= lo:
            continue
        d = post_doc[lo:hi].astype(np.int64)
        q_freq = post_tf[lo:hi].astype(np.int64)
        score[d] += float(idf[t]) * (
            q_freq * (k1 + 1) / (q_freq + k1 * (1 - b + b * dl[d] / avgdl)))
    return score


def bm25_topk(scores: np.ndarray, k: int, row_offset: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Stable descending sort, zeros included (bm25.py:199)."""
    order = order_desc_then_index(scores)[:k]
    return order.astype(np.int64) + row_offset, scores[order]


# ---- tokeniser / filters (rag/retrieval/bm25.py:34-107) -------------------

TOKEN_RE = re.compile(r"[A-Za-zÀ-ÖØ-öø-ÿ]+")

STOP_EN = frozenset(
    "a an the and or but if then else for to of in on at by with from as is are was were be been "
    "being it its this that these those i you he she we they them his her their my your our me us "
    "not no yes do does did doing can could should would may might will shall about into over under "
    "again further there here when where why how what which who whom".split())

STOP_IT = frozenset(
    "un uno una le la il lo gli i l e o ma se allora altrimenti per di a da in su con come è era "
    "sono siamo siete fui fu furono essere stato questo questa questi queste quello quella quelli "
    "quelle ciò cio io tu lui lei noi voi loro mio mia tuo tua suo sua nostro vostro non no si sia "
    "fare fa fatto posso può puo puoi possono dovrebbe potrebbe sarà sara sarebbe saremmo sarete "
    "siano che perché perche quando dove cosa quale chi".split())


def choose_stopwords(lang_hint: Optional[str]) -> frozenset:
    lang = (lang_hint or "").lower()
    if lang.startswith("it"):
        return STOP_IT
    return STOP_EN  # "en*" and unknown both map to EN (bm25.py:54-61)


def tokenize(text: str, lang_hint: Optional[str] = None) -> List[str]:
    """bm25.py:63-70: letter runs, lowercase, drop stopwords and len<=1."""
    sw = choose_stopwords(lang_hint)
    out = []
    for m in TOKEN_RE.finditer(text or ""):
        t = m.group(0).lower()
        if len(t) > 1 and t not in sw:
            out.append(t)
    return out


FILTER_FIELDS = ("course", "unit", "language", "doc_type", "author", "semester")


def matches_filter(meta: Mapping[str, Any], where: Optional[Mapping[str, Any]]) -> bool:
    """bm25.py:79-107 (including quirk Q1: a None-valued filter key only
    matches docs lacking the field)."""
    if not where:
        return True
    if "$and" in where:
        return all(matches_filter(meta, c) for c in where["$and"])
    tg = where.get("tags") if "tags" in where else None
    if isinstance(tg, dict) and "$contains" in tg:
        want = tg["$contains"]
        if not want:
            return True
        want_set = {want} if isinstance(want, str) else set(want)
        return want_set.issubset(set(meta.get("tags") or []))
    for f in FILTER_FIELDS:
        if f in where and meta.get(f) != where[f]:
            return False
    return True


def slug_tag(t: str) -> str:
    """vector_chroma.py:22-26."""
    s = re.sub(r"[^a-z0-9]+", "_", (t or "").lower().strip())
    return s.strip("_")


def parse_tags(obj) -> List[str]:
    """vector_chroma.py:29-42."""
    if not obj:
        return []
    vals = [str(x) for x in obj] if isinstance(obj, (list, tuple)) else str(obj).split(",")
    return [v.strip() for v in vals if v.strip()]


def build_where_filter(meta_like: Mapping[str, Any]) -> Optional[Dict[str, Any]]:
    """vector_chroma.py:45-78."""
    if not meta_like:
        return None
    clauses: List[Dict[str, Any]] = []
    for f in FILTER_FIELDS:
        v = meta_like.get(f)
        if v is None:
            continue
        if isinstance(v, str):
            v = v.strip()
            if not v or (f == "doc_type" and v.lower() == "other"):
                continue
        clauses.append({f: v})
    for t in parse_tags(meta_like.get("tags")):
        s = slug_tag(t)
        if s:
            clauses.append({f"tag_{s}": True})
    if not clauses:
        return None
    return clauses[0] if len(clauses) == 1 else {"$and": clauses}


def chroma_where_matches(meta: Mapping[str, Any], where: Optional[Mapping[str, Any]]) -> bool:
    """Chroma's metadata ``where`` semantics for the only shapes
    build_where_filter emits: ``{field: value}`` equality and ``{"$and": [...]}``."""
    if not where:
        return True
    if "$and" in where:
        return all(chroma_where_matches(meta, c) for c in where["$and"])
    for key, val in where.items():
        if key not in meta or meta[key] != val:
            return False
    return True


def bm25_store_search(entries: Sequence[Tuple[str, List[str], Mapping[str, Any]]],
                      query: str, where: Optional[Mapping[str, Any]], top_k: int,
                      query_lang: str = "en") -> List[Tuple[str, float]]:
    """BM25Store.search (bm25.py:175-212) over ``entries`` = [(id, tokens, meta)]
    in insertion order: filter -> BM25Okapi over the SUBSET -> stable sort."""
    if not query.strip() or not entries:
        return []
    cand = [e for e in entries if matches_filter(e[2], where)]
    if not cand:
        return []
    bm = BM25Okapi([e[1] for e in cand] or [[""]])
    q_tokens = tokenize(query, lang_hint=query_lang)
    scores = bm.get_scores(q_tokens)
    ranked = sorted(zip([e[0] for e in cand], scores), key=lambda x: x[1], reverse=True)[:top_k]
    return [(i, float(s)) for i, s in ranked]


# --------------------------------------------------------------------------
# A3: RRF (rag/retrieval/fusion.py:17-36)
# --------------------------------------------------------------------------


def rrf_fuse(rank_lists: Sequence[Sequence[Any]], weights: Optional[Sequence[float]] = None,
             rrf_k: int = 60) -> Dict[Any, float]:
    if not rank_lists:
        return {}
    if weights is None:
        weights = [1.0] * len(rank_lists)
    elif len(weights) != len(rank_lists):
        raise ValueError("weights length must match rank_lists length")
    fused: Dict[Any, float] = {}
    for w, ids in zip(weights, rank_lists):
        w = float(w)
        for pos, _id in enumerate(ids):
            fused[_id] = fused.get(_id, 0.0) + w * (1.0 / (rrf_k + (pos + 1)))
    return fused


# --------------------------------------------------------------------------
# A4: MMR (rag/retrieval/fusion.py:39-61), precision pinned to float64
# --------------------------------------------------------------------------


def mmr_order(sims_q: np.ndarray, sims_cc: np.ndarray, k: int, lambd: float = 0.5) -> List[int]:
    """Greedy MMR over precomputed float64 similarities.  First pick = argmax
    sims_q (lowest index on ties); then strict '>' over ascending index."""
    n = int(sims_q.shape[0])
    if n == 0:
        return []
    sims_q = np.asarray(sims_q, dtype=np.float64)
    sims_cc = np.asarray(sims_cc, dtype=np.float64)
    selected = [int(np.argmax(sims_q))]
    remaining = [i for i in range(n) if i != selected[0]]
    while remaining and len(selected) < min(k, n):
        best_i, best_s = None, -1e9
        for i in remaining:  # ascending index
            div = max(float(sims_cc[i, j]) for j in selected)
            s = lambd * float(sims_q[i]) - (1.0 - lambd) * div
            if s > best_s:
                best_s, best_i = s, i
        selected.append(int(best_i))
        remaining.remove(int(best_i))
    return selected


def mmr_order_bf16(q_bits: np.ndarray, cand_bits: np.ndarray, k: int, lambd: float = 0.5) -> List[int]:
    """MMR with the pinned exact-dot similarities on bf16 inputs."""
    cand_bits = np.asarray(cand_bits, dtype=np.uint16)
    n = cand_bits.shape[0]
    if n == 0:
        return []
    sims_q = exact_dots(q_bits, cand_bits)
    sims_cc = np.empty((n, n), dtype=np.float64)
    for i in range(n):
        sims_cc[i] = exact_dots(cand_bits[i], cand_bits)
    return mmr_order(sims_q, sims_cc, k, lambd)


# --------------------------------------------------------------------------
# A5: HybridRetriever.retrieve merge + final order (fusion.py:108-167)
# --------------------------------------------------------------------------


def hybrid_merge(vec: Sequence[Tuple[Any, float]], bm: Sequence[Tuple[Any, float]],
                 top_k: int, rrf_k: int = 60, w_vec: float = 1.0, w_bm: float = 1.0,
                 hybrid: bool = True) -> List[Dict[str, Any]]:
    """``vec`` = [(id, distance)] in post-MMR order, ``bm`` = [(id, score)].
    Returns the fused items in final order, each
    {id, fused, vector_distance|None, bm25_score|None}."""
    vec_ids = [i for i, _ in vec]
    bm_ids = [i for i, _ in bm] if hybrid else []
    fused = rrf_fuse([vec_ids, bm_ids] if hybrid else [vec_ids],
                     [w_vec, w_bm] if hybrid else [1.0], rrf_k)
    items: Dict[Any, Dict[str, Any]] = {}
    for i, dist in vec:
        it = items.setdefault(i, {"id": i, "fused": 0.0, "vector_distance": None, "bm25_score": None})
        it["vector_distance"] = dist
    if hybrid:
        for i, sc in bm:
            it = items.setdefault(i, {"id": i, "fused": 0.0, "vector_distance": None, "bm25_score": None})
            it["bm25_score"] = sc
    for i, s in fused.items():
        if i in items:
            items[i]["fused"] = float(s)

    def key(it):
        vd = it["vector_distance"]
        return (it["fused"] or 0.0, -(vd if isinstance(vd, (int, float)) else 0.0))

    return sorted(items.values(), key=key, reverse=True)[:top_k]


# --------------------------------------------------------------------------
# A6: neighbor expansion (rag/retrieval/expand.py:63-153, rag/utils/ids.py:17-29)
# --------------------------------------------------------------------------


def stable_chunk_id(source_path, page: int, chunk_index: int, course: Optional[str] = None,
                    unit: Optional[str] = None, prefix: str = "cm_") -> str:
    sp = str(Path(source_path).resolve())
    key = "|".join([sp, str(page), str(chunk_index), course or "", unit or ""])
    return prefix + blake2b(key.encode("utf-8"), digest_size=16).hexdigest()


def neighbor_ids(meta: Mapping[str, Any], radius: int) -> List[str]:
    sp, page, cid = meta.get("source_path"), meta.get("page"), meta.get("chunk_id")
    if sp is None or page is None or cid is None:
        return []
    try:
        page_i, cid_i = int(page), int(cid)
    except Exception:
        return []
    course, unit = meta.get("course") or None, meta.get("unit") or None
    return [stable_chunk_id(Path(str(sp)), page_i, cid_i + d, course, unit)
            for d in range(-radius, radius + 1) if d != 0]


def expand_with_neighbors(results: Sequence[Mapping[str, Any]],
                          catalog: Mapping[str, Tuple[str, Mapping[str, Any]]],
                          radius: int = 1, max_per_doc: Optional[int] = None,
                          neighbor_penalty: float = 0.001) -> List[Dict[str, Any]]:
    seen, out = set(), []
    for r in results:
        rid = str(r.get("id") or "")
        if not rid or rid in seen:
            continue
        seen.add(rid)
        sc = float(r.get("score") or 0.0)  # retrieve() never sets "score" (quirk Q3)
        meta = dict(r.get("metadata") or {})
        out.append({"id": rid, "document": str(r.get("document") or ""), "score": sc, "metadata": meta})
        if radius > 0:
            for nid in neighbor_ids(meta, radius):
                if nid in seen or nid not in catalog:
                    continue
                ntext, nmeta = catalog[nid]
                if not (ntext or "").strip():
                    continue
                out.append({"id": nid, "document": ntext, "score": sc - neighbor_penalty, "metadata": nmeta})
                seen.add(nid)
    if max_per_doc and max_per_doc > 0:
        counts: Dict[str, int] = {}
        kept = []
        for it in out:
            sp = str(it["metadata"].get("source_path") or "")
            if counts.get(sp, 0) < max_per_doc:
                kept.append(it)
                counts[sp] = counts.get(sp, 0) + 1
        out = kept
    return out


# --------------------------------------------------------------------------
# A9: near-duplicate cosine filter (extension; greedy keep-first rule of
#     rag/utils/dedup.py:40-55 with the '>=' comparison of :50)
# --------------------------------------------------------------------------


def neardup_keep_mask(c_bits: np.ndarray, threshold: float = 0.95) -> np.ndarray:
    """keep[i] = no previously KEPT j<i has exact_dot(c_i, c_j) >= threshold."""
    c_bits = np.asarray(c_bits, dtype=np.uint16)
    n = c_bits.shape[0]
    keep = np.zeros(n, dtype=bool)
    kept_rows: List[int] = []
    for i in range(n):
        dup = False
        if kept_rows:
            sims = exact_dots(c_bits[i], c_bits[np.asarray(kept_rows)])
            dup = bool(np.any(sims >= threshold))
        if not dup:
            keep[i] = True
            kept_rows.append(i)
    return keep


# --------------------------------------------------------------------------
# N4: ingest-side text dedup (rag/utils/dedup.py:19-55): Jaccard similarity of
#     token 5-gram shingle sets, greedy keep-first, comparison '>=' (:50).
#     Pinned by tests/golden/reference_dedup.json (live reference outputs).
# --------------------------------------------------------------------------
_DEDUP_PUNCT = re.compile(r"[^\w\s]", re.UNICODE)


def dedup_norm_tokens(text: str) -> List[str]:
    """dedup.py:19-23: lower-case, punctuation -> space, split on whitespace."""
    return _DEDUP_PUNCT.sub(" ", (text or "").lower()).split()


def dedup_shingles(tokens: Sequence[str], k: int = 5) -> set:
    """dedup.py:25-29: all k-grams; a shorter non-empty text is one shingle; empty -> empty set."""
    toks = tuple(tokens)
    if not toks:
        return set()
    if len(toks) < k:
        return {toks}
    return {toks[i:i + k] for i in range(len(toks) - k + 1)}


def dedup_jaccard(a: set, b: set) -> float:
    """dedup.py:31-38."""
    if not a and not b:
        return 1.0
    if not a or not b:
        return 0.0
    inter = len(a & b)
    return inter / (len(a) + len(b) - inter)


def dedup_keep_indices(blocks: Sequence[str], threshold: float = 0.92) -> List[int]:
    """dedup.py:40-55 as indices: block i is kept iff no previously KEPT block has Jaccard >= threshold."""
    kept: List[int] = []
    kept_sets: List[set] = []
    for i, text in enumerate(blocks):
        sh = dedup_shingles(dedup_norm_tokens(text))
        if not any(dedup_jaccard(sh, other) >= threshold for other in kept_sets):
            kept.append(i)
            kept_sets.append(sh)
    return kept
