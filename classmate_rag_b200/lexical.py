"""Device-resident BM25 inverted index (term-major CSR + tile skip table).

Build side of A7 (reference: BM25Store.upsert_many/_rebuild,
rag/retrieval/bm25.py:140-166, which rebuilds rank_bm25.BM25Okapi from the
token lists).  The statistics follow rank_bm25 exactly: idf uses
``math.log`` on the host, its epsilon floor uses the average idf summed
sequentially in vocabulary first-appearance order, avgdl = total tokens / N.
torch is used for sorting and scans (plumbing); scoring is in libcmrag.
"""
from __future__ import annotations

import ctypes as C
import itertools
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import torch

BM25_K1 = 1.5
BM25_B = 0.75
BM25_EPS = 0.25
DEFAULT_TILE_DOCS = 2048   # 16 documents per thread of a 128-thread CTA (csrc/bm25.cu)
DENSE_DENSITY = 0.125    # terms in at least this share of the documents get a factor column
DENSE_MAX_TERMS = 64     # 8 B x n_docs each
HEAD_TERMS = 64          # columns of the fp16 head matrix (one 128-byte row per document; csrc/bm25_mma.cu)
SKIP_BUDGET_BYTES = 1 << 30   # skip-table rows ((n_tiles + 1) x 4 B each) go to the most frequent terms within this


class LexIndexStruct(C.Structure):
    """Mirror of ``cmr_lex_index`` (include/cmrag.h)."""
    _fields_ = [
        ("term_ptr", C.c_void_p), ("tile_skip", C.c_void_p), ("post_pack", C.c_void_p), ("imp_table", C.c_void_p),
        ("post_doc", C.c_void_p), ("post_imp", C.c_void_p), ("post_tf", C.c_void_p), ("doc_len", C.c_void_p),
        ("idf", C.c_void_p), ("dense_imp", C.c_void_p), ("dense_slot", C.c_void_p),
        ("head_mat", C.c_void_p), ("head_slot", C.c_void_p), ("skip_row", C.c_void_p),
        ("n_docs", C.c_int64), ("n_terms", C.c_int32), ("tile_docs", C.c_int32), ("n_tiles", C.c_int32),
        ("n_codes", C.c_int32), ("n_dense", C.c_int32), ("n_head", C.c_int32),
        ("avgdl", C.c_double), ("k1", C.c_double), ("b", C.c_double),
    ]


def idf_table(df: np.ndarray, n_docs: int, vocab_order: np.ndarray, epsilon: float = BM25_EPS):
    """rank_bm25 idf with the epsilon floor.  ``vocab_order``: term ids with
    df > 0 in first-appearance order (the order the running sum is taken in).

    Bit-identical to rank_bm25's Python loop without being one: the logarithms are libm's
    (``math.log`` mapped over the values -- NumPy's vectorised log may differ in the last bit),
    the subtraction is the same IEEE operation in NumPy, and the running sum is
    ``np.cumsum`` (a sequential left-to-right scan, unlike ``np.sum``'s pairwise tree)."""
    idf = np.zeros(df.shape[0], dtype=np.float64)
    order = np.asarray(vocab_order, dtype=np.int64)
    if order.size == 0:
        return idf, 0.0
    n_t = df[order].astype(np.float64)
    log_a = np.fromiter(map(math.log, (float(n_docs) - n_t + 0.5).tolist()), dtype=np.float64, count=order.size)
    log_b = np.fromiter(map(math.log, (n_t + 0.5).tolist()), dtype=np.float64, count=order.size)
    v = log_a - log_b
    idf[order] = v
    average_idf = float(np.cumsum(v)[-1]) / max(1, order.size)
    neg = order[v < 0]
    if neg.size:
        idf[neg] = epsilon * average_idf
    return idf, average_idf


@dataclass
class LexicalIndex:
    term_ptr: torch.Tensor      # int64 [V+1]
    tile_skip: torch.Tensor     # int32 (uint32 bits) [n_skip_rows, n_tiles+1]: rows of the most frequent terms
    post_doc: torch.Tensor      # int32 [P]
    post_imp: Optional[torch.Tensor]   # float64 [P] (wide format only)
    post_tf: torch.Tensor       # int16 (uint16 bits) [P]
    doc_len: torch.Tensor       # int32 [N]
    idf: torch.Tensor           # float64 [V]
    n_docs: int
    n_terms: int
    tile_docs: int
    n_tiles: int
    avgdl: float
    k1: float = BM25_K1
    b: float = BM25_B
    post_pack: Optional[torch.Tensor] = None   # int32 (uint32 bits) [P]: code << 16 | tile-local doc
    imp_table: Optional[torch.Tensor] = None   # float64 [n_codes]: factor of each distinct (tf, doc_len) pair
    pair_tf: Optional[torch.Tensor] = None     # int32 [n_codes]
    pair_dl: Optional[torch.Tensor] = None     # int32 [n_codes]
    dense_imp: Optional[torch.Tensor] = None   # float64 [n_dense, N]: factor column of each dense term
    dense_slot: Optional[torch.Tensor] = None  # int32 [V]: column of the term or -1
    dense_terms: Optional[np.ndarray] = field(default=None, repr=False)  # term id of each column
    head_mat: Optional[torch.Tensor] = None    # float16 [N, 64]: fp16(idf * factor) of the head terms, 0 where absent
    head_slot: Optional[torch.Tensor] = None   # int32 [V]: column of the term in head_mat or -1
    head_terms: Optional[np.ndarray] = field(default=None, repr=False)   # term id of each head column
    skip_row: Optional[torch.Tensor] = None    # int32 [V]: row of the term in tile_skip or -1 (bisection on post_doc)
    idf_host: np.ndarray = field(default=None, repr=False)
    df_host: np.ndarray = field(default=None, repr=False)        # corpus-wide df
    shard_df_host: np.ndarray = field(default=None, repr=False)  # postings per term in THIS shard
    _struct: Optional[LexIndexStruct] = field(default=None, repr=False)

    @property
    def n_postings(self) -> int:
        return int(self.post_doc.numel())

    @property
    def device(self):
        return self.post_doc.device

    def struct(self) -> LexIndexStruct:
        if self._struct is None:
            opt = lambda t: None if t is None else t.data_ptr()
            self._struct = LexIndexStruct(
                self.term_ptr.data_ptr(), self.tile_skip.data_ptr(), opt(self.post_pack), opt(self.imp_table),
                self.post_doc.data_ptr(), opt(self.post_imp), self.post_tf.data_ptr(), self.doc_len.data_ptr(),
                self.idf.data_ptr(), opt(self.dense_imp), opt(self.dense_slot),
                opt(self.head_mat), opt(self.head_slot), opt(self.skip_row),
                self.n_docs, self.n_terms, self.tile_docs, self.n_tiles,
                0 if self.imp_table is None else int(self.imp_table.numel()),
                0 if self.dense_imp is None else int(self.dense_imp.shape[0]),
                0 if self.head_terms is None else int(len(self.head_terms)), self.avgdl, self.k1, self.b)
        return self._struct

    def posting_bytes(self, terms: Sequence[int]) -> int:
        """Algorithmic bytes one query streams, multiplicity counted: 4 B per packed
        posting (12 B wide: int32 doc + float64 factor) of every sparse query token,
        8 B per document (the float64 factor column) of every dense one."""
        per = 4 if self.post_pack is not None else 12
        dense = set() if self.dense_terms is None else set(int(t) for t in self.dense_terms)
        total = 0
        for t in terms:
            if 0 <= t < self.n_terms:
                total += 8 * self.n_docs if t in dense else per * int(self.shard_df_host[t])
        return total


@dataclass
class GlobalStats:
    """Corpus-wide statistics for a doc-partitioned (sharded) index, so every
    shard scores with the same idf / avgdl as the unsharded index."""
    n_docs: int
    total_tokens: int
    df: np.ndarray            # int64 [V]
    vocab_order: np.ndarray   # term ids with df>0, first-appearance order


def corpus_stats(doc_ptr: torch.Tensor, tokens: torch.Tensor, n_terms: int) -> GlobalStats:
    """df, total length and first-appearance vocabulary order of a token corpus."""
    n_docs = doc_ptr.numel() - 1
    tok = tokens.long()
    total = int(tok.numel())
    dev = tok.device
    if total == 0:
        return GlobalStats(n_docs, 0, np.zeros(n_terms, dtype=np.int64), np.zeros(0, dtype=np.int64))
    doc_len = (doc_ptr[1:] - doc_ptr[:-1]).long()
    doc_of = torch.repeat_interleave(torch.arange(n_docs, device=dev), doc_len)
    key = torch.unique(tok * n_docs + doc_of)
    df = torch.bincount(key // n_docs, minlength=n_terms)
    first_pos = torch.full((n_terms,), total, dtype=torch.int64, device=dev)
    first_pos.scatter_reduce_(0, tok, torch.arange(total, device=dev), reduce="amin")
    order = torch.argsort(first_pos, stable=True)
    n_seen = int((df > 0).sum())
    return GlobalStats(n_docs, total, df.cpu().numpy().astype(np.int64), order[:n_seen].cpu().numpy())


def bm25_factor(tf: torch.Tensor, dl: torch.Tensor, avgdl: float, k1: float, b: float) -> torch.Tensor:
    """float64 tf*(k1+1) / (tf + k1*(1 - b + b*dl/avgdl)), operation for operation as
    rank_bm25 evaluates it.  (avgdl goes in as a device tensor: torch turns a division
    by a Python scalar into a multiplication by its reciprocal -- a different rounding.)"""
    tf64, dl64 = tf.to(torch.float64), dl.to(torch.float64)
    avgdl_t = torch.tensor(avgdl, dtype=torch.float64, device=tf.device)
    return tf64 * (k1 + 1) / (tf64 + k1 * ((1 - b) + (b * dl64) / avgdl_t))


def build_lexical_index(doc_ptr: torch.Tensor, tokens: torch.Tensor, n_terms: int, *,
                        device=None, tile_docs: int = DEFAULT_TILE_DOCS,
                        stats: Optional[GlobalStats] = None, fmt: str = "auto",
                        k1: float = BM25_K1, b: float = BM25_B, epsilon: float = BM25_EPS,
                        dense_density: Optional[float] = DENSE_DENSITY,
                        dense_max_terms: int = DENSE_MAX_TERMS, head_terms: int = HEAD_TERMS,
                        skip_budget_bytes: int = SKIP_BUDGET_BYTES) -> LexicalIndex:
    """Build the CSR index of the documents ``tokens[doc_ptr[i]:doc_ptr[i+1]]``.

    ``stats`` (optional) supplies corpus-wide df / N / total tokens when this
    call builds one shard of a larger corpus.  ``dense_density`` (None = off):
    terms present in at least that share of the documents also get a dense
    float64 factor column, swept instead of scattered by the kernel.  ``head_terms`` (0 = off):
    the terms with the longest posting lists in this shard, at most 64, get a column of the fp16
    head matrix the batched kernels score on the tensor cores (packed postings only).
    ``skip_budget_bytes``: the skip table (posting offset of every tile start) gets rows for the
    terms with the longest lists in this shard, as many as fit; the rest (short lists) are
    bisected by the kernels, so the table does not grow with the vocabulary."""
    if device is None:
        device = tokens.device
    doc_ptr = doc_ptr.to(device).long()
    tok = tokens.to(device).long()
    n_docs = doc_ptr.numel() - 1
    total = int(tok.numel())
    if tile_docs % 512 != 0 or tile_docs <= 0:
        raise ValueError("tile_docs must be a positive multiple of 512")
    n_tiles = max(1, (n_docs + tile_docs - 1) // tile_docs)
    doc_len = (doc_ptr[1:] - doc_ptr[:-1])
    if stats is None:
        stats = corpus_stats(doc_ptr, tok, n_terms)
    avgdl = (stats.total_tokens / stats.n_docs) if stats.n_docs > 0 else 0.0
    idf_host, _ = idf_table(stats.df, stats.n_docs, stats.vocab_order, epsilon)

    if total > 0:
        doc_of = torch.repeat_interleave(torch.arange(n_docs, device=device), doc_len)
        key, _ = torch.sort(tok * max(n_docs, 1) + doc_of)
        del doc_of
        uniq, counts = torch.unique_consecutive(key, return_counts=True)
        del key
        term = uniq // max(n_docs, 1)
        doc = uniq - term * max(n_docs, 1)
    else:
        uniq = torch.zeros(0, dtype=torch.int64, device=device)
        counts = term = doc = uniq
    term_ptr = torch.searchsorted(term, torch.arange(n_terms + 1, device=device))
    # tile skip table: posting offset (relative to the term) where each tile starts -- one row per
    # term for the most frequent terms (within the budget), none for the rest
    shard_df_all = (term_ptr[1:] - term_ptr[:-1])
    max_rows = max(1, int(skip_budget_bytes) // ((n_tiles + 1) * 4))
    n_listed = int((shard_df_all > 0).sum())
    if n_listed <= max_rows:
        row_terms = torch.nonzero(shard_df_all > 0).flatten()
    else:
        row_terms = torch.argsort(shard_df_all, descending=True, stable=True)[:max_rows].sort().values
    skip_row = torch.full((n_terms,), -1, dtype=torch.int32, device=device)
    skip_row[row_terms] = torch.arange(row_terms.numel(), dtype=torch.int32, device=device)
    tile_starts = torch.clamp(torch.arange(n_tiles + 1, device=device) * tile_docs, max=max(n_docs, 1))
    skip = torch.zeros((max(1, row_terms.numel()), n_tiles + 1), dtype=torch.int64, device=device)
    for lo_r in range(0, row_terms.numel(), 4096):      # bounded temporaries for large vocabularies
        rt = row_terms[lo_r:lo_r + 4096]
        bounds = rt[:, None] * max(n_docs, 1) + tile_starts[None, :]
        skip[lo_r:lo_r + rt.numel()] = (torch.searchsorted(uniq, bounds.reshape(-1)).reshape(rt.numel(), n_tiles + 1)
                                        - term_ptr[rt][:, None])
        del bounds
    if tile_docs > 65536:
        raise ValueError("tile_docs must be <= 65536")
    if fmt not in ("auto", "packed", "wide"):
        raise ValueError("fmt must be auto, packed or wide")
    # distinct (tf, doc_len) pairs -> 16-bit codes into an exact float64 factor table
    post_pack = imp_table = pair_tf = pair_dl = imp = None
    dl_post = doc_len[doc] if total > 0 else doc_len[:0]
    if fmt != "wide" and total > 0 and int(doc_len.max()) < (1 << 31) // 65536:
        pair_key = counts.clamp(max=65535) * (int(doc_len.max()) + 1) + dl_post
        pairs, code = torch.unique(pair_key, return_inverse=True)
        if pairs.numel() <= 65536 and int(counts.max()) <= 65535:
            pair_tf = (pairs // (int(doc_len.max()) + 1)).to(torch.int32)
            pair_dl = (pairs % (int(doc_len.max()) + 1)).to(torch.int32)
            imp_table = (bm25_factor(pair_tf, pair_dl, avgdl, k1, b) if avgdl > 0
                         else torch.zeros(pairs.numel(), dtype=torch.float64, device=device))
            local = doc - (doc // tile_docs) * tile_docs
            pk = (code << 16) | local                      # < 2^32, held in int64 here
            post_pack = torch.where(pk >= (1 << 31), pk - (1 << 32), pk).to(torch.int32).contiguous()
            del pk, local
        del pair_key, code
    elif fmt != "wide" and total == 0:
        post_pack = torch.zeros(1, dtype=torch.int32, device=device)
        imp_table = torch.zeros(1, dtype=torch.float64, device=device)
    if post_pack is None:
        if fmt == "packed":
            raise ValueError("corpus has more than 65536 distinct (tf, doc_len) pairs: packed postings impossible")
        imp = (bm25_factor(counts, dl_post, avgdl, k1, b) if avgdl > 0
               else torch.zeros(counts.numel(), dtype=torch.float64, device=device))
    # dense terms: full float64 factor columns (cmr_lex_index.dense_imp)
    dense_imp = dense_slot = dense_terms = None
    if dense_density is not None and total > 0 and n_docs > 0:
        shard_df = (term_ptr[1:] - term_ptr[:-1])
        cand = torch.nonzero(shard_df.double() >= dense_density * n_docs).flatten()
        if cand.numel() > dense_max_terms:
            cand = cand[torch.argsort(shard_df[cand], descending=True)[:dense_max_terms]].sort().values
        if cand.numel() > 0:
            dense_terms = cand.cpu().numpy()
            dense_imp = torch.zeros((cand.numel(), n_docs), dtype=torch.float64, device=device)
            slot = torch.full((n_terms,), -1, dtype=torch.int32, device=device)
            slot[cand] = torch.arange(cand.numel(), dtype=torch.int32, device=device)
            dense_slot = slot
            for c, t in enumerate(dense_terms.tolist()):
                a, z = int(term_ptr[t]), int(term_ptr[t + 1])
                d_t = doc[a:z]
                if imp_table is not None:
                    f_t = imp_table[((post_pack[a:z].long() >> 16) & 0xFFFF)]
                else:
                    f_t = imp[a:z]
                dense_imp[c, d_t] = f_t
    # head terms: fp16(idf * factor) columns of a [n_docs, 64] matrix (cmr_lex_index.head_mat)
    head_mat = head_slot = head_ids = None
    if head_terms and post_pack is not None and total > 0 and n_docs > 0:
        if not 0 < head_terms <= HEAD_TERMS:
            raise ValueError(f"head_terms must be in 0..{HEAD_TERMS}")
        shard_df = (term_ptr[1:] - term_ptr[:-1])
        top = torch.argsort(shard_df, descending=True, stable=True)[:head_terms]
        top = top[shard_df[top] > 0].sort().values
        if top.numel() > 0:
            head_ids = top.cpu().numpy()
            head_mat = torch.zeros((n_docs, HEAD_TERMS), dtype=torch.float16, device=device)
            hs = torch.full((n_terms,), -1, dtype=torch.int32, device=device)
            hs[top] = torch.arange(top.numel(), dtype=torch.int32, device=device)
            head_slot = hs
            for c, t in enumerate(head_ids.tolist()):
                a, z = int(term_ptr[t]), int(term_ptr[t + 1])
                f_t = imp_table[((post_pack[a:z].long() >> 16) & 0xFFFF)]
                head_mat[doc[a:z], c] = (float(idf_host[t]) * f_t).to(torch.float16)
    tf32 = counts.clamp(max=65535).to(torch.int32)
    tf16 = torch.where(tf32 >= 32768, tf32 - 65536, tf32).to(torch.int16)
    return LexicalIndex(
        term_ptr=term_ptr.contiguous(), tile_skip=skip.to(torch.int32).contiguous(),
        post_doc=doc.to(torch.int32).contiguous(), post_imp=None if imp is None else imp.contiguous(),
        post_pack=post_pack, imp_table=imp_table, pair_tf=pair_tf, pair_dl=pair_dl,
        dense_imp=dense_imp, dense_slot=dense_slot, dense_terms=dense_terms,
        head_mat=head_mat, head_slot=head_slot, head_terms=head_ids, skip_row=skip_row,
        post_tf=tf16.contiguous(), doc_len=doc_len.to(torch.int32).contiguous(),
        idf=torch.from_numpy(idf_host).to(device), n_docs=n_docs, n_terms=n_terms, tile_docs=tile_docs,
        n_tiles=n_tiles, avgdl=float(avgdl), k1=k1, b=b, idf_host=idf_host, df_host=stats.df,
        shard_df_host=(term_ptr[1:] - term_ptr[:-1]).cpu().numpy())


def pack_queries(queries: Sequence[Sequence[int]]):
    """[[term ids]] -> (q_terms int32, q_ptr int32 [B+1]) host tensors."""
    ptr = np.zeros(len(queries) + 1, dtype=np.int32)
    np.cumsum(np.fromiter(map(len, queries), dtype=np.int64, count=len(queries)), out=ptr[1:])
    flat = np.fromiter(itertools.chain.from_iterable(queries), dtype=np.int32, count=int(ptr[-1]))
    if flat.size == 0:
        flat = np.zeros(1, dtype=np.int32)  # keep a valid device pointer
    return torch.from_numpy(flat), torch.from_numpy(ptr)


# --------------------------------------------------------------------------
# N2: binary snapshot of a built index (replaces the per-ask JSONL re-parse +
# BM25Okapi rebuild of the reference, rag/retrieval/bm25.py:220-248, rag/pipeline/rag.py:532)
# --------------------------------------------------------------------------
_SNAPSHOT_TENSORS = ("term_ptr", "tile_skip", "post_doc", "post_imp", "post_tf", "doc_len", "idf", "post_pack",
                     "imp_table", "pair_tf", "pair_dl", "dense_imp", "dense_slot", "head_mat", "head_slot", "skip_row")
_SNAPSHOT_VERSION = 1


def save_lexical_index(ix: LexicalIndex, path) -> None:
    """Write the index as one .npy file per array plus meta.json (loadable with mmap)."""
    import json
    from pathlib import Path
    path = Path(path)
    path.mkdir(parents=True, exist_ok=True)
    present = []
    for name in _SNAPSHOT_TENSORS:
        t = getattr(ix, name)
        if t is not None:
            np.save(path / f"{name}.npy", t.detach().cpu().numpy())
            present.append(name)
    for name in ("idf_host", "df_host", "shard_df_host", "dense_terms", "head_terms"):
        a = getattr(ix, name)
        if a is not None:
            np.save(path / f"{name}.npy", np.asarray(a))
            present.append(name)
    meta = {"version": _SNAPSHOT_VERSION, "n_docs": ix.n_docs, "n_terms": ix.n_terms, "tile_docs": ix.tile_docs,
            "n_tiles": ix.n_tiles, "avgdl_hex": float(ix.avgdl).hex(), "k1_hex": float(ix.k1).hex(),
            "b_hex": float(ix.b).hex(), "arrays": present}
    (path / "meta.json").write_text(json.dumps(meta))


def load_lexical_index(path, device) -> LexicalIndex:
    """Inverse of save_lexical_index: arrays go straight to ``device`` (no rebuild)."""
    import json
    from pathlib import Path
    path = Path(path)
    meta = json.loads((path / "meta.json").read_text())
    if meta.get("version") != _SNAPSHOT_VERSION:
        raise ValueError(f"unsupported lexical snapshot version {meta.get('version')}")
    kw = {}
    for name in meta["arrays"]:
        a = np.load(path / f"{name}.npy", mmap_mode="r")
        if name in _SNAPSHOT_TENSORS:
            kw[name] = torch.from_numpy(np.array(a)).to(device)
        else:
            kw[name] = np.array(a)
    for name in _SNAPSHOT_TENSORS:
        kw.setdefault(name, None)
    return LexicalIndex(n_docs=int(meta["n_docs"]), n_terms=int(meta["n_terms"]), tile_docs=int(meta["tile_docs"]),
                        n_tiles=int(meta["n_tiles"]), avgdl=float.fromhex(meta["avgdl_hex"]),
                        k1=float.fromhex(meta["k1_hex"]), b=float.fromhex(meta["b_hex"]), **kw)
