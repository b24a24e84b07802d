"""Lexical store on one B200: the drop-in for the reference's BM25Store
(rag/retrieval/bm25.py:110-256).

Same dataclass fields, method names, keyword-only arguments, JSONL persistence format and
result dicts.  The per-query ``BM25Okapi`` rebuild + Python scoring + full sort of the
reference is replaced by a device-resident inverted index (classmate_rag_b200.lexical) and
``cmr_bm25_topk``; scores are bit-identical to rank_bm25's float64 arithmetic and ties keep
insertion order, so the ranked ids match the reference's stable sort exactly.

Filters follow ``_matches_filter`` to the letter (filters.bm25_clauses).  Like the
reference, a filtered search scores over the statistics of the FILTERED subset (N, df,
avgdl all change, and with them idf and every BM25 factor).  No subset index is built:
``_FilteredView`` derives the subset's statistics with masked reductions over the resident
CSR (df = postings whose document passes, first appearances from the token array), the
idf table and the (tf, doc_len) factor table are recomputed for them, and the exact kernel
scores the FULL index with those tables and the filter's row mask.  A view is a mask and
two small tables, cached per distinct filter until the store changes.
"""
from __future__ import annotations

import dataclasses
import json
from collections import OrderedDict
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Dict, List, Mapping, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .. import _lib, lexical, ops
from .filters import SIMPLE_FIELDS, MetaColumns, bm25_clauses
from .ids import REGISTRY
from .text import detect_lang_tag, tokenize


@dataclass
class _Entry:
    id: str
    text: str
    tokens: List[str]
    metadata: Dict[str, Any]


class _DeviceIndex:
    """Inverted index of a list of entries (the whole store or a filtered subset)."""

    def __init__(self, rows: Sequence[int], doc_ptr: torch.Tensor, tokens: torch.Tensor, n_terms: int,
                 all_doc_ptr_host: np.ndarray, device, lex: Optional["lexical.LexicalIndex"] = None):
        # rows: positions (in store order) of the documents this index covers
        self.rows = np.asarray(rows, dtype=np.int64)
        # the same map on the device (None: the index covers every row of the store, in order)
        self.rows_dev = (None if len(self.rows) == all_doc_ptr_host.shape[0] - 1
                         else torch.from_numpy(self.rows).to(device))
        self.buffers: Dict[Tuple[int, int], ops.TopkBuffers] = {}
        if lex is not None:     # loaded from a snapshot
            self.lex = lex
            return
        if len(rows) == all_doc_ptr_host.shape[0] - 1:
            sub_ptr, sub_tok = doc_ptr, tokens
        else:
            r = torch.from_numpy(self.rows).to(device)
            lens = (doc_ptr[1:] - doc_ptr[:-1])[r]
            sub_ptr = torch.zeros(len(rows) + 1, dtype=torch.int64, device=device)
            sub_ptr[1:] = torch.cumsum(lens, 0)
            total = int(sub_ptr[-1])
            if total:
                doc_of = torch.repeat_interleave(torch.arange(len(rows), device=device), lens)
                pos = torch.arange(total, device=device) - sub_ptr[doc_of] + doc_ptr[r][doc_of]
                sub_tok = tokens[pos]
            else:
                sub_tok = tokens[:0]
        n_docs = len(rows)
        tile = 512
        while tile < 2048 and tile * 4 < n_docs:
            tile *= 2
        self.lex = lexical.build_lexical_index(sub_ptr, sub_tok, max(n_terms, 1), device=device, tile_docs=tile)


class _FilteredView:
    """The full index seen through a filter (reference: the per-query BM25Okapi rebuild over the
    filtered entries, rag/retrieval/bm25.py:184-191).  ``lex`` shares every posting array with
    the full index; only idf, the factor table and avgdl are the subset's.  Dense columns and
    the head matrix hold factors of the FULL corpus's avgdl, so they are dropped: the exact
    kernel walks the packed postings."""

    def __init__(self, full: _DeviceIndex, mask: torch.Tensor, doc_ptr: torch.Tensor, tokens: torch.Tensor):
        lex = full.lex
        dev = mask.device
        self.rows = full.rows
        self.rows_dev = None
        self.buffers: Dict[Tuple[int, int], ops.TopkBuffers] = {}
        self.mask = mask
        import os
        import time
        prof = os.environ.get("CMRAG_PROFILE_FILTER") == "1"
        marks = []

        def mark(name):
            if prof:
                torch.cuda.synchronize(dev)
                marks.append((name, time.perf_counter()))
        mark("start")
        keep = mask.bool()
        n_terms, n_post = lex.n_terms, lex.n_postings
        doc_len = (doc_ptr[1:] - doc_ptr[:-1])
        # df of the subset and every term's first passing posting: one pass over the CSR (cmr_masked_df)
        if n_post < 2 ** 31 - 1 and lex.term_ptr.dtype == torch.int64 and lex.post_doc.dtype == torch.int32:
            df32, first32 = ops.masked_df(lex.term_ptr, lex.post_doc, mask.to(torch.uint8))
            df = df32.to(torch.int64)
            cs = None
        else:   # very large shards: segmented sum with library kernels
            hit = keep.index_select(0, lex.post_doc).to(torch.int64)
            cs = torch.zeros(n_post + 1, dtype=hit.dtype, device=dev)
            torch.cumsum(hit, 0, out=cs[1:])
            before = cs[lex.term_ptr[:-1]]
            df = (cs[lex.term_ptr[1:]] - before).to(torch.int64)
            del hit
        mark("masked df")
        # rank_bm25 sums idf in dict order = first appearance of each term in the subset's token stream:
        # the term's first posting whose document passes (postings are in document order), then its first
        # position inside that document.  V-sized work instead of a pass over every token of the corpus.
        seen = torch.nonzero(df > 0).flatten()
        order = seen
        if seen.numel():
            if cs is None:
                first_post = first32[seen].long()
            else:
                first_post = torch.searchsorted(cs, (before[seen] + 1).contiguous()) - 1   # cs[j + 1] is the first to reach before + 1
            mark("  first postings")
            d_first = lex.post_doc[first_post].long()
            starts, lens = doc_ptr[d_first], doc_len[d_first]
            l_max = int(lens.max())
            pos_in_doc = torch.empty(seen.numel(), dtype=torch.int64, device=dev)
            ar = torch.arange(l_max, device=dev)
            n_tok = int(tokens.numel())
            for lo_c in range(0, seen.numel(), 65536):          # bounded temporaries for large vocabularies
                sl = slice(lo_c, lo_c + 65536)
                idx = (starts[sl, None] + ar[None, :]).clamp_(max=max(n_tok - 1, 0))
                eq = (tokens[idx] == seen[sl, None]) & (ar[None, :] < lens[sl, None])
                pos_in_doc[sl] = torch.where(eq, ar[None, :], l_max).amin(dim=1)
            mark("  positions in documents")
            order = seen[torch.argsort(d_first * (l_max + 1) + pos_in_doc, stable=True)]
        del cs
        mark("first appearances")
        # one read-back: subset size, its token total, df, the vocabulary order
        n_sub_t = torch.stack([keep.sum(), (doc_len * keep).sum()])
        n_sub, total = [int(x) for x in n_sub_t.tolist()]
        self.n_docs = n_sub
        df_h = df.cpu().numpy().astype(np.int64)
        order_h = order.cpu().numpy()
        mark("read-back")
        idf_host, _ = lexical.idf_table(df_h, n_sub, order_h)
        avgdl = (total / n_sub) if n_sub > 0 else 0.0
        imp = (lexical.bm25_factor(lex.pair_tf, lex.pair_dl, avgdl, lex.k1, lex.b) if avgdl > 0
               else torch.zeros_like(lex.imp_table))
        mark("idf + factor table")
        if prof:
            import sys
            print("[_FilteredView] " + ", ".join(f"{b[0]} {1e3 * (b[1] - a[1]):.1f} ms" for a, b in zip(marks, marks[1:])),
                  file=sys.stderr, flush=True)
        self.lex = dataclasses.replace(lex, idf=torch.from_numpy(idf_host).to(dev), imp_table=imp, avgdl=float(avgdl),
                                       dense_imp=None, dense_slot=None, dense_terms=None, head_mat=None, head_slot=None,
                                       head_terms=None, idf_host=idf_host, df_host=df_h, _struct=None)


@dataclass
class BM25Store:
    index_dir: Path = Path("./indexes/bm25")
    index_file: str = "bm25_index.jsonl"
    device: str = "cuda"

    _entries: Dict[str, _Entry] = field(default_factory=dict)   # id -> entry, insertion ordered
    _id_list: List[str] = field(default_factory=list)           # order of the index rows
    _vocab: Dict[str, int] = field(default_factory=dict, repr=False)
    _full: Optional[_DeviceIndex] = field(default=None, repr=False)
    _subsets: "OrderedDict[tuple, Any]" = field(default_factory=OrderedDict, repr=False)   # filter -> _FilteredView
    _columns: Optional[MetaColumns] = field(default=None, repr=False)
    _dirty: bool = field(default=True, repr=False)
    _dev_tokens: Optional[Tuple[torch.Tensor, torch.Tensor, np.ndarray]] = field(default=None, repr=False)
    _gids: Optional[torch.Tensor] = field(default=None, repr=False)
    loaded_from_snapshot: bool = field(default=False, repr=False)
    _snapshot_ok: bool = field(default=False, repr=False)   # entries came straight from load(): the sidecar may describe them
    device_tokenize_from: int = 64          # batches of at least this many queries are tokenised on the device
    _tokenizer: Any = field(default=None, repr=False)

    # ---------- core ops ----------
    def _rebuild(self, from_disk: bool = False) -> None:
        """Reference: rebuild BM25Okapi from the token lists after every mutation.  Here the
        id order is refreshed at once and the device index lazily, on the next search.
        ``from_disk``: the entries are exactly what load() read from the JSONL file -- only
        then may the binary sidecar stand in for a rebuild (an upsert that replaces an id or a
        delete + upsert leaves the file stamp and the entry count unchanged)."""
        self._snapshot_ok = bool(from_disk)
        self._id_list = list(self._entries.keys())
        self._dirty = True
        self._full = None
        self._subsets.clear()
        self.loaded_from_snapshot = False
        self._tokenizer = None

    def _ensure_index(self) -> None:
        if not self._dirty and self._full is not None:
            return
        if not torch.cuda.is_available():
            raise RuntimeError("BM25Store needs a CUDA device: classmate_rag_b200 has no CPU path")
        dev = torch.device(self.device)
        if self._snapshot_ok and self._load_snapshot(dev):
            return
        vocab = self._vocab
        lens = np.zeros(len(self._id_list), dtype=np.int64)
        flat: List[int] = []
        for i, cid in enumerate(self._id_list):
            toks = self._entries[cid].tokens
            lens[i] = len(toks)
            for t in toks:
                n = vocab.get(t)
                if n is None:
                    n = vocab[t] = len(vocab)
                flat.append(n)
        ptr = np.zeros(len(lens) + 1, dtype=np.int64)
        ptr[1:] = np.cumsum(lens)
        doc_ptr = torch.from_numpy(ptr).to(dev)
        tokens = torch.from_numpy(np.asarray(flat, dtype=np.int32)).to(dev)
        self._dev_tokens = (doc_ptr, tokens, ptr)
        self._columns = MetaColumns(dev)
        self._columns.reset([self._entries[c].metadata for c in self._id_list])
        self._columns.prepare([("get", f) for f in SIMPLE_FIELDS])   # dictionary-code the filterable fields now,
        # while the build walks every entry anyway: the first filtered search then pays no O(n) Python loop
        self._gids = torch.tensor([REGISTRY.intern(c) for c in self._id_list], dtype=torch.int64, device=dev)
        self._full = _DeviceIndex(range(len(self._id_list)), doc_ptr, tokens, len(vocab), ptr, dev)
        self._subsets.clear()
        self._dirty = False

    def upsert_many(self, *, ids: Sequence[str], texts: Sequence[str], metadatas: Sequence[Mapping[str, Any]]) -> None:
        """Add or replace documents; tokenised with the metadata language, detected when it is
        missing or "auto" (and then written back into the stored metadata)."""
        if not (len(ids) == len(texts) == len(metadatas)):
            raise ValueError("ids, texts, metadatas must have the same length")
        for i, doc_id in enumerate(ids):
            text = texts[i] or ""
            meta = dict(metadatas[i] or {})
            lang = meta.get("language")
            if not lang or lang == "auto":
                lang = detect_lang_tag(text)
                meta["language"] = lang
            self._entries[doc_id] = _Entry(id=doc_id, text=text, tokens=tokenize(text, lang_hint=lang), metadata=meta)
        self._rebuild()

    def delete_many(self, ids: Sequence[str]) -> int:
        """Returns how many ids were present (the reference returns None although its
        callers report the value as a count, rag/admin/manage.py:191-195)."""
        n = 0
        for doc_id in ids:
            n += self._entries.pop(doc_id, None) is not None
        self._rebuild()
        return n

    def count(self) -> int:
        return len(self._entries)

    # ---------- query ----------
    def _index_for(self, where: Optional[Mapping[str, Any]]) -> Optional[_DeviceIndex]:
        """The index a search with this filter scores against; None = no candidate."""
        self._ensure_index()
        clauses = bm25_clauses(where)
        if not clauses:
            return self._full
        try:
            key = tuple((k, repr(v)) for k, v in clauses)
        except Exception:
            key = None
        if key is not None and key in self._subsets:
            self._subsets.move_to_end(key)
            return self._subsets[key]
        mask = self._columns.mask(clauses)
        doc_ptr, tokens, ptr = self._dev_tokens
        if self._full.lex.post_pack is not None and self._full.lex.pair_tf is not None:
            sub = _FilteredView(self._full, mask, doc_ptr, tokens)      # subset statistics, no rebuild
            if sub.n_docs == 0:
                return None
            if sub.n_docs == len(self._id_list):
                return self._full
        else:   # wide postings (> 65536 distinct (tf, doc_len) pairs): build the subset's own index
            rows = np.nonzero(mask.cpu().numpy().astype(bool))[0]
            if rows.size == 0:
                return None
            if rows.size == len(self._id_list):
                return self._full
            sub = _DeviceIndex(rows, doc_ptr, tokens, len(self._vocab), ptr, torch.device(self.device))
        if key is not None:
            self._subsets[key] = sub
            while len(self._subsets) > 32:
                self._subsets.popitem(last=False)
        return sub

    def _query_terms(self, query: str) -> List[int]:
        q_tokens = tokenize(query, lang_hint=detect_lang_tag(query))
        return [self._vocab.get(t, -1) for t in q_tokens]

    def search_device(self, queries: Sequence[str], where: Optional[Mapping[str, Any]], top_k: int):
        """Batched device-level search: (index, scores f64 [B,k], local docs i64 [B,k], counts
        i32 [B]) on the device, not synchronised; None when no document qualifies."""
        ix = self._index_for(where)
        if ix is None:
            return None
        k = min(int(top_k), _lib.CMR_MAX_K)
        if k <= 0:
            raise ValueError("top_k must be positive")
        dev = torch.device(self.device)
        if len(queries) >= self.device_tokenize_from:
            # large batches: tokenise on the device (the language tag is still decided here)
            if self._tokenizer is None:
                from .device_tokenizer import DeviceTokenizer
                self._tokenizer = DeviceTokenizer(self._vocab, dev)
            longest = max(len(q.encode("utf-8")) for q in queries)
            qt, qp, _ = self._tokenizer(queries, langs=[detect_lang_tag(q) for q in queries],
                                        max_terms=max(8, (longest + 1) // 3 + 1))
        else:
            qt, qp = lexical.pack_queries([self._query_terms(q) for q in queries])
            qt, qp = qt.to(dev), qp.to(dev)
        key = (len(queries), k)
        buf = ix.buffers.get(key)
        if buf is None:
            import ctypes as C
            st = ix.lex.struct()
            with torch.cuda.device(dev):
                nbytes = _lib.load().cmr_bm25_workspace_bytes(C.byref(st), len(queries), k)
            if nbytes == 0:
                raise ValueError("unsupported bm25 shape: " + _lib.last_error())
            buf = ix.buffers[key] = ops.TopkBuffers(len(queries), k, nbytes, dev)
        sc, docs, cnt, _ = ops.bm25_topk(ix.lex, qt, qp, k, buffers=buf, row_mask=getattr(ix, "mask", None))
        return ix, sc, docs, cnt

    def search(self, *, query: str, where: Optional[Mapping[str, Any]] = None, top_k: int = 8) -> List[Dict[str, Any]]:
        """BM25 over the (optionally) filtered subset: id, document, metadata, score (higher
        is better).  Zero-score documents are ranked too, in insertion order."""
        if not query.strip() or not self._entries:
            return []
        res = self.search_batch(queries=[query], where=where, top_k=top_k)
        return res[0]

    def search_batch(self, *, queries: Sequence[str], where: Optional[Mapping[str, Any]] = None,
                     top_k: int = 8) -> List[List[Dict[str, Any]]]:
        """Extension: many queries in one launch (blank queries give [])."""
        if not self._entries:
            return [[] for _ in queries]
        live = [i for i, q in enumerate(queries) if q.strip()]
        out: List[List[Dict[str, Any]]] = [[] for _ in queries]
        if not live:
            return out
        got = self.search_device([queries[i] for i in live], where, top_k)
        if got is None:
            return out
        ix, sc, docs, cnt = got
        sc_h, docs_h, cnt_h = sc.cpu().numpy(), docs.cpu().numpy(), cnt.cpu().numpy()
        for j, qi in enumerate(live):
            items = []
            for r in range(int(cnt_h[j])):
                e = self._entries[self._id_list[int(ix.rows[int(docs_h[j, r])])]]
                items.append({"id": e.id, "document": e.text, "metadata": e.metadata, "score": float(sc_h[j, r])})
            out[qi] = items
        return out

    # ---------- persistence (same JSONL records as the reference) ----------
    @property
    def index_path(self) -> Path:
        return Path(self.index_dir) / self.index_file

    def save(self) -> None:
        Path(self.index_dir).mkdir(parents=True, exist_ok=True)
        with self.index_path.open("w", encoding="utf-8") as f:
            for e in self._entries.values():
                f.write(json.dumps({"id": e.id, "text": e.text, "tokens": e.tokens, "metadata": e.metadata},
                                   ensure_ascii=False) + "\n")
        if self._entries and torch.cuda.is_available():
            self._save_snapshot()

    # Binary sidecar of the device index (N2): the next process loads the arrays instead of
    # re-deriving them from the token lists.  Valid only for exactly this JSONL file.
    @property
    def snapshot_path(self) -> Path:
        return Path(self.index_dir) / (self.index_file + ".cmrag")

    def _jsonl_stamp(self):
        st = self.index_path.stat()
        return [int(st.st_size), int(st.st_mtime_ns)]

    def _save_snapshot(self) -> None:
        self._ensure_index()
        snap = self.snapshot_path
        lexical.save_lexical_index(self._full.lex, snap)
        doc_ptr, tokens, _ = self._dev_tokens
        np.save(snap / "corpus_doc_ptr.npy", doc_ptr.cpu().numpy())
        np.save(snap / "corpus_tokens.npy", tokens.cpu().numpy())
        words = [None] * len(self._vocab)
        for w, n in self._vocab.items():
            words[n] = w
        (snap / "store.json").write_text(json.dumps({"jsonl": self._jsonl_stamp(), "n_entries": len(self._id_list),
                                                     "vocab": words}, ensure_ascii=False), encoding="utf-8")

    def _load_snapshot(self, dev) -> bool:
        snap = self.snapshot_path
        try:
            if not (snap / "store.json").exists() or not self.index_path.exists():
                return False
            info = json.loads((snap / "store.json").read_text(encoding="utf-8"))
            if info.get("jsonl") != self._jsonl_stamp() or info.get("n_entries") != len(self._id_list):
                return False
            lex = lexical.load_lexical_index(snap, dev)
            if lex.n_docs != len(self._id_list):
                return False
            ptr = np.load(snap / "corpus_doc_ptr.npy")
            tokens = torch.from_numpy(np.load(snap / "corpus_tokens.npy")).to(dev)
        except Exception:
            return False
        self._vocab = {w: n for n, w in enumerate(info["vocab"])}
        self._dev_tokens = (torch.from_numpy(ptr).to(dev), tokens, ptr)
        self._columns = MetaColumns(dev)
        self._columns.reset([self._entries[c].metadata for c in self._id_list])
        self._gids = torch.tensor([REGISTRY.intern(c) for c in self._id_list], dtype=torch.int64, device=dev)
        self._full = _DeviceIndex(range(len(self._id_list)), self._dev_tokens[0], tokens, len(self._vocab), ptr, dev,
                                  lex=lex)
        self._subsets.clear()
        self._dirty = False
        self.loaded_from_snapshot = True
        return True

    def load(self) -> None:
        self._entries.clear()
        self._vocab = {}
        if self.index_path.exists():
            with self.index_path.open("r", encoding="utf-8") as f:
                for line in f:
                    if not line.strip():
                        continue
                    rec = json.loads(line)
                    self._entries[rec["id"]] = _Entry(id=rec["id"], text=rec.get("text", ""),
                                                      tokens=list(rec.get("tokens", [])),
                                                      metadata=dict(rec.get("metadata", {})))
        self._rebuild(from_disk=True)

    def catalog(self) -> Dict[str, Tuple[str, Dict[str, Any]]]:
        """id -> (text, metadata): what expand_with_neighbors reads from the JSONL file."""
        return {e.id: (e.text, dict(e.metadata)) for e in self._entries.values()}

    @classmethod
    def load_or_create(cls, index_dir: Union[str, Path] = "./indexes/bm25") -> "BM25Store":
        store = cls(index_dir=Path(index_dir))
        store.load()
        return store
