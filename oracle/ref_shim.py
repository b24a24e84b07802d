"""Import the LIVE reference glue from /root/reference with stub third-party
modules.  TEST INFRASTRUCTURE -- only usable in the build container (the GPU
box has no /root/reference); used by ``tests/golden/make_golden.py`` to
generate the committed golden vectors and by the optional live differential
tests (skipped when the reference tree is absent).

Stubs injected:
  * ``rank_bm25.BM25Okapi``      -> oracle.np_oracle.BM25Okapi (our restatement)
  * ``sentence_transformers``    -> a class that is never instantiated
  * ``langdetect``               -> detect() == "en", DetectorFactory
  * ``chromadb``                 -> an in-memory exact brute-force cosine
                                    collection (stand-in for hnswlib)
"""
from __future__ import annotations

import importlib
import sys
import types
from pathlib import Path
from typing import Any, Dict, List

import numpy as np

REFERENCE_ROOT = Path("/root/reference")


def available() -> bool:
    return (REFERENCE_ROOT / "rag" / "retrieval" / "fusion.py").exists()


class _FakeCollection:
    """Exact cosine top-k in float32, ties by insertion order."""

    def __init__(self):
        self.ids: List[str] = []
        self.docs: List[str] = []
        self.metas: List[Dict[str, Any]] = []
        self.embs: List[np.ndarray] = []

    def delete(self, ids):
        drop = set(ids)
        keep = [i for i, x in enumerate(self.ids) if x not in drop]
        self.ids = [self.ids[i] for i in keep]
        self.docs = [self.docs[i] for i in keep]
        self.metas = [self.metas[i] for i in keep]
        self.embs = [self.embs[i] for i in keep]

    def add(self, ids, documents, metadatas, embeddings):
        for i, d, m, e in zip(ids, documents, metadatas, embeddings):
            self.ids.append(i)
            self.docs.append(d)
            self.metas.append(dict(m))
            self.embs.append(np.asarray(e, dtype=np.float32))

    def count(self):
        return len(self.ids)

    def query(self, query_embeddings, n_results, include, where=None):
        from oracle.np_oracle import chroma_where_matches
        out = {"ids": [], "documents": [], "metadatas": [], "distances": [], "embeddings": []}
        for q in query_embeddings:
            q = np.asarray(q, dtype=np.float32)
            rows = [i for i in range(len(self.ids)) if chroma_where_matches(self.metas[i], where)]
            if rows:
                mat = np.stack([self.embs[i] for i in rows]).astype(np.float64)
                sims = mat @ q.astype(np.float64)
                dist = 1.0 - sims
                order = np.lexsort((np.arange(len(rows)), dist))[:n_results]
            else:
                dist, order = np.zeros(0), []
            sel = [rows[i] for i in order]
            out["ids"].append([self.ids[i] for i in sel])
            out["documents"].append([self.docs[i] for i in sel])
            out["metadatas"].append([self.metas[i] for i in sel])
            out["distances"].append([float(dist[i]) for i in order])
            out["embeddings"].append([self.embs[i].tolist() for i in sel])
        return out


class _FakeClient:
    _collections: Dict[str, _FakeCollection] = {}

    def __init__(self, *a, **k):
        pass

    def get_or_create_collection(self, name, metadata=None, embedding_function=None):
        return self._collections.setdefault(name, _FakeCollection())

    def delete_collection(self, name):
        self._collections.pop(name, None)


def install_stubs() -> None:
    from oracle import np_oracle

    rb = types.ModuleType("rank_bm25")
    rb.BM25Okapi = np_oracle.BM25Okapi
    sys.modules.setdefault("rank_bm25", rb)

    st = types.ModuleType("sentence_transformers")

    class SentenceTransformer:  # never instantiated by the retrieval glue
        def __init__(self, *a, **k):
            raise RuntimeError("stub")

    st.SentenceTransformer = SentenceTransformer
    sys.modules.setdefault("sentence_transformers", st)

    ld = types.ModuleType("langdetect")
    ld.detect = lambda text: "en"

    class DetectorFactory:
        seed = 0

    ld.DetectorFactory = DetectorFactory
    sys.modules.setdefault("langdetect", ld)

    ch = types.ModuleType("chromadb")
    ch.PersistentClient = _FakeClient
    ch.HttpClient = _FakeClient
    sys.modules.setdefault("chromadb", ch)


def load():
    """Return a namespace with the reference's retrieval modules imported."""
    if not available():
        raise RuntimeError("reference tree not present")
    install_stubs()
    root = str(REFERENCE_ROOT)
    if root not in sys.path:
        sys.path.insert(0, root)
    ns = types.SimpleNamespace()
    ns.fusion = importlib.import_module("rag.retrieval.fusion")
    ns.bm25 = importlib.import_module("rag.retrieval.bm25")
    ns.vector_chroma = importlib.import_module("rag.retrieval.vector_chroma")
    ns.expand = importlib.import_module("rag.retrieval.expand")
    ns.ids = importlib.import_module("rag.utils.ids")
    return ns
