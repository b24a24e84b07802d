"""Drop-in for ``rag.retrieval`` (reference rag/retrieval/__init__.py:1-11): the same five
names, backed by the B200 kernels of libcmrag.  ``expand_with_neighbors`` lives in
``classmate_rag_b200.retrieval.expand`` exactly as it does in the reference."""
from .vector_store import ChromaVectorStore
from .filters import build_where_filter
from .bm25_store import BM25Store
from .fusion import rrf_fuse, HybridRetriever

__all__ = [
    "ChromaVectorStore",
    "build_where_filter",
    "BM25Store",
    "rrf_fuse",
    "HybridRetriever",
]
