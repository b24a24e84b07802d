"""classmate_rag_b200 -- B200-native (sm_100a) hybrid retrieval hot path of
CLASSMATE-RAG behind the reference's ``rag.retrieval`` API.

Only what the path needs: ``csrc/`` (CUDA kernels + C ABI, built into
``libcmrag.so``), ``_lib`` (ctypes binding), ``ops`` (tensor-level calls) and
``retrieval`` (host-side mirror of ``rag/retrieval``).
"""
__version__ = "0.1.0"
