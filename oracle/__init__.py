"""CPU oracle for the hybrid-retrieval hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker (or as
the timed CPU baseline), never as the thing shipped.  The product package
``classmate_rag_b200`` never imports this package.

Parity status: the reference's own glue (``rrf_fuse``, ``_mmr_order``,
``HybridRetriever.retrieve``, ``_tokenize``, ``_matches_filter``,
``build_where_filter``, ``expand_with_neighbors``, ``stable_chunk_id``) is
pinned by golden vectors generated from the live reference code
(``tests/golden/make_golden.py``).  The two third-party numeric cores are NOT
under ``/root/reference`` and are not installed here -- ``rank_bm25``
(requirements.txt:4, ``>=0.2.2,<0.3``) and ``chromadb``/hnswlib
(requirements.txt:2-3) -- so for those two the restatement follows the
published algorithm and says: **parity unpinned** (no reference test or golden
vector exists for them).
"""
