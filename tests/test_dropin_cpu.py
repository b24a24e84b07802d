"""CPU: the host side of the drop-in retrieval package (classmate_rag_b200.retrieval)
against the golden vectors generated from the LIVE reference (tests/golden/make_golden.py):
tokeniser, filter construction, filter flattening + dictionary coding, chunk ids, neighbor
expansion, and the argument errors the reference raises."""
import json

import numpy as np
import pytest

from classmate_rag_b200.retrieval import BM25Store, ChromaVectorStore, HybridRetriever, build_where_filter, rrf_fuse  # noqa: F401
from classmate_rag_b200.retrieval import expand as ex
from classmate_rag_b200.retrieval.filters import MetaColumns, _Present, bm25_clauses, chroma_clauses
from classmate_rag_b200.retrieval.ids import IdRegistry, stable_chunk_id
from classmate_rag_b200.retrieval.text import STOPWORDS_EN, STOPWORDS_IT, detect_lang_tag, tokenize
from oracle import np_oracle as o
from tests.helpers import unhex


def test_import_surface_matches_reference():
    import classmate_rag_b200.retrieval as r
    assert r.__all__ == ["ChromaVectorStore", "build_where_filter", "BM25Store", "rrf_fuse", "HybridRetriever"]
    assert callable(ex.expand_with_neighbors)
    import dataclasses
    names = [f.name for f in dataclasses.fields(HybridRetriever) if not f.name.startswith("_")]
    assert names == ["vector_store", "bm25_store", "embedder", "k_vector", "k_bm25", "rrf_k", "weight_vector",
                     "weight_bm25", "use_mmr", "mmr_lambda", "mmr_max_pool"]
    d = {f.name: f.default for f in dataclasses.fields(HybridRetriever)}
    assert (d["k_vector"], d["k_bm25"], d["rrf_k"], d["use_mmr"], d["mmr_lambda"], d["mmr_max_pool"]) == (8, 8, 60, True, 0.5, 24)


def test_tokenizer_golden(golden):
    assert STOPWORDS_EN == o.STOP_EN and STOPWORDS_IT == o.STOP_IT
    for c in golden["tokenize"]:
        assert tokenize(c["text"], c["lang"]) == c["out"], c


def test_detect_lang_tag_fallback():
    assert detect_lang_tag("") == "en"
    assert detect_lang_tag("the derivative of a function is a limit") == "en"
    assert detect_lang_tag("perché la derivata di una funzione è un limite") in ("it", "en")


def test_build_where_filter_golden(golden):
    for c in golden["build_where_filter"]:
        assert build_where_filter(c["meta_like"]) == c["out"], c


def test_stable_chunk_id_golden(golden):
    for c in golden["stable_chunk_id"]:
        a = c["args"]
        assert stable_chunk_id(source_path=a[0], page=a[1], chunk_index=a[2], course=a[3], unit=a[4]) == c["out"]


def _host_mask(mc: MetaColumns, clauses, n):
    """What cmr_filter_mask computes, from the same columns and codes (test-side restatement)."""
    keep = np.ones(n, dtype=bool)
    for key, want in clauses:
        mc._ensure(key)
        col = mc._host_cols[key]
        if isinstance(want, _Present):
            code = -2 if want.value is None else mc._dicts[key].get(want.value, -2)
        elif want is None:
            code = -1
        else:
            try:
                code = mc._dicts[key].get(want, -2)
            except TypeError:
                code = -2
        keep &= col == code
    return keep


def test_bm25_filter_flattening_matches_reference_matches_filter(golden):
    cases = golden["matches_filter"]
    metas, seen = [], []
    for c in cases:
        if c["meta"] not in seen:
            seen.append(c["meta"])
    metas = seen
    mc = MetaColumns("cpu")
    mc.reset(metas)
    for c in cases:
        got = _host_mask(mc, bm25_clauses(c["where"]), len(metas))[metas.index(c["meta"])]
        assert bool(got) == c["out"], c


def test_chroma_filter_flattening():
    metas = [{"course": "Math101", "tag_exam": True}, {"course": "Phys202"}, {}, {"course": "Math101", "unit": 3}]
    mc = MetaColumns("cpu")
    mc.reset(metas)
    for where in (None, {"course": "Math101"}, {"$and": [{"course": "Math101"}, {"tag_exam": True}]},
                  {"course": {"$eq": "Phys202"}}, {"unit": 3}, {"course": "Nope"}, {"missing": 1}):
        want = [o.chroma_where_matches(m, where if not (where and "course" in where and isinstance(where["course"], dict))
                                       else {"course": where["course"]["$eq"]}) for m in metas]
        assert _host_mask(mc, chroma_clauses(where), len(metas)).tolist() == want, where
    with pytest.raises(ValueError):
        chroma_clauses({"$or": [{"a": 1}]})
    with pytest.raises(ValueError):
        chroma_clauses({"a": {"$gt": 1}})


def test_expand_with_neighbors_golden(golden):
    corpus = golden["corpus"]
    catalog = {i: (d, m) for i, d, m in zip(corpus["ids"], corpus["docs"], corpus["metas"])}
    ex.use_catalog(catalog)
    try:
        for c in golden["expand"]:
            results = [{"id": corpus["ids"][i], "document": corpus["docs"][i], "metadata": corpus["metas"][i],
                        "scores": {"fused": 0.1}} for i in c["seeds"]]
            out = ex.expand_with_neighbors(results, radius=c["radius"], max_per_doc=c["max_per_doc"])
            assert [[x["id"], x["score"]] for x in out] == [[i, unhex(s)] for i, s in c["out"]]
            assert all(set(x) == {"id", "document", "score", "metadata"} for x in out)
    finally:
        ex.use_catalog(None)


def test_expand_reads_the_jsonl_catalog_like_the_reference(tmp_path, golden, monkeypatch):
    corpus = golden["corpus"]
    store = BM25Store(index_dir=tmp_path / "bm25")
    store.upsert_many(ids=corpus["ids"], texts=corpus["docs"], metadatas=corpus["metas"])
    store.save()
    lines = store.index_path.read_text(encoding="utf-8").splitlines()
    assert len(lines) == len(corpus["ids"])
    rec = json.loads(lines[0])
    assert set(rec) == {"id", "text", "tokens", "metadata"} and rec["tokens"] == tokenize(corpus["docs"][0], "en")
    monkeypatch.setattr(ex, "_BM25_JSONL", store.index_path)
    c = golden["expand"][0]
    results = [{"id": corpus["ids"][i], "document": corpus["docs"][i], "metadata": corpus["metas"][i]} for i in c["seeds"]]
    out = ex.expand_with_neighbors(results, radius=c["radius"], max_per_doc=c["max_per_doc"])
    assert [[x["id"], x["score"]] for x in out] == [[i, unhex(s)] for i, s in c["out"]]
    again = BM25Store.load_or_create(tmp_path / "bm25")
    assert again.count() == len(corpus["ids"]) and again._id_list == corpus["ids"]
    assert again.delete_many([corpus["ids"][0], "nope"]) == 1 and again.count() == len(corpus["ids"]) - 1


def test_argument_errors_match_the_reference(tmp_path):
    with pytest.raises(ValueError):
        rrf_fuse(rank_lists=[["a"]], weights=[1.0, 2.0])
    assert rrf_fuse(rank_lists=[]) == {}
    store = BM25Store(index_dir=tmp_path)
    with pytest.raises(ValueError):
        store.upsert_many(ids=["a"], texts=["x", "y"], metadatas=[{}])
    assert store.search(query="   ") == [] and store.search(query="anything") == []
    vs = ChromaVectorStore(persist_dir=tmp_path / "chroma")
    with pytest.raises(ValueError):
        vs.upsert(ids=["a", "b"], documents=["x"], metadatas=[{}, {}], embeddings=np.zeros((2, 8), dtype=np.float32))


def test_id_registry_is_stable():
    r = IdRegistry()
    assert [r.intern(x) for x in ("a", "b", "a", "c")] == [0, 1, 0, 2] and r.name(2) == "c" and len(r) == 3
