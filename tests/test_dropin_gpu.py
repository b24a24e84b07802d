"""GPU parity of the drop-in retrieval classes (classmate_rag_b200.retrieval) against the
golden vectors produced by the LIVE reference code (tests/golden/make_golden.py):
BM25Store.search, HybridRetriever.retrieve, rrf_fuse, the filters, neighbor expansion.

The reference's third-party numeric cores were substituted when the vectors were made
(rank_bm25 -> oracle restatement, Chroma/hnswlib -> exact brute-force cosine); everything
else in the expected outputs is the reference's own code."""
import math

import numpy as np
import pytest
import torch

from tests.helpers import bits_from_hex, unhex
from oracle import np_oracle as o

pytestmark = pytest.mark.gpu


class _Emb:
    """Embedder stub with the E5 contract: encode_queries(list[str]) -> float32 [B, dim]."""

    def __init__(self):
        self.vec = None

    def encode_queries(self, texts):
        return np.stack([self.vec for _ in texts])


@pytest.fixture(scope="module")
def world(golden, tmp_path_factory):
    from classmate_rag_b200.retrieval import BM25Store, ChromaVectorStore
    c = golden["corpus"]
    td = tmp_path_factory.mktemp("dropin")
    emb_bits = bits_from_hex(c["emb_bits"], (c["n"], c["d"]))
    emb_f32 = o.bf16_bits_to_f32(emb_bits)
    store = BM25Store(index_dir=td / "bm25")
    store.upsert_many(ids=c["ids"], texts=c["docs"], metadatas=c["metas"])
    vs = ChromaVectorStore(persist_dir=td / "chroma", collection_name="golden")
    vs.upsert(ids=c["ids"], documents=c["docs"], metadatas=c["metas"], embeddings=emb_f32)
    return store, vs, c, emb_bits, emb_f32, td


def test_bm25store_search_golden(world, golden):
    store = world[0]
    for case in golden["bm25_search"]:
        res = store.search(query=case["query"], where=case["where"], top_k=case["top_k"])
        assert [[x["id"], x["score"]] for x in res] == [[i, unhex(s)] for i, s in case["out"]], case["query"]
        for x in res:
            assert set(x) == {"id", "document", "metadata", "score"} and isinstance(x["score"], float)


def test_bm25store_search_batch_equals_single(world, golden):
    store = world[0]
    queries = sorted({c["query"] for c in golden["bm25_search"]})
    for where in (None, {"course": "Math101"}):
        batch = store.search_batch(queries=queries, where=where, top_k=8)
        for q, got in zip(queries, batch):
            assert got == store.search(query=q, where=where, top_k=8)


def test_hybrid_retriever_golden(world, golden):
    from classmate_rag_b200.retrieval import HybridRetriever
    store, vs, c, emb_bits, emb_f32, _ = world
    emb = _Emb()
    for case in golden["retrieve"]:
        emb.vec = o.bf16_bits_to_f32(bits_from_hex(case["q_bits"], (c["d"],)))
        hr = HybridRetriever(vector_store=vs, bm25_store=store, embedder=emb, k_vector=8, k_bm25=8,
                             use_mmr=case["use_mmr"])
        out = hr.retrieve(question=case["question"], filters=case["filters"], top_k=8, hybrid=case["hybrid"])
        assert [x["id"] for x in out] == [w["id"] for w in case["out"]], case
        for x, w in zip(out, case["out"]):
            assert set(x) == {"id", "document", "metadata", "scores"}
            s = x["scores"]
            assert s["fused"] == unhex(w["fused"])                      # RRF: bit-exact
            assert s["bm25_score"] == unhex(w["bm25_score"])            # BM25: bit-exact
            wd = unhex(w["vector_distance"])
            if wd is None:
                assert s["vector_distance"] is None
            else:  # same bf16 inputs, float64 sums in a different order than NumPy's matmul
                assert math.isclose(s["vector_distance"], wd, rel_tol=0, abs_tol=1e-12)
            i = c["ids"].index(x["id"])
            assert x["metadata"] == c["metas"][i]
            assert x["document"] == (c["docs"][i] or (None if s["vector_distance"] is None else c["docs"][i]))


def test_rrf_fuse_golden(golden):
    from classmate_rag_b200.retrieval import rrf_fuse
    for case in golden["rrf_fuse"]:
        out = rrf_fuse(rank_lists=case["rank_lists"], weights=case["weights"], rrf_k=case["rrf_k"])
        assert list(out.items()) == [(i, unhex(s)) for i, s in case["out"]]


def test_device_filter_masks_golden(golden):
    from classmate_rag_b200.retrieval.filters import MetaColumns, bm25_clauses
    metas = []
    for case in golden["matches_filter"]:
        if case["meta"] not in metas:
            metas.append(case["meta"])
    mc = MetaColumns(torch.device("cuda"))
    mc.reset(metas)
    for case in golden["matches_filter"]:
        got = mc.mask(bm25_clauses(case["where"])).cpu().numpy()[metas.index(case["meta"])]
        assert bool(got) == case["out"], case


def test_vector_store_behaviour(world, golden):
    from classmate_rag_b200.retrieval import ChromaVectorStore, build_where_filter
    store, vs, c, emb_bits, emb_f32, td = world
    assert vs.count() == c["n"]
    q = emb_f32[5]
    res = vs.query(query_embeddings=q, top_k=5, include_embeddings=True)
    want_ids, want_sc = o.dense_topk(emb_bits[5], emb_bits, 5)
    assert [r["id"] for r in res] == [c["ids"][i] for i in want_ids]
    assert [r["distance"] for r in res] == [1.0 - float(s) for s in want_sc]
    assert res[0]["id"] == c["ids"][5] and np.array_equal(res[0]["embedding"], emb_f32[5])
    assert set(res[0]) == {"id", "document", "metadata", "distance", "embedding"}
    # 2-D input: only the first query's hits come back (reference behaviour); batch API returns all
    two = np.stack([emb_f32[5], emb_f32[9]])
    assert [r["id"] for r in vs.query(query_embeddings=two, top_k=3)] == [r["id"] for r in res[:3]]
    both = vs.query_batch(query_embeddings=two, top_k=3)
    assert len(both) == 2 and both[1][0]["id"] == c["ids"][9]
    # where filter (Chroma-style) == oracle mask
    where = build_where_filter({"course": "Phys202", "tags": ["exam"]})
    keep = np.array([o.chroma_where_matches(m, where) for m in c["metas"]], dtype=np.uint8)
    got = vs.query(query_embeddings=q, where=where, top_k=6)
    want_ids, _ = o.dense_topk(emb_bits[5], emb_bits, 6, mask=keep)
    assert [r["id"] for r in got] == [c["ids"][i] for i in want_ids]
    assert vs.query(query_embeddings=q, where={"course": "Nope"}, top_k=6) == []
    # a second handle on the same collection sees the same rows; upsert moves an id to the end
    vs2 = ChromaVectorStore(persist_dir=td / "chroma", collection_name="golden")
    assert vs2.count() == c["n"]
    side = ChromaVectorStore(persist_dir=td / "side", collection_name="t")
    side.upsert(ids=["a", "b", "c"], documents=["A", "B", "C"], metadatas=[{}, {}, {}], embeddings=emb_f32[:3])
    side.upsert(ids=["a"], documents=["A2"], metadatas=[{"v": 2}], embeddings=emb_f32[3:4])
    assert side.count() == 3
    r = side.query(query_embeddings=emb_f32[3], top_k=3)
    assert r[0]["id"] == "a" and r[0]["document"] == "A2" and r[0]["metadata"] == {"v": 2}
    assert side.delete(["b", "zzz"]) == 1 and side.count() == 2
    assert [x["id"] for x in side.query(query_embeddings=emb_f32[1], top_k=3)] != [] and \
        "b" not in [x["id"] for x in side.query(query_embeddings=emb_f32[1], top_k=3)]
    # durable without persist(), like chromadb.PersistentClient: a new process (simulated by dropping the live
    # collection) replays the operation log -- upserts, the re-upsert that moved "a", the delete
    from classmate_rag_b200.retrieval import vector_store as vsm
    before = side.query(query_embeddings=emb_f32[1], top_k=3)
    vsm._COLLECTIONS.pop(side._key())
    replay = ChromaVectorStore(persist_dir=td / "side", collection_name="t")
    assert replay.count() == 2 and replay.query(query_embeddings=emb_f32[1], top_k=3) == before
    assert replay.query(query_embeddings=emb_f32[3], top_k=1)[0]["document"] == "A2"
    snap = replay.persist()             # compacted rewrite (atomic rename); 2 rows, no tombstones
    assert (snap / "rows.bf16").stat().st_size == 2 * emb_f32.shape[1] * 2
    assert not snap.with_name(snap.name + ".tmp").exists() and not snap.with_name(snap.name + ".old").exists()
    vsm._COLLECTIONS.pop(side._key())
    again = ChromaVectorStore(persist_dir=td / "side", collection_name="t")
    assert again.count() == 2 and again.query(query_embeddings=emb_f32[3], top_k=1)[0]["document"] == "A2"
    assert again.query(query_embeddings=emb_f32[1], top_k=3) == before
    # rows that are not unit length are normalised on the way in, like hnswlib's cosine space
    again.upsert(ids=["long"], documents=["L"], metadatas=[{}], embeddings=emb_f32[7:8] * 5.0)
    hit = again.query(query_embeddings=emb_f32[7] * 3.0, top_k=1)[0]
    assert hit["id"] == "long" and abs(hit["distance"]) < 1e-2
    again.reset_collection()
    assert again.count() == 0 and again.query(query_embeddings=emb_f32[3], top_k=1) == []


def test_expand_uses_store_catalog(world, golden):
    from classmate_rag_b200.retrieval import expand as ex
    store, vs, c = world[0], world[1], world[2]
    ex.use_catalog(store)
    try:
        for case in golden["expand"]:
            results = [{"id": c["ids"][i], "document": c["docs"][i], "metadata": c["metas"][i]} for i in case["seeds"]]
            out = ex.expand_with_neighbors(results, radius=case["radius"], max_per_doc=case["max_per_doc"])
            assert [[x["id"], x["score"]] for x in out] == [[i, unhex(s)] for i, s in case["out"]]
    finally:
        ex.use_catalog(None)


def test_bm25store_snapshot_reload_gives_identical_results(world, golden, tmp_path):
    """N2: save() leaves a binary sidecar of the device index; a new process-like store loads it
    (no rebuild) and answers exactly as before; a changed JSONL invalidates it."""
    from classmate_rag_b200.retrieval import BM25Store
    c = world[2]
    a = BM25Store(index_dir=tmp_path / "bm25")
    a.upsert_many(ids=c["ids"], texts=c["docs"], metadatas=c["metas"])
    a.save()
    assert (a.snapshot_path / "meta.json").exists()
    b = BM25Store.load_or_create(tmp_path / "bm25")
    cases = golden["bm25_search"][:40]
    for case in cases:
        res = b.search(query=case["query"], where=case["where"], top_k=case["top_k"])
        assert [[x["id"], x["score"]] for x in res] == [[i, unhex(s)] for i, s in case["out"]]
    assert b.loaded_from_snapshot
    b.upsert_many(ids=["new"], texts=["gradient descent gradient"], metadatas=[{"language": "en"}])
    assert not b.loaded_from_snapshot and b.search(query="gradient", top_k=100)[0]["id"] is not None
    # a JSONL that no longer matches the sidecar: rebuilt from the token lists
    with a.index_path.open("a", encoding="utf-8") as f:
        f.write('{"id": "extra", "text": "kernel kernel", "tokens": ["kernel", "kernel"], "metadata": {"language": "en"}}\n')
    d = BM25Store.load_or_create(tmp_path / "bm25")
    assert d.count() == len(c["ids"]) + 1
    assert d.search(query="kernel", top_k=3)[0]["id"] == "extra" and not d.loaded_from_snapshot


def test_bm25store_snapshot_is_not_used_after_a_mutation(tmp_path):
    """The sidecar is validated by the JSONL stamp and the entry count only, so it may stand in for a rebuild
    only while the entries are exactly what load() read: replacing an id (count unchanged, file untouched) or
    filling a fresh store over a directory that holds an equally long snapshot must rebuild."""
    from classmate_rag_b200.retrieval import BM25Store
    ids = [f"d{i}" for i in range(6)]
    texts = ["alpha beta", "beta gamma", "gamma delta", "delta alpha", "alpha alpha gamma", "beta delta delta"]
    metas = [{"language": "en"}] * 6
    a = BM25Store(index_dir=tmp_path / "bm25")
    a.upsert_many(ids=ids, texts=texts, metadatas=metas)
    a.save()
    b = BM25Store.load_or_create(tmp_path / "bm25")
    b.upsert_many(ids=["d0"], texts=["zeta zeta zeta"], metadatas=[{"language": "en"}])   # same count, same file
    hit = b.search(query="zeta", top_k=1)
    assert hit[0]["id"] == "d0" and hit[0]["score"] > 0 and not b.loaded_from_snapshot
    assert all(x["id"] != "d0" or x["score"] == 0.0 for x in b.search(query="alpha", top_k=6))
    b.save()
    c = BM25Store.load_or_create(tmp_path / "bm25")
    hit = c.search(query="zeta", top_k=1)
    assert hit[0]["id"] == "d0" and hit[0]["score"] > 0 and c.loaded_from_snapshot
    # a fresh store over the same directory, populated with as many (other) documents
    d = BM25Store(index_dir=tmp_path / "bm25")
    d.upsert_many(ids=ids, texts=["omega"] * 6, metadatas=metas)
    assert d.search(query="omega", top_k=1)[0]["id"] == "d0" and not d.loaded_from_snapshot
    assert d.search(query="zeta", top_k=1)[0]["score"] == 0.0


def test_device_tokenizer_matches_host_tokenizer(golden):
    """N3: cmr_tokenize_queries == tokenize() + vocabulary lookup, for both stopword lists,
    accented letters, the two excluded signs, other scripts, empty and truncated queries."""
    from classmate_rag_b200.retrieval.device_tokenizer import DeviceTokenizer
    from classmate_rag_b200.retrieval.text import tokenize
    texts = sorted({c["text"] for c in golden["tokenize"]}) + [
        "ÀÉÎÕÜ Þorn ßtraße ÿes ×no÷ yes×no", "日本語 text mixed 한국어 and emoji 🙂 ok", "école école ÉCOLE",
        "a b c I x", "come dove quando the of and perché PERCHÉ", "x" * 300 + " tail", "",
        "gradient descent kernel memory bandwidth cache latency pipeline fusion retrieval ranking lexical dense"]
    rng = np.random.default_rng(0)
    alphabet = list("abcXYZ éÈñ×÷ß-_'1 ") + ["ÿ", "À", "Ā", "€"]
    texts += ["".join(rng.choice(alphabet, size=int(rng.integers(0, 60)))) for _ in range(300)]
    words = sorted({t for x in texts for t in tokenize(x, "en")} | {t for x in texts for t in tokenize(x, "it")})
    vocab = {w: i for i, w in enumerate(words) if i % 3 != 0}          # a third of the words stays unknown
    vocab["come"] = len(words) + 1                                     # a term that is an Italian stopword
    tok = DeviceTokenizer(vocab, "cuda")
    for lang in ("en", "it", None):
        langs = None if lang is None else [lang] * len(texts)
        q_terms, q_ptr, counts = tok(texts, langs=langs, max_terms=24)
        q_terms, q_ptr, counts = q_terms.cpu().numpy(), q_ptr.cpu().numpy(), counts.cpu().numpy()
        assert q_ptr.tolist() == [24 * i for i in range(len(texts) + 1)]
        for i, x in enumerate(texts):
            want = [vocab.get(t, -1) for t in tokenize(x, lang)]
            assert counts[i] == len(want), (x, lang)
            row = q_terms[24 * i: 24 * (i + 1)].tolist()
            assert row[: min(len(want), 24)] == want[:24], (x, lang)
            assert all(v == -1 for v in row[len(want):])


def test_bm25store_batch_on_device_tokenizer_equals_host_path(world, golden):
    store = world[0]
    queries = [c["query"] for c in golden["bm25_search"]][:80] + ["Gradient DESCENT kernel", "memória ×bandwidth÷"]
    want = [store.search(query=q, top_k=8) for q in queries]
    store.device_tokenize_from = 1
    try:
        got = store.search_batch(queries=queries, top_k=8)
    finally:
        store.device_tokenize_from = 64
    assert got == want


def test_dropin_edge_cases(tmp_path, golden):
    """Empty stores, non-hybrid, no MMR, top_k beyond the corpus, filters that match nothing."""
    from classmate_rag_b200.retrieval import BM25Store, ChromaVectorStore, HybridRetriever
    c = golden["corpus"]
    emb_f32 = o.bf16_bits_to_f32(bits_from_hex(c["emb_bits"], (c["n"], c["d"])))
    emb = _Emb()
    emb.vec = emb_f32[3]
    vs = ChromaVectorStore(persist_dir=tmp_path / "chroma", collection_name="edge")
    bm = BM25Store(index_dir=tmp_path / "bm25")
    hr = HybridRetriever(vector_store=vs, bm25_store=bm, embedder=emb)
    assert hr.retrieve(question="gradient", top_k=8) == []                      # both stores empty
    assert vs.query(query_embeddings=emb_f32[0], top_k=3) == [] and vs.count() == 0
    bm.upsert_many(ids=c["ids"][:10], texts=c["docs"][:10], metadatas=c["metas"][:10])
    only_bm = hr.retrieve(question="gradient descent kernel", top_k=8)          # vector store still empty
    assert only_bm and all(x["scores"]["vector_distance"] is None for x in only_bm)
    assert [x["id"] for x in only_bm] == [x["id"] for x in bm.search(query="gradient descent kernel", top_k=8)]
    vs.upsert(ids=c["ids"][:10], documents=c["docs"][:10], metadatas=c["metas"][:10], embeddings=emb_f32[:10])
    out = hr.retrieve(question="gradient descent kernel", top_k=50)             # top_k beyond what exists
    assert 0 < len(out) <= 16 and out[0]["scores"]["fused"] >= out[-1]["scores"]["fused"]
    dense_only = HybridRetriever(vector_store=vs, bm25_store=bm, embedder=emb, use_mmr=False).retrieve(
        question="anything", top_k=4, hybrid=False)
    assert [x["id"] for x in dense_only] == [x["id"] for x in vs.query(query_embeddings=emb.vec, top_k=8)][:4]
    assert dense_only[0]["id"] == c["ids"][3] and all(x["scores"]["bm25_score"] is None for x in dense_only)
    assert hr.retrieve(question="gradient", filters={"course": "NoSuchCourse"}, top_k=8) == []
    with pytest.raises(ValueError):
        vs.query(query_embeddings=np.zeros(c["d"] * 2, dtype=np.float32), top_k=3)   # wrong dimension


def test_vector_store_dedupe_hook(tmp_path):
    """`rag rebuild` hook: near-duplicate rows (cos >= 0.95 to an earlier kept row) are removed."""
    from classmate_rag_b200.retrieval import ChromaVectorStore
    rng = np.random.default_rng(4)
    n, d = 600, 64
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    dup_of = {50: 10, 51: 50, 300: 299, 599: 0}
    for i, j in dup_of.items():
        v = x[j] + 0.02 * rng.standard_normal(d).astype(np.float32) / np.sqrt(d)
        x[i] = v / np.linalg.norm(v)
    x = o.bf16_bits_to_f32(o.f32_to_bf16_bits(x))
    ids = [f"c{i}" for i in range(n)]
    vs = ChromaVectorStore(persist_dir=tmp_path / "chroma", collection_name="dd")
    vs.upsert(ids=ids, documents=ids, metadatas=[{} for _ in ids], embeddings=x)
    want_keep = o.neardup_keep_mask(o.f32_to_bf16_bits(x), 0.95)
    dropped = vs.dedupe(0.95)
    assert dropped == [ids[i] for i in range(n) if not want_keep[i]] == ["c50", "c51", "c300", "c599"]
    assert vs.count() == n - 4 and vs.dedupe(0.95) == []
    assert vs.query(query_embeddings=x[50], top_k=1)[0]["id"] == "c10"
