"""CPU: the oracle restatement against the golden vectors produced by the live
reference glue (tests/golden/make_golden.py) and hand-computed BM25 answers."""
import math

import numpy as np
import pytest

from oracle import np_oracle as o
from tests.helpers import bits_from_hex, unhex


def test_rrf_bit_exact(golden):
    for c in golden["rrf_fuse"]:
        got = o.rrf_fuse(c["rank_lists"], c["weights"], c["rrf_k"])
        want = [(i, unhex(s)) for i, s in c["out"]]
        assert list(got.items()) == want  # same keys, same insertion order, same bits


def test_rrf_weight_mismatch_raises():
    with pytest.raises(ValueError):
        o.rrf_fuse([["a"]], [1.0, 2.0])
    assert o.rrf_fuse([]) == {}


def test_tokenize(golden):
    for c in golden["tokenize"]:
        assert o.tokenize(c["text"], c["lang"]) == c["out"]


def test_matches_filter(golden):
    for c in golden["matches_filter"]:
        assert o.matches_filter(c["meta"], c["where"]) == c["out"], c


def test_build_where_filter(golden):
    for c in golden["build_where_filter"]:
        assert o.build_where_filter(c["meta_like"]) == c["out"], c


def test_stable_chunk_id(golden):
    for c in golden["stable_chunk_id"]:
        a = c["args"]
        assert o.stable_chunk_id(a[0], a[1], a[2], a[3], a[4]) == c["out"]


def test_mmr_matches_reference_order(golden):
    """The reference multiplies in fp32 BLAS; the oracle pins fp64 exact dots.
    On non-degenerate inputs the greedy order is the same."""
    for c in golden["mmr_order"]:
        q = bits_from_hex(c["q_bits"], (c["d"],))
        cand = bits_from_hex(c["cand_bits"], (c["n"], c["d"]))
        assert o.mmr_order_bf16(q, cand, c["k"], c["lambda"]) == c["out"]


def _corpus(golden):
    c = golden["corpus"]
    emb = bits_from_hex(c["emb_bits"], (c["n"], c["d"]))
    entries = []
    for i, (cid, text, meta) in enumerate(zip(c["ids"], c["docs"], c["metas"])):
        entries.append((cid, o.tokenize(text, meta.get("language")), meta))
    return c, emb, entries


def test_bm25_store_search(golden):
    c, _, entries = _corpus(golden)
    for case in golden["bm25_search"]:
        got = o.bm25_store_search(entries, case["query"], case["where"], case["top_k"])
        want = [(i, unhex(s)) for i, s in case["out"]]
        assert got == want, case["query"]


def test_bm25_csr_equals_dict_form(golden):
    """The vectorised CSR scorer is bit-identical to the rank_bm25-style one."""
    c, _, entries = _corpus(golden)
    docs = [e[1] for e in entries]
    bm = o.BM25Okapi(docs)
    vocab = list(bm.nd.keys())            # first-appearance order
    tid = {w: i for i, w in enumerate(vocab)}
    rows = sorted((tid[w], d, tf) for d, f in enumerate(bm.doc_freqs) for w, tf in f.items())
    term_ptr = np.zeros(len(vocab) + 1, dtype=np.int64)
    for t, _, _ in rows:
        term_ptr[t + 1] += 1
    term_ptr = np.cumsum(term_ptr)
    post_doc = np.array([r[1] for r in rows], dtype=np.int32)
    post_tf = np.array([r[2] for r in rows], dtype=np.int32)
    df = np.diff(term_ptr)
    idf, avg = o.bm25_idf_table(df, len(docs))
    assert avg == bm.average_idf
    for w in vocab:
        assert idf[tid[w]] == bm.idf[w]
    for q in (["gradient", "descent", "kernel"], ["memory", "memory", "bandwidth"], ["nope"], []):
        want = bm.get_scores(q)
        got = o.bm25_scores_csr(term_ptr, post_doc, post_tf, np.array(bm.doc_len), idf, bm.avgdl,
                                [tid.get(w, -1) for w in q])
        assert np.array_equal(got, want)


def test_bm25_known_answers():
    """Hand-computed from the published Okapi formula (independent code path)."""
    docs = [["aa", "bb", "aa"], ["bb", "cc"], ["aa", "bb", "cc", "dd"], ["bb"]]
    bm = o.BM25Okapi(docs)
    n = 4
    avgdl = (3 + 2 + 4 + 1) / 4
    assert bm.avgdl == avgdl
    raw = {w: math.log(n - nd + 0.5) - math.log(nd + 0.5) for w, nd in
           {"aa": 2, "bb": 4, "cc": 2, "dd": 1}.items()}
    avg_idf = (raw["aa"] + raw["bb"] + raw["cc"] + raw["dd"]) / 4
    assert raw["bb"] < 0                      # term in every doc -> negative -> epsilon floor
    assert bm.idf["bb"] == 0.25 * avg_idf
    assert bm.idf["aa"] == raw["aa"] == 0.0   # n_t == N/2 -> idf exactly 0

    def contrib(idf, tf, dl):
        return idf * (tf * 2.5 / (tf + 1.5 * (1 - 0.75 + 0.75 * dl / avgdl)))

    s = bm.get_scores(["dd", "bb", "bb", "zz"])   # repeated + unseen token
    want2 = contrib(raw["dd"], 1, 4) + contrib(bm.idf["bb"], 1, 4) + contrib(bm.idf["bb"], 1, 4)
    assert s[2] == pytest.approx(want2, rel=1e-15)
    assert s[0] == pytest.approx(2 * contrib(bm.idf["bb"], 1, 3), rel=1e-15)
    # zero-score docs are still ranked, in insertion order (bm25.py:199)
    ids, sc = o.bm25_topk(bm.get_scores(["zz"]), 3)
    assert ids.tolist() == [0, 1, 2] and sc.tolist() == [0.0, 0.0, 0.0]
    # empty-corpus guard BM25Okapi([[""]]) (bm25.py:145)
    e = o.BM25Okapi([[""]])
    assert e.get_scores(["x"]).tolist() == [0.0]


def test_exact_dot_order_is_the_documented_one():
    rng = np.random.default_rng(5)
    for d in (8, 32, 40, 264, 768, 1024):
        a = o.f32_to_bf16_bits(rng.standard_normal(d).astype(np.float32))
        b = o.f32_to_bf16_bits(rng.standard_normal(d).astype(np.float32))
        af, bf = o.bf16_bits_to_f64(a), o.bf16_bits_to_f64(b)
        acc = [0.0] * 32
        for v in range(d // 8):            # vector v belongs to lane v % 32
            for e in range(8):
                i = 8 * v + e
                acc[v % 32] = acc[v % 32] + af[i] * bf[i]
        for off in (16, 8, 4, 2, 1):
            for l in range(off):
                acc[l] = acc[l] + acc[l + off]
        assert o.exact_dot_pair(a, b) == acc[0]
        assert abs(acc[0] - float(np.dot(af, bf))) < 1e-12


def test_bf16_rounding_is_rne():
    x = np.array([1.0, 1.00390625, 1.01171875, -2.5, 3.3895314e38, 1e-40, 0.0], dtype=np.float32)
    b = o.f32_to_bf16_bits(x)
    import torch
    want = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(b, want)
    r = np.random.default_rng(0).standard_normal(10000).astype(np.float32)
    want = torch.from_numpy(r).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(o.f32_to_bf16_bits(r), want)


def test_dense_topk_tie_break_and_mask():
    rng = np.random.default_rng(3)
    c = o.f32_to_bf16_bits(rng.standard_normal((200, 64)).astype(np.float32))
    c[150] = c[20]
    c[90] = c[20]
    q = c[20].copy()
    ids, sc = o.dense_topk(q, c, 5)
    assert ids[:3].tolist() == [20, 90, 150] and sc[0] == sc[1] == sc[2]
    mask = np.ones(200, dtype=np.uint8)
    mask[20] = 0
    ids2, _ = o.dense_topk(q, c, 5, mask=mask, row_offset=1000)
    assert ids2[:2].tolist() == [1090, 1150]


def test_retrieve_pipeline_against_reference(golden):
    """Oracle pipeline (exact dense -> MMR -> BM25 subset -> RRF -> final sort)
    against HybridRetriever.retrieve run live on the reference."""
    c, emb, entries = _corpus(golden)
    d = c["d"]
    for case in golden["retrieve"]:
        q = bits_from_hex(case["q_bits"], (d,))
        filt = case["filters"] or {}
        chroma_where = o.build_where_filter(filt) if filt else None
        bm_where = filt or None
        mask = np.array([o.chroma_where_matches(m, chroma_where) for m in c["metas"]], dtype=np.uint8)
        hybrid, use_mmr = case["hybrid"], case["use_mmr"]
        k = 8 if hybrid else max(8, 8)
        pool = max(k, 24) if use_mmr else k
        ids, sc = o.dense_topk(q, emb, pool, mask=mask)
        if use_mmr and len(ids):
            order = o.mmr_order_bf16(q, emb[ids], k, 0.5)
            ids, sc = ids[order], sc[order]
        else:
            ids, sc = ids[:k], sc[:k]
        vec = [(c["ids"][i], 1.0 - s) for i, s in zip(ids.tolist(), sc.tolist())]
        bm = o.bm25_store_search(entries, case["question"], bm_where, 8) if hybrid else []
        got = o.hybrid_merge(vec, bm, top_k=8, hybrid=hybrid)
        want = case["out"]
        assert [g["id"] for g in got] == [w["id"] for w in want], case["question"]
        for g, w in zip(got, want):
            assert g["fused"] == unhex(w["fused"])
            assert g["bm25_score"] == unhex(w["bm25_score"])
            if w["vector_distance"] is None:
                assert g["vector_distance"] is None
            else:
                assert g["vector_distance"] == pytest.approx(unhex(w["vector_distance"]), abs=1e-12)


def test_expand_with_neighbors(golden):
    c = golden["corpus"]
    catalog = {i: (t, m) for i, t, m in zip(c["ids"], c["docs"], c["metas"])}
    for case in golden["expand"]:
        results = [{"id": c["ids"][i], "document": c["docs"][i], "metadata": c["metas"][i],
                    "scores": {"fused": 0.1}} for i in case["seeds"]]
        got = o.expand_with_neighbors(results, catalog, radius=case["radius"], max_per_doc=case["max_per_doc"])
        assert [[g["id"], g["score"]] for g in got] == [[i, unhex(s)] for i, s in case["out"]]


def test_final_sort_tie_rules():
    """equal fused -> BM25-only item precedes a vector item with
    positive distance; distance exactly 0.0 keeps insertion order."""
    got = o.hybrid_merge([("v", 0.25)], [("b", 3.0)], top_k=5)
    assert [g["id"] for g in got] == ["b", "v"]
    got = o.hybrid_merge([("v", 0.0)], [("b", 3.0)], top_k=5)
    assert [g["id"] for g in got] == ["v", "b"]


def test_neardup_keep_first():
    rng = np.random.default_rng(9)
    x = rng.standard_normal((30, 64)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    x[10] = x[3]
    y = x[5] + 0.01 * rng.standard_normal(64).astype(np.float32)
    x[20] = y / np.linalg.norm(y)
    keep = o.neardup_keep_mask(o.f32_to_bf16_bits(x), 0.95)
    assert not keep[10] and not keep[20] and keep.sum() == 28
