#!/usr/bin/env python
"""Generate tests/golden/reference_glue.json by running the LIVE reference glue
(/root/reference/rag/retrieval/*, rag/utils/ids.py) through oracle/ref_shim.py.

Run in the build container only:  python tests/golden/make_golden.py
The reference tree cannot travel to the GPU box, so the vectors are committed.

What is and is not pinned:
  * pinned by the reference's own code: rrf_fuse, _mmr_order, _tokenize,
    _matches_filter, build_where_filter, stable_chunk_id, expand_with_neighbors,
    HybridRetriever.retrieve (merge + final sort), BM25Store.search control
    flow (filter -> subset rebuild -> stable sort, zeros included).
  * NOT pinned (third-party, absent): rank_bm25 arithmetic and Chroma/hnswlib
    distances; the shim substitutes oracle.np_oracle.BM25Okapi and an exact
    brute-force cosine collection.
Floats are stored as float.hex() strings so comparisons can be bit-exact.
"""
from __future__ import annotations

import json
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import np_oracle as o  # noqa: E402
from oracle import ref_shim  # noqa: E402

WORDS = ("gradient descent matrix vector tensor kernel memory bandwidth cache latency "
         "pipeline fusion retrieval ranking lexical dense sparse token shard stripe "
         "posting index query chunk embedding cosine neighbor lecture exam course unit "
         "integral derivative theorem proof lemma corollary entropy compiler parser "
         "lattice algebra topology manifold quantum photon").split()


def hexf(x):
    return float(x).hex()


def bits_hex(a: np.ndarray) -> str:
    return np.ascontiguousarray(a, dtype=np.uint16).tobytes().hex()


def synth_docs(rng, n, lo=4, hi=18):
    docs = []
    for _ in range(n):
        ln = int(rng.integers(lo, hi))
        # zipf-ish pick
        idx = np.minimum((rng.pareto(1.1, ln)).astype(int), len(WORDS) - 1)
        docs.append(" ".join(WORDS[i] for i in idx))
    return docs


def unit_bf16(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    bits = o.f32_to_bf16_bits(x)
    return bits, o.bf16_bits_to_f32(bits)


def main():
    ref = ref_shim.load()
    g = {}

    # stopword tables identical
    assert ref.bm25._STOP_EN == set(o.STOP_EN) and ref.bm25._STOP_IT == set(o.STOP_IT)

    # ---- rrf_fuse ---------------------------------------------------------
    cases = []
    rng = np.random.default_rng(11)
    raw = [
        ([["a", "b", "c"], ["b", "d"]], None, 60),
        ([["a", "b", "c"], ["b", "d"]], [1.0, 1.0], 60),
        ([["x1", "x2"], ["x2", "x1"], ["x3"]], [0.3, 1.7, 2.0], 10),
        ([[], ["only"]], [1.0, 0.5], 60),
        ([[f"d{i}" for i in rng.permutation(100)], [f"d{i}" for i in rng.permutation(100)]], [1.0, 1.0], 60),
        ([[f"d{i}" for i in range(8)], [f"d{i}" for i in range(4, 12)]], [0.7, 1.3], 1),
    ]
    for lists, w, k in raw:
        out = ref.fusion.rrf_fuse(rank_lists=lists, weights=w, rrf_k=k)
        cases.append({"rank_lists": lists, "weights": w, "rrf_k": k,
                      "out": [[i, hexf(s)] for i, s in out.items()]})
    g["rrf_fuse"] = cases
    try:
        ref.fusion.rrf_fuse(rank_lists=[["a"]], weights=[1.0, 2.0])
        raise AssertionError("expected ValueError")
    except ValueError:
        pass

    # ---- tokenizer ----------------------------------------------------------
    texts = [
        "The Chain-rule's été l'Hôpital 3x2 a b2b",
        "Perché la derivata di una funzione è un limite? Non lo so.",
        "",
        "   ",
        "A I x YY zz_ww THE the The tHe",
        "naïve café Ærø øl ÿ÷× Øre ×times÷",
        "snake_case camelCase kebab-case 123abc456",
    ]
    g["tokenize"] = [{"text": t, "lang": l, "out": ref.bm25._tokenize(t, l)}
                     for t in texts for l in (None, "en", "it", "EN-us", "fr")]

    # ---- filters ------------------------------------------------------------
    metas = [
        {"course": "Math101", "unit": "1", "language": "en", "doc_type": "slides", "tags": ["exam", "week1"]},
        {"course": "Math101", "language": "it", "tags": []},
        {"language": "en"},
        {},
        {"course": None, "author": "Rossi", "semester": "2024S", "tags": ["exam"]},
    ]
    wheres = [
        None, {}, {"course": "Math101"}, {"course": "Math101", "language": "en"},
        {"course": None}, {"author": None, "course": None},
        {"tags": {"$contains": "exam"}}, {"tags": {"$contains": ["exam", "week1"]}},
        {"tags": {"$contains": ""}}, {"tags": ["exam"]},
        {"$and": [{"course": "Math101"}, {"language": "it"}]},
        {"$and": [{"tags": {"$contains": "exam"}}, {"semester": "2024S"}]},
        {"unknown_field": "zzz"},
    ]
    g["matches_filter"] = [{"meta": m, "where": w, "out": bool(ref.bm25._matches_filter(m, w))}
                           for m in metas for w in wheres]
    likes = [
        {}, {"course": "Math101"}, {"course": " Math101 ", "unit": "", "doc_type": "other"},
        {"doc_type": "Other", "language": "it"}, {"tags": ["Exam Prep", "week-1", " "]},
        {"tags": "a, b ,,c"}, {"course": None, "unit": None, "author": None},
        {"course": "C", "unit": "U", "language": "en", "doc_type": "pdf", "author": "A", "semester": "S", "tags": ["t"]},
        {"semester": 2024},
    ]
    g["build_where_filter"] = [{"meta_like": m, "out": ref.vector_chroma.build_where_filter(m)} for m in likes]

    # ---- stable_chunk_id ---------------------------------------------------
    idc = [("/data/course/lec1.pdf", 1, 0, None, None), ("/data/course/lec1.pdf", 3, 17, "Math101", "U2"),
           ("/tmp/ünï.txt", 0, 5, "", "x"), ("/a/b/../c.md", 2, 2, "K", None)]
    g["stable_chunk_id"] = [{"args": list(a), "out": ref.ids.stable_chunk_id(
        source_path=a[0], page=a[1], chunk_index=a[2], course=a[3], unit=a[4])} for a in idc]

    # ---- MMR (reference fp32 BLAS vs our pinned fp64) ----------------------
    mm = []
    for seed, n, d, k, lam in [(0, 24, 64, 8, 0.5), (1, 24, 96, 8, 0.5), (2, 10, 32, 8, 0.3),
                               (3, 5, 32, 8, 0.5), (4, 24, 768, 8, 0.5), (5, 1, 32, 8, 0.5),
                               (6, 24, 64, 24, 0.9)]:
        r = np.random.default_rng(1000 + seed)
        cb, cf = unit_bf16(r, n, d)
        base = cf[r.integers(0, n)] + 0.6 * r.standard_normal(d).astype(np.float32) / np.sqrt(d)
        qb = o.f32_to_bf16_bits(base / np.linalg.norm(base))
        qf = o.bf16_bits_to_f32(qb)
        # order candidates by similarity first, as the store would return them
        order = o.order_desc_then_index(o.exact_dots(qb, cb))
        cb, cf = cb[order], cf[order]
        ids = [f"c{i}" for i in range(n)]
        out = ref.fusion._mmr_order(q=qf, cands=cf, ids=ids, k=k, lambd=lam)
        mm.append({"n": n, "d": d, "k": k, "lambda": lam, "q_bits": bits_hex(qb),
                   "cand_bits": bits_hex(cb), "out": [int(i) for i in out]})
    g["mmr_order"] = mm

    # ---- BM25Store.search + HybridRetriever.retrieve ----------------------
    r = np.random.default_rng(77)
    n_docs, d = 60, 32
    docs = synth_docs(r, n_docs)
    docs[7] = ""            # empty text
    docs[13] = docs[12]     # exact duplicate text
    metas_c = []
    for i in range(n_docs):
        f = i // 6  # six chunks per file, two pages of three chunks; course/unit are per file
        m = {"language": "en", "course": "Math101" if f % 3 else "Phys202", "source_path": f"/data/f{f}.pdf",
             "page": 1 + (i % 6) // 3, "chunk_id": i % 6, "unit": "U1" if f % 2 else "U2"}
        if i % 5 == 0:
            m["tag_exam"] = True
        metas_c.append(m)
    ids_c = [ref.ids.stable_chunk_id(source_path=m["source_path"], page=m["page"], chunk_index=m["chunk_id"],
                                     course=m["course"], unit=m["unit"]) for m in metas_c]
    emb_bits, emb_f32 = unit_bf16(r, n_docs, d)
    emb_bits[21] = emb_bits[20]
    emb_f32[21] = emb_f32[20]   # exact duplicate embedding -> tie

    with tempfile.TemporaryDirectory() as td:
        store = ref.bm25.BM25Store(index_dir=Path(td) / "bm25")
        store.upsert_many(ids=ids_c, texts=docs, metadatas=metas_c)
        store.save()
        vs = ref.vector_chroma.ChromaVectorStore(persist_dir=Path(td) / "chroma", collection_name="golden")
        vs.upsert(ids=ids_c, documents=docs, metadatas=metas_c, embeddings=emb_f32)

        queries = ["gradient descent kernel", "memory memory bandwidth", "zzzunknownzzz", "the of and",
                   "   ", "quantum photon lattice algebra topology", "matrix"]
        wheres_b = [None, {"course": "Math101"}, {"course": "Phys202", "unit": "U2"}, {"course": "Nope"},
                    {"course": None}]
        bs = []
        for qtext in queries:
            for w in wheres_b:
                for k in (3, 8, 100):
                    res = store.search(query=qtext, where=w, top_k=k)
                    bs.append({"query": qtext, "where": w, "top_k": k,
                               "out": [[x["id"], hexf(x["score"])] for x in res]})
        g["bm25_search"] = bs

        class Emb:
            def __init__(self):
                self.vec = None

            def encode_queries(self, texts):
                return np.stack([self.vec for _ in texts])

        emb = Emb()
        rt = []
        for qi, qtext in enumerate(queries):
            for filt in ({}, {"course": "Math101"}, {"course": "Phys202", "tags": ["exam"]}):
                for hybrid in (True, False):
                    for use_mmr in (True, False):
                        rr = np.random.default_rng(500 + qi)
                        tgt = int(rr.integers(0, n_docs))
                        v = emb_f32[tgt] + 0.5 * rr.standard_normal(d).astype(np.float32) / np.sqrt(d)
                        qb = o.f32_to_bf16_bits(v / np.linalg.norm(v))
                        emb.vec = o.bf16_bits_to_f32(qb)
                        hr = ref.fusion.HybridRetriever(vector_store=vs, bm25_store=store, embedder=emb,
                                                        k_vector=8, k_bm25=8, use_mmr=use_mmr)
                        out = hr.retrieve(question=qtext, filters=filt, top_k=8, hybrid=hybrid)
                        rt.append({
                            "question": qtext, "filters": filt, "hybrid": hybrid, "use_mmr": use_mmr,
                            "q_bits": bits_hex(qb),
                            "out": [{"id": x["id"],
                                     "fused": hexf(x["scores"]["fused"]),
                                     "vector_distance": None if x["scores"]["vector_distance"] is None
                                     else hexf(x["scores"]["vector_distance"]),
                                     "bm25_score": None if x["scores"]["bm25_score"] is None
                                     else hexf(x["scores"]["bm25_score"])} for x in out]})
        g["retrieve"] = rt

        # ---- expand_with_neighbors ---------------------------------------
        ref.expand._BM25_JSONL = store.index_path
        ex = []
        for seeds in ([0, 5, 30], [2, 3, 4], [7, 8], [59, 0], [12, 13, 14, 15, 16, 17]):
            results = [{"id": ids_c[i], "document": docs[i], "metadata": metas_c[i],
                        "scores": {"fused": 0.1}} for i in seeds]
            for radius, cap in ((1, 3), (2, None), (0, 1), (1, 1)):
                out = ref.expand.expand_with_neighbors(results, radius=radius, max_per_doc=cap)
                ex.append({"seeds": seeds, "radius": radius, "max_per_doc": cap,
                           "out": [[x["id"], hexf(x["score"])] for x in out]})
        g["expand"] = ex

    g["corpus"] = {"ids": ids_c, "docs": docs, "metas": metas_c, "emb_bits": bits_hex(emb_bits),
                   "n": n_docs, "d": d}

    out_path = Path(__file__).with_name("reference_glue.json")
    out_path.write_text(json.dumps(g, ensure_ascii=False, indent=0, sort_keys=True))
    print("wrote", out_path, out_path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
