"""GPU parity: MMR, RRF/merge/final sort and the shard merge vs the oracle and
the golden vectors produced by the live reference glue."""
import math

import numpy as np
import pytest
import torch

from oracle import np_oracle as o
from tests.helpers import bits_from_hex, unhex

pytestmark = pytest.mark.gpu


def _dev_bits(bits):
    return torch.from_numpy(np.ascontiguousarray(bits).view(np.int16)).cuda().view(torch.bfloat16)


def test_mmr_matches_reference_golden(golden):
    from classmate_rag_b200 import ops
    for c in golden["mmr_order"]:
        n, d, k = c["n"], c["d"], c["k"]
        if d % 8:
            continue
        q = bits_from_hex(c["q_bits"], (d,))
        cand = bits_from_hex(c["cand_bits"], (n, d))
        sims = o.exact_dots(q, cand)
        rows = _dev_bits(cand)[None]
        ids = torch.arange(n, dtype=torch.int64, device="cuda")[None]
        out_ids, out_sims, cnt = ops.mmr_select(rows, torch.from_numpy(sims).cuda()[None], ids,
                                                torch.tensor([n], dtype=torch.int32, device="cuda"), k, c["lambda"])
        torch.cuda.synchronize()
        m = int(cnt.item())
        assert out_ids[0, :m].cpu().tolist() == c["out"]            # the reference's own order
        assert out_ids[0, :m].cpu().tolist() == o.mmr_order_bf16(q, cand, k, c["lambda"])


def test_mmr_batch_random_vs_oracle():
    from classmate_rag_b200 import ops
    rng = np.random.default_rng(5)
    b, pool, d, k = 6, 24, 768, 8
    cand = o.f32_to_bf16_bits(rng.standard_normal((b, pool, d)).astype(np.float32) / np.sqrt(d))
    cand[2, 5] = cand[2, 1]                     # exact duplicate inside a pool
    q = o.f32_to_bf16_bits(rng.standard_normal((b, d)).astype(np.float32) / np.sqrt(d))
    counts = np.array([24, 24, 24, 3, 1, 0], dtype=np.int32)
    sims = np.stack([o.exact_dots(q[i], cand[i]) for i in range(b)])
    order = np.argsort(-sims, axis=1, kind="stable")
    cand = np.take_along_axis(cand, order[:, :, None], axis=1)
    sims = np.take_along_axis(sims, order, axis=1)
    ids = torch.arange(b * pool, dtype=torch.int64).reshape(b, pool).cuda()
    out_ids, out_sims, cnt = ops.mmr_select(_dev_bits(cand.reshape(-1, d)).reshape(b, pool, d),
                                            torch.from_numpy(sims).cuda(), ids, torch.from_numpy(counts).cuda(), k, 0.5)
    torch.cuda.synchronize()
    for i in range(b):
        n = int(counts[i])
        want = o.mmr_order_bf16(q[i], cand[i, :n], k, 0.5) if n else []
        assert int(cnt[i]) == len(want)
        assert (out_ids[i, :len(want)].cpu().numpy() - i * pool).tolist() == want
        assert out_sims[i, :len(want)].cpu().numpy().tobytes() == sims[i, want].tobytes()


def _fuse_case(vec, bm, top_k, rrf_k=60, w_vec=1.0, w_bm=1.0, hybrid=True):
    from classmate_rag_b200 import ops
    kv, kb = max(1, len(vec)), max(1, len(bm))
    v_ids = torch.full((1, kv), -1, dtype=torch.int64)
    v_s = torch.zeros((1, kv), dtype=torch.float64)
    for i, (a, s) in enumerate(vec):
        v_ids[0, i], v_s[0, i] = a, s
    b_ids = torch.full((1, kb), -1, dtype=torch.int64)
    b_s = torch.zeros((1, kb), dtype=torch.float64)
    for i, (a, s) in enumerate(bm):
        b_ids[0, i], b_s[0, i] = a, s
    vt = (v_ids.cuda(), v_s.cuda(), torch.tensor([len(vec)], dtype=torch.int32).cuda())
    bt = (b_ids.cuda(), b_s.cuda(), torch.tensor([len(bm)], dtype=torch.int32).cuda()) if hybrid else None
    ids, fused, vd, bms, cnt = ops.hybrid_fuse(vt, bt, top_k=top_k, rrf_k=rrf_k, w_vec=w_vec, w_bm=w_bm)
    torch.cuda.synchronize()
    n = int(cnt.item())
    got = []
    for i in range(n):
        got.append({"id": int(ids[0, i]), "fused": float(fused[0, i]),
                    "vector_distance": None if math.isnan(float(vd[0, i])) else float(vd[0, i]),
                    "bm25_score": None if math.isnan(float(bms[0, i])) else float(bms[0, i])})
    want = o.hybrid_merge([(a, 1.0 - s) for a, s in vec], list(bm), top_k, rrf_k, w_vec, w_bm, hybrid)
    assert got == want
    return got


def test_hybrid_fuse_vs_oracle_random():
    rng = np.random.default_rng(1)
    for trial in range(40):
        nv, nb = int(rng.integers(0, 25)), int(rng.integers(0, 25))
        pool = rng.permutation(60)
        vec = [(int(pool[i]), float(rng.random())) for i in range(nv)]
        bm_ids = rng.permutation(60)[:nb]
        bm = [(int(i), float(rng.random() * 9)) for i in bm_ids]
        _fuse_case(vec, bm, top_k=int(rng.integers(1, 30)), rrf_k=int(rng.choice([1, 10, 60])),
                   w_vec=float(rng.choice([1.0, 0.3, 1.7])), w_bm=float(rng.choice([1.0, 2.0])))
    _fuse_case([(5, 0.75)], [(9, 3.0)], 5)          # equal fused: BM25-only first
    _fuse_case([(5, 1.0)], [(9, 3.0)], 5)           # distance exactly 0.0: insertion order
    _fuse_case([(1, 0.9), (2, 0.8)], [], 8, hybrid=False)
    _fuse_case([], [(3, 1.0), (4, 0.0)], 8)


def test_hybrid_fuse_golden_retrieve(golden):
    """Feed the reference's own dense/BM25 lists (recovered from the golden
    output) through the kernel: fused scores and order must be bit-exact."""
    c = golden["corpus"]
    emb = bits_from_hex(c["emb_bits"], (c["n"], c["d"]))
    row_of = {cid: i for i, cid in enumerate(c["ids"])}
    entries = [(cid, o.tokenize(t, m.get("language")), m) for cid, t, m in zip(c["ids"], c["docs"], c["metas"])]
    for case in golden["retrieve"]:
        q = bits_from_hex(case["q_bits"], (c["d"],))
        filt = case["filters"] or {}
        where = o.build_where_filter(filt) if filt else None
        mask = np.array([o.chroma_where_matches(m, where) for m in c["metas"]], dtype=np.uint8)
        hybrid, use_mmr = case["hybrid"], case["use_mmr"]
        pool = 24 if use_mmr else 8
        ids, sc = o.dense_topk(q, emb, pool, mask=mask)
        if use_mmr and len(ids):
            order = o.mmr_order_bf16(q, emb[ids], 8, 0.5)
            ids, sc = ids[order], sc[order]
        else:
            ids, sc = ids[:8], sc[:8]
        vec = list(zip(ids.tolist(), sc.tolist()))
        bm = [(row_of[i], s) for i, s in o.bm25_store_search(entries, case["question"], filt or None, 8)] if hybrid else []
        got = _fuse_case(vec, bm, 8, hybrid=hybrid)
        assert [c["ids"][g["id"]] for g in got] == [w["id"] for w in case["out"]]
        for g, w in zip(got, case["out"]):
            assert g["fused"] == unhex(w["fused"]) and g["bm25_score"] == unhex(w["bm25_score"])


def test_topk_merge_and_gather():
    from classmate_rag_b200 import ops
    rng = np.random.default_rng(2)
    g, b, k = 4, 5, 10
    scores = np.round(rng.random((g, b, k)), 2)      # coarse -> ties across shards
    scores = -np.sort(-scores, axis=2)
    ids = rng.permutation(g * b * k).reshape(g, b, k).astype(np.int64)
    counts = rng.integers(0, k + 1, (g, b)).astype(np.int32)
    counts[0, 0] = 0
    s, i, c = ops.topk_merge(torch.from_numpy(scores).cuda(), torch.from_numpy(ids).cuda(), torch.from_numpy(counts).cuda())
    torch.cuda.synchronize()
    for q in range(b):
        allv = [(scores[p, q, r], ids[p, q, r]) for p in range(g) for r in range(counts[p, q])]
        allv.sort(key=lambda x: (-x[0], x[1]))
        want = allv[:k]
        n = int(c[q])
        assert n == len(want)
        assert i[q, :n].cpu().tolist() == [int(x[1]) for x in want]
        assert s[q, :n].cpu().tolist() == [float(x[0]) for x in want]
    emb = torch.randn(100, 64, device="cuda").to(torch.bfloat16)
    pick = torch.tensor([[1005, 1099, -1, 7, 1100]], dtype=torch.int64, device="cuda")
    rows = ops.gather_rows(emb, pick, row_offset=1000)
    torch.cuda.synchronize()
    assert torch.equal(rows[0, 0], emb[5]) and torch.equal(rows[0, 1], emb[99])
    assert rows[0, 2].abs().sum() == 0 and rows[0, 3].abs().sum() == 0 and rows[0, 4].abs().sum() == 0
