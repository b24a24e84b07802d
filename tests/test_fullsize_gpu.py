"""GPU: the BASELINE.json configurations at FULL size, checked through size-independent
properties (the oracle only on a few queries, where it finishes in seconds):

  C2  1M x 768 exact dense top-10, batch 1 and batch 1024
  C3  10M x 1024 hybrid row-sharded over 2/4/8 (here: 1M x 1024, ranks emulated in sequence)
  C4  BM25 over ~50M postings, V = 30 000, batch 4096, top-100
  C5  near-duplicate filter (threshold 0.95) over 2M x 768
"""
import numpy as np
import pytest
import torch

from oracle import c_oracle, np_oracle as o

pytestmark = pytest.mark.gpu


def _bits(t):
    return t.contiguous().view(torch.int16).cpu().numpy().view(np.uint16)


def _sorted_desc_then_id(scores, ids, counts):
    s, i = scores.cpu().numpy(), ids.cpu().numpy()
    for b in range(s.shape[0]):
        n = int(counts[b])
        assert (np.diff(s[b, :n]) <= 0).all()
        same = np.diff(s[b, :n]) == 0
        assert (np.diff(i[b, :n])[same] > 0).all()          # ties: ascending id
        assert len(set(i[b, :n].tolist())) == n


def test_c2_dense_1m_768_batch1_and_batch1024():
    from classmate_rag_b200 import ops, synth
    n, d, k = 1_000_000, 768, 10
    emb = synth.dense_corpus(n, d, "cuda")
    q, planted = synth.dense_queries(n, d, 1024, "cuda")
    qb = ops.f32_to_bf16(q)
    s_m, i_m, c_m, f_m = [t.clone() for t in ops.dense_topk(emb, qb, k, algo="mma")]
    torch.cuda.synchronize()
    assert int(f_m.sum()) == 0 and int(c_m.min()) == k
    _sorted_desc_then_id(s_m, i_m, c_m)
    assert torch.equal(i_m[:, 0].cpu(), planted)              # the planted row is the exact top-1
    # batch 1 (scan path) == row b of the batch-1024 GEMM path, bit for bit
    for b in (0, 1, 511, 1023):
        s1, i1, c1, f1 = ops.dense_topk(emb, qb[b:b + 1], k, algo="scan")
        torch.cuda.synchronize()
        assert torch.equal(i1[0], i_m[b]) and s1[0].cpu().numpy().tobytes() == s_m[b].cpu().numpy().tobytes()
    # 32 queries through the scan path as well
    s_s, i_s, c_s, f_s = ops.dense_topk(emb, qb[:32], k, algo="scan")
    torch.cuda.synchronize()
    assert torch.equal(i_s, i_m[:32]) and s_s.cpu().numpy().tobytes() == s_m[:32].cpu().numpy().tobytes()
    # oracle on two queries
    emb_bits, q_bits = _bits(emb), _bits(qb[:2])
    for b in range(2):
        want_ids, want_sc = o.dense_topk(q_bits[b], emb_bits, k)
        assert i_m[b].cpu().numpy().tolist() == want_ids.tolist()
        assert s_m[b].cpu().numpy().tobytes() == want_sc.tobytes()
    # linearity of the exact score: score(q, c) is the float64 dot -> within 1e-3 relative of fp32 torch
    ref = (emb[i_m[0]].float() @ qb[0].float()).double().cpu().numpy()
    assert np.allclose(s_m[0].cpu().numpy(), ref, rtol=1e-3, atol=0)


def test_c4_bm25_50m_postings_batch4096_top100():
    from classmate_rag_b200 import lexical, ops, sharding, synth
    n, vocab, k, b = 1_000_000, 30000, 100, 4096
    doc_ptr, tokens = synth.lexical_corpus(n, vocab, 64, "cuda")
    lex = lexical.build_lexical_index(doc_ptr, tokens, vocab)
    assert 45_000_000 < lex.n_postings < 60_000_000
    terms = synth.lexical_queries(b, vocab)
    qt, qp = lexical.pack_queries(terms)
    qt, qp = qt.cuda(), qp.cuda()
    sc, ids, cnt, fl = [t.clone() for t in ops.bm25_topk(lex, qt, qp, k)]
    torch.cuda.synchronize()
    assert int(cnt.min()) == k and int(fl.sum()) == 0
    _sorted_desc_then_id(sc[:256], ids[:256], cnt[:256])
    # idempotence / batch independence: a query scores the same alone as inside the batch
    for j in (0, 7, 4095):
        q1, p1 = lexical.pack_queries([terms[j]])
        s1, i1, c1, _ = ops.bm25_topk(lex, q1.cuda(), p1.cuda(), k)
        torch.cuda.synchronize()
        assert torch.equal(i1[0], ids[j]) and s1[0].cpu().numpy().tobytes() == sc[j].cpu().numpy().tobytes()
    # dense columns off == on
    plain = lexical.build_lexical_index(doc_ptr, tokens, vocab, dense_density=None)
    s2, i2, c2, _ = ops.bm25_topk(plain, qt[: int(qp[64])], qp[:65].contiguous(), k)
    torch.cuda.synchronize()
    assert torch.equal(i2, ids[:64]) and s2.cpu().numpy().tobytes() == sc[:64].cpu().numpy().tobytes()
    # shard invariance: two document shards with corpus-wide statistics, merged
    stats = lexical.corpus_stats(doc_ptr, tokens, vocab)
    parts = []
    for r in range(2):
        lo, hi = sharding.shard_range(n, r, 2)
        t_lo, t_hi = int(doc_ptr[lo]), int(doc_ptr[hi])
        sh = lexical.build_lexical_index(doc_ptr[lo:hi + 1] - doc_ptr[lo], tokens[t_lo:t_hi], vocab, stats=stats)
        parts.append([t.clone() for t in ops.bm25_topk(sh, qt[: int(qp[64])], qp[:65].contiguous(), k, row_offset=lo)])
    ms, mi, mc = ops.topk_merge(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]),
                                torch.stack([p[2] for p in parts]))
    torch.cuda.synchronize()
    assert torch.equal(mi, ids[:64]) and ms.cpu().numpy().tobytes() == sc[:64].cpu().numpy().tobytes()
    # oracle (C restatement of rank_bm25) on three queries
    tp, pd = lex.term_ptr.cpu().numpy(), lex.post_doc.cpu().numpy()
    tf = (lex.post_tf.cpu().to(torch.int32) & 0xFFFF).numpy()
    dl = lex.doc_len.cpu().numpy()
    for j in (0, 1, 2):
        full = c_oracle.bm25_scores(tp, pd, tf, dl, lex.idf_host, lex.avgdl, terms[j])
        wi, ws = o.bm25_topk(full, k)
        assert ids[j].cpu().numpy().tolist() == wi.tolist() and sc[j].cpu().numpy().tobytes() == ws.tobytes()


def test_c3_hybrid_dim1024_sharded_2_4_8():
    from classmate_rag_b200 import lexical, ops, sharding, synth
    from classmate_rag_b200.engine import HybridEngine, SearchParams
    n, d, vocab, nq = 1_000_000, 1024, 30000, 16
    emb = synth.dense_corpus(n, d, "cuda")
    doc_ptr, tokens = synth.lexical_corpus(n, vocab, 64, "cuda")
    stats = lexical.corpus_stats(doc_ptr, tokens, vocab)
    eng = HybridEngine(emb, lexical.build_lexical_index(doc_ptr, tokens, vocab, stats=stats))
    q, planted = synth.dense_queries(n, d, nq, "cuda")
    terms = synth.lexical_queries(nq, vocab)
    qb = ops.f32_to_bf16(q)
    qt, qp = lexical.pack_queries(terms)
    qt, qp = qt.cuda(), qp.cuda()
    p = SearchParams(top_k=10)
    want = [t.clone() for t in eng.search(qb, qt, qp, p)]
    torch.cuda.synchronize()
    assert int(eng.last_dense_flags.sum()) == 0
    del eng
    pool = p.pool
    for g in (2, 4, 8):
        msgs = []
        for r in range(g):
            lo, hi = sharding.shard_range(n, r, g)
            t_lo, t_hi = int(doc_ptr[lo]), int(doc_ptr[hi])
            sh_lex = lexical.build_lexical_index(doc_ptr[lo:hi + 1] - doc_ptr[lo], tokens[t_lo:t_hi], vocab, stats=stats)
            sh = HybridEngine(emb[lo:hi], sh_lex, row_offset=lo)
            dense = sh.dense_pool(qb, pool)
            b_sc, b_ids, b_cnt, _ = sh.lexical_topk(qt, qp, p.k_bm25)
            msgs.append(ops.shard_pack(dense, (b_sc, b_ids, b_cnt), sh.emb, row_offset=lo).clone())
            del sh, sh_lex
        d_s, d_i, d_c, d_f, rows, b_s, b_i, b_c = ops.shard_merge(torch.stack(msgs), pool, p.k_bm25, d)
        v_ids, v_sims, v_cnt = ops.mmr_select(rows, d_s, d_i, d_c, p.k_vector, p.mmr_lambda)
        got = ops.hybrid_fuse((v_ids, v_sims, v_cnt), (b_i, b_s, b_c), top_k=p.top_k, rrf_k=p.rrf_k)
        torch.cuda.synchronize()
        for a, b_ in zip(got, want):
            assert a.cpu().numpy().tobytes() == b_.cpu().numpy().tobytes(), g


def test_c5_neardup_2m_768_properties():
    """2M x 768 with planted structure: 5 % noisy copies (cos ~ 0.999) and 1 % exact copies of
    EARLIER rows; everything else is random (cos << 0.95).  Under the keep-first rule a row is
    dropped iff it is a copy (of a kept row or, through a chain, of a dropped copy whose own
    source is within 0.95 -- with noise this small the whole chain collapses onto its root)."""
    from classmate_rag_b200 import neardup
    n, d = 2_000_000, 768
    g = torch.Generator(device="cuda").manual_seed(5)
    emb = torch.empty((n, d), dtype=torch.bfloat16, device="cuda")
    step = 1 << 18
    for lo in range(0, n, step):
        x = torch.randn((min(step, n - lo), d), generator=g, device="cuda")
        emb[lo:lo + x.shape[0]] = torch.nn.functional.normalize(x, dim=1).to(torch.bfloat16)
    u = torch.rand(n, generator=g, device="cuda")
    src = (torch.rand(n, generator=g, device="cuda") * torch.arange(n, device="cuda")).long()   # src[i] < i
    is_near = (u < 0.05) & (torch.arange(n, device="cuda") > 0)
    is_exact = (u >= 0.05) & (u < 0.06) & (torch.arange(n, device="cuda") > 0)
    # resolve chains to ORIGINAL rows so that the planted structure is a forest of depth 1
    root = torch.arange(n, device="cuda")
    copy = is_near | is_exact
    root[copy] = src[copy]
    for _ in range(40):
        nxt = root[root]
        if torch.equal(nxt, root):
            break
        root = nxt
    idx = torch.nonzero(is_exact).flatten()
    emb[idx] = emb[root[idx]]
    idx = torch.nonzero(is_near).flatten()
    for lo in range(0, idx.numel(), step):
        ii = idx[lo:lo + step]
        noise = torch.randn((ii.numel(), d), generator=g, device="cuda") * (0.03 / d ** 0.5)
        emb[ii] = torch.nn.functional.normalize(emb[root[ii]].float() + noise, dim=1).to(torch.bfloat16)
    keep = neardup.neardup_keep_mask(emb, 0.95)
    torch.cuda.synchronize()
    assert torch.equal(keep.bool(), ~copy)
    # idempotence: the kept rows contain no near-duplicates any more
    kept = emb[keep.bool()][:300_000].contiguous()
    assert int(neardup.neardup_keep_mask(kept, 0.95).sum()) == kept.shape[0]
