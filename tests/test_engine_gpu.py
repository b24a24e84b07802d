"""GPU parity of the whole hot path (dense -> MMR -> BM25 -> RRF -> final
order) through the engine, eager and CUDA-graph, vs the oracle pipeline."""
import math

import numpy as np
import pytest
import torch

from oracle import np_oracle as o

pytestmark = pytest.mark.gpu


def _bits(t: torch.Tensor) -> np.ndarray:
    return t.contiguous().view(torch.int16).cpu().numpy().view(np.uint16)


def _oracle_hybrid(emb_bits, q_bits, lex_arrays, terms, p):
    tp, pd, tf, dl, idf, avgdl = lex_arrays
    hybrid = p.hybrid and terms is not None
    k_vec = p.k_vector if hybrid else max(p.top_k, p.k_vector)
    pool = max(k_vec, p.mmr_max_pool) if p.use_mmr else k_vec
    ids, sc = o.dense_topk(q_bits, emb_bits, pool)
    if p.use_mmr and len(ids):
        order = o.mmr_order_bf16(q_bits, emb_bits[ids], k_vec, p.mmr_lambda)
        ids, sc = ids[order], sc[order]
    else:
        ids, sc = ids[:k_vec], sc[:k_vec]
    vec = [(int(i), 1.0 - float(s)) for i, s in zip(ids, sc)]
    bm = []
    if hybrid:
        full = o.bm25_scores_csr(tp, pd, tf, dl, idf, avgdl, terms)
        bi, bs = o.bm25_topk(full, p.k_bm25)
        bm = [(int(i), float(s)) for i, s in zip(bi, bs)]
    return o.hybrid_merge(vec, bm, p.top_k, p.rrf_k, p.weight_vector, p.weight_bm25, hybrid)


def _check(out, want_lists):
    ids, fused, vd, bm, cnt = out
    for b, want in enumerate(want_lists):
        n = int(cnt[b])
        assert n == len(want)
        for i, w in enumerate(want):
            assert int(ids[b, i]) == w["id"]
            assert float(fused[b, i]) == w["fused"]
            g_vd = None if math.isnan(float(vd[b, i])) else float(vd[b, i])
            g_bm = None if math.isnan(float(bm[b, i])) else float(bm[b, i])
            assert g_vd == w["vector_distance"] and g_bm == w["bm25_score"]


@pytest.fixture(scope="module")
def small_world():
    from classmate_rag_b200 import lexical, synth
    from classmate_rag_b200.engine import HybridEngine
    n, d, vocab = 30000, 256, 3000
    emb = synth.dense_corpus(n, d, "cuda")
    doc_ptr, tokens = synth.lexical_corpus(n, vocab, 24, "cuda")
    lex = lexical.build_lexical_index(doc_ptr, tokens, vocab, tile_docs=2048)
    eng = HybridEngine(emb, lex)
    q, planted = synth.dense_queries(n, d, 6, "cuda")
    terms = synth.lexical_queries(6, vocab)
    terms[2][1] = -1
    terms[3][5] = terms[3][0]
    lex_arrays = (lex.term_ptr.cpu().numpy(), lex.post_doc.cpu().numpy(),
                  (lex.post_tf.cpu().to(torch.int32) & 0xFFFF).numpy(), lex.doc_len.cpu().numpy(),
                  lex.idf_host, lex.avgdl)
    return eng, emb, lex, q, planted, terms, lex_arrays


@pytest.mark.parametrize("hybrid,use_mmr,top_k", [(True, True, 8), (True, False, 10), (False, True, 12), (False, False, 5)])
def test_engine_matches_oracle_pipeline(small_world, hybrid, use_mmr, top_k):
    from classmate_rag_b200 import lexical, ops
    from classmate_rag_b200.engine import SearchParams
    eng, emb, lex, q, planted, terms, lex_arrays = small_world
    p = SearchParams(top_k=top_k, hybrid=hybrid, use_mmr=use_mmr)
    q_bf16 = ops.f32_to_bf16(q)
    qt, qp = lexical.pack_queries(terms)
    out = eng.search(q_bf16, qt.cuda(), qp.cuda(), p)
    torch.cuda.synchronize()
    out = [t.cpu().numpy() for t in out]
    emb_bits, q_bits = _bits(emb), _bits(q_bf16)
    want = [_oracle_hybrid(emb_bits, q_bits[b], lex_arrays, terms[b] if hybrid else None, p) for b in range(len(terms))]
    _check(out, want)
    assert int(eng.last_dense_flags.sum()) == 0
    if not use_mmr:  # the planted row is the exact top-1 of the dense list
        for b in range(len(terms)):
            assert int(planted[b]) in [w["id"] for w in want[b]] or hybrid


def test_side_stream_schedules_give_identical_results(small_world):
    """Batches on the tcgen05 path (> 8 queries) run BM25 on a side stream (engine overlap): serial,
    BM25-first and dense-first schedules, eager and from a CUDA graph, must return the same bytes,
    and those must be the oracle pipeline's."""
    from classmate_rag_b200 import lexical, ops, synth
    from classmate_rag_b200.engine import GraphedSearch, HybridEngine, SearchParams
    eng0, emb, lex, _, _, _, lex_arrays = small_world
    nq = 20
    q, _ = synth.dense_queries(emb.shape[0], emb.shape[1], nq, "cuda")
    terms = synth.lexical_queries(nq, 3000)
    q_bf16 = ops.f32_to_bf16(q)
    qt, qp = [t.cuda() for t in lexical.pack_queries(terms)]
    p = SearchParams(top_k=10)
    ref = None
    for overlap, first in ((False, True), (True, True), (True, False)):
        eng = HybridEngine(emb, lex, overlap=overlap)
        eng.bm25_first = first
        for rep in range(3):
            out = [t.clone() for t in eng.search(q_bf16, qt, qp, p)]
            torch.cuda.synchronize()
            got = [t.cpu().numpy().tobytes() for t in out]
            ref = ref or got
            assert got == ref, (overlap, first, rep)
        gs = GraphedSearch(eng, p, nq, max_terms=16)
        for rep in range(3):
            g = gs(q.cpu().numpy(), terms)
            assert [a.tobytes() for a in g] == ref, (overlap, first, "graph", rep)
    emb_bits, q_bits = _bits(emb), _bits(q_bf16)
    want = [_oracle_hybrid(emb_bits, q_bits[b], lex_arrays, terms[b], p) for b in range(4)]
    _check([t.cpu().numpy()[:4] for t in eng.search(q_bf16, qt, qp, p)], want)


def test_graphed_search_equals_eager_and_is_deterministic(small_world):
    from classmate_rag_b200 import lexical, ops
    from classmate_rag_b200.engine import GraphedSearch, SearchParams
    eng, emb, lex, q, planted, terms, lex_arrays = small_world
    p = SearchParams(top_k=10)
    gs = GraphedSearch(eng, p, n_queries=1, max_terms=16)
    assert gs.graph is not None
    for b in range(3):
        got = gs(q[b:b + 1].cpu().numpy(), [terms[b]])
        got = [g.copy() for g in got]
        again = gs(q[b:b + 1].cpu().numpy(), [terms[b]])
        for a, c in zip(got, again):
            assert a.tobytes() == c.tobytes()           # run twice, bit-compare
        qt, qp = lexical.pack_queries([terms[b]])
        eager = eng.search(ops.f32_to_bf16(q[b:b + 1]), qt.cuda(), qp.cuda(), p)
        torch.cuda.synchronize()
        for a, c in zip(got, eager):
            assert a.tobytes() == c.cpu().numpy().tobytes()
    assert gs.h2d_bytes > 0 and gs.d2h_bytes > 0


@pytest.mark.parametrize("n_shards,hybrid,use_mmr", [(2, True, True), (3, True, False), (8, False, True), (5, True, True)])
def test_single_exchange_shard_merge_equals_unsharded(small_world, n_shards, hybrid, use_mmr):
    """The sharded step's message protocol (cmr_shard_pack -> all-gather -> cmr_shard_merge),
    with the ranks emulated one after the other on one GPU: the concatenation of the ranks'
    messages is exactly what the all-gather delivers.  Result must equal the unsharded
    search bit for bit, for any number of shards."""
    from classmate_rag_b200 import lexical, ops, sharding, synth
    from classmate_rag_b200.engine import HybridEngine, SearchParams
    eng, emb, lex, q, planted, terms, lex_arrays = small_world
    n, d = emb.shape
    vocab = lex.n_terms
    p = SearchParams(top_k=9, hybrid=hybrid, use_mmr=use_mmr)
    q_bf16 = ops.f32_to_bf16(q)
    qt, qp = lexical.pack_queries(terms)
    qt, qp = qt.cuda(), qp.cuda()
    want = [t.clone() for t in eng.search(q_bf16, qt, qp, p)]

    doc_ptr, tokens = synth.lexical_corpus(n, vocab, 24, "cuda")
    stats = lexical.corpus_stats(doc_ptr, tokens, vocab)
    k_vec = p.k_vector if hybrid else max(p.top_k, p.k_vector)
    pool = min(max(k_vec, p.mmr_max_pool), 64) if use_mmr else k_vec
    msgs = []
    for r in range(n_shards):
        lo, hi = sharding.shard_range(n, r, n_shards)
        t_lo, t_hi = int(doc_ptr[lo]), int(doc_ptr[hi])
        sh_lex = lexical.build_lexical_index(doc_ptr[lo:hi + 1] - doc_ptr[lo], tokens[t_lo:t_hi], vocab,
                                             tile_docs=2048, stats=stats)
        sh = HybridEngine(emb[lo:hi].contiguous(), sh_lex, row_offset=lo)
        dense = sh.dense_pool(q_bf16, pool)
        bm = None
        if hybrid:
            b_sc, b_ids, b_cnt, _ = sh.lexical_topk(qt, qp, p.k_bm25)
            bm = (b_sc, b_ids, b_cnt)
        msgs.append(ops.shard_pack(dense, bm, sh.emb if use_mmr else None, row_offset=lo).clone())
    gathered = torch.stack(msgs)
    assert gathered.shape[2] == ops.shard_msg_bytes(pool, p.k_bm25 if hybrid else 0, d if use_mmr else 0)
    d_s, d_i, d_c, d_f, rows, b_s, b_i, b_c = ops.shard_merge(gathered, pool, p.k_bm25 if hybrid else 0,
                                                               d if use_mmr else 0)
    if use_mmr:
        v_ids, v_sims, v_cnt = ops.mmr_select(rows, d_s, d_i, d_c, min(k_vec, pool), p.mmr_lambda)
    else:
        v_ids, v_sims, v_cnt = d_i, d_s, d_c
    got = ops.hybrid_fuse((v_ids, v_sims, v_cnt), (b_i, b_s, b_c) if hybrid else None, top_k=p.top_k, rrf_k=p.rrf_k,
                          w_vec=p.weight_vector if hybrid else 1.0, w_bm=p.weight_bm25)
    torch.cuda.synchronize()
    assert int(d_f.sum()) == 0
    for a, b in zip(got, want):
        assert a.cpu().numpy().tobytes() == b.cpu().numpy().tobytes()


@pytest.mark.parametrize("hybrid,use_mmr", [(True, True), (False, True), (True, False)])
def test_search_in_two_halves_equals_search(small_world, hybrid, use_mmr):
    """engine.search_heavy (the scans) + engine.search_tail (merge, MMR, fusion) -- the two graphs of the
    pipelined form -- give the bytes of engine.search, on either buffer slot."""
    from classmate_rag_b200 import lexical, ops
    from classmate_rag_b200.engine import SearchParams
    eng, emb, lex, q, planted, terms, lex_arrays = small_world
    p = SearchParams(top_k=10, hybrid=hybrid, use_mmr=use_mmr)
    qb = ops.f32_to_bf16(q[:6])
    qt, qp = lexical.pack_queries(terms[:6])
    qt, qp = qt.cuda(), qp.cuda()
    want = [t.clone() for t in eng.search(qb, qt, qp, p)]
    for slot in (0, 1):
        state = eng.search_heavy(qb, qt, qp, p, slot=slot)
        got = eng.search_tail(state, p)
        torch.cuda.synchronize()
        for a, b in zip(got, want):
            assert a.cpu().numpy().tobytes() == b.cpu().numpy().tobytes()


def test_pipelined_search_resident_rotation(small_world):
    """launch_resident alternates the two slots; each launch's outputs stay valid for one more launch."""
    from classmate_rag_b200 import lexical
    from classmate_rag_b200.engine import GraphedSearch, PipelinedSearch, SearchParams
    eng, emb, lex, q, planted, terms, lex_arrays = small_world
    p = SearchParams(top_k=10)
    one = GraphedSearch(eng, p, n_queries=2, max_terms=16)
    qh = q.cpu().numpy()
    ps = PipelinedSearch(eng, p, 2, max_terms=16)
    assert ps.independent and ps.slots[0].graph_h is not None
    prev = None
    for i in (0, 2, 4, 0, 2):
        want = [x.copy() for x in one(qh[i:i + 2], terms[i:i + 2])]
        qt, qp = lexical.pack_queries(terms[i:i + 2])
        out = ps.launch_resident(q[i:i + 2].contiguous(), qt.cuda(), qp.cuda())
        if prev is not None:      # the launch before this one is still intact
            ps.wait()
            torch.cuda.synchronize()
            for a, b in zip(prev[0], prev[1]):
                assert a.cpu().numpy().tobytes() == b.tobytes()
        prev = (out, want)
    ps.wait()
    torch.cuda.synchronize()
    for a, b in zip(prev[0], prev[1]):
        assert a.cpu().numpy().tobytes() == b.tobytes()


def test_pipelined_search_returns_each_batch_in_order(small_world):
    """PipelinedSearch (two graph objects, scans in order on one stream, tails on their own streams) hands
    back, one submit late, exactly what the single-slot GraphedSearch returns for the same batch."""
    from classmate_rag_b200.engine import GraphedSearch, PipelinedSearch, SearchParams
    eng, emb, lex, q, planted, terms, lex_arrays = small_world
    p = SearchParams(top_k=10)
    qh = q.cpu().numpy()
    one = GraphedSearch(eng, p, n_queries=2, max_terms=16)
    want = [[x.copy() for x in one(qh[i:i + 2], terms[i:i + 2])] for i in (0, 2, 4, 0)]
    ps = PipelinedSearch(eng, p, 2, max_terms=16)
    got = []
    for i in (0, 2, 4, 0):
        r = ps.submit(qh[i:i + 2], terms[i:i + 2])
        if r is not None:
            got.append(r)
    got.append(ps.drain())
    assert ps.drain() is None and len(got) == 4
    for g, w in zip(got, want):
        for a, b in zip(g, w):
            assert a.tobytes() == b.tobytes()


@pytest.mark.timeout(300)
def test_multi_gpu_sharded_search_peer_memory_and_nccl():
    """Needs >= 2 GPUs on the box (skipped otherwise): tools/shard_check.py under torchrun --
    peer-memory exchange, NCCL exchange and the unsharded engine agree bit for bit."""
    import subprocess
    import sys
    from pathlib import Path
    n_gpu = torch.cuda.device_count()
    if n_gpu < 2:
        pytest.skip("one GPU only")
    root = Path(__file__).resolve().parents[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(n_gpu, 4)}",
           "--master-addr", "127.0.0.1", "--master-port", "29577", str(root / "tools" / "shard_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "shard_check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
