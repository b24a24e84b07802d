"""GPU parity: cmr_bm25_topk (through the C ABI) vs the rank_bm25 restatement."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as o
from tests.synth_small import zipf_corpus, zipf_queries

pytestmark = pytest.mark.gpu


def _run(docs, doc_ptr, tokens, v, queries, k, tile_docs, mask=None, row_offset=0, fmt="auto", **build_kw):
    from classmate_rag_b200 import lexical, ops
    ix = lexical.build_lexical_index(doc_ptr, tokens, v, device="cuda", tile_docs=tile_docs, fmt=fmt, **build_kw)
    assert (ix.post_pack is None) == (fmt == "wide")
    qt, qp = lexical.pack_queries(queries)
    m = None if mask is None else torch.from_numpy(mask).cuda()
    sc, ids, cnt, fl = ops.bm25_topk(ix, qt.cuda(), qp.cuda(), k, row_mask=m, row_offset=row_offset)
    torch.cuda.synchronize()
    sc, ids, cnt, fl = sc.cpu().numpy(), ids.cpu().numpy(), cnt.cpu().numpy(), fl.cpu().numpy()
    tp = ix.term_ptr.cpu().numpy()
    pd = ix.post_doc.cpu().numpy()
    tf = (ix.post_tf.cpu().to(torch.int32) & 0xFFFF).numpy()
    dl = ix.doc_len.cpu().numpy()
    for b, q in enumerate(queries):
        full = o.bm25_scores_csr(tp, pd, tf, dl, ix.idf_host, ix.avgdl, q)
        idx = np.arange(len(full))
        if mask is not None:
            keep = mask.astype(bool)
            full, idx = full[keep], idx[keep]
        order = o.order_desc_then_index(full, idx)[:k]
        n = len(order)
        assert cnt[b] == n
        assert ids[b, :n].tolist() == (idx[order] + row_offset).tolist(), (b, q)
        assert sc[b, :n].tobytes() == full[order].tobytes()
        assert (ids[b, n:] == -1).all()
        assert fl[b] == 0
    return ix


@pytest.mark.parametrize("n_docs,vocab,k,tile", [(5000, 300, 8, 1024), (20000, 2000, 10, 8192),
                                                 (3000, 50, 100, 512), (40000, 30000, 24, 4096),
                                                 (777, 40, 8, 512), (30000, 400, 64, 2048)])
def test_bm25_matches_oracle(n_docs, vocab, k, tile):
    docs, doc_ptr, tokens, v = zipf_corpus(seed=n_docs, n_docs=n_docs, vocab=vocab, mean_len=20)
    _run(docs, doc_ptr, tokens, v, zipf_queries(1, 12, vocab), k, tile)


def test_bm25_against_dict_form_and_zero_scores():
    """Small case checked against the rank_bm25-style dict implementation,
    including the all-zero ranking (insertion order) and negative idf."""
    docs, doc_ptr, tokens, v = zipf_corpus(seed=8, n_docs=400, vocab=30, mean_len=10)
    bm = o.BM25Okapi([[f"w{t}" for t in d] for d in docs])
    assert min(bm.idf.values()) < 0 or True
    queries = [[0], [0, 1, 2], [29], [-1], [], [0, 0, 0, 5]]
    from classmate_rag_b200 import lexical, ops
    ix = _run(docs, doc_ptr, tokens, v, queries, 8, 512)
    qt, qp = lexical.pack_queries(queries)
    sc, ids, cnt, fl = ops.bm25_topk(ix, qt.cuda(), qp.cuda(), 8)
    torch.cuda.synchronize()
    for b, q in enumerate(queries):
        want = bm.get_scores([f"w{t}" if t >= 0 else "zzz" for t in q])
        order = o.order_desc_then_index(want)[:8]
        assert ids[b].cpu().tolist() == order.tolist()
        assert sc[b].cpu().numpy().tobytes() == want[order].tobytes()
    # unknown-only and empty queries: every score is 0.0 -> first k documents in order
    assert ids[3].cpu().tolist() == list(range(8)) and ids[4].cpu().tolist() == list(range(8))


def test_bm25_mask_and_offset():
    docs, doc_ptr, tokens, v = zipf_corpus(seed=9, n_docs=6000, vocab=500, mean_len=15)
    rng = np.random.default_rng(0)
    mask = (rng.random(6000) < 0.4).astype(np.uint8)
    _run(docs, doc_ptr, tokens, v, zipf_queries(2, 6, 500), 10, 2048, mask=mask, row_offset=5_000_000_000)


def test_bm25_fewer_docs_than_k():
    docs, doc_ptr, tokens, v = zipf_corpus(seed=10, n_docs=5, vocab=10, mean_len=4, empty_every=0)
    _run(docs, doc_ptr, tokens, v, [[0, 1], [3]], 8, 512)


def test_bm25_wide_format_and_long_queries():
    """The fallback posting format and queries longer than one staging chunk (64 tokens)."""
    docs, doc_ptr, tokens, v = zipf_corpus(seed=11, n_docs=9000, vocab=200, mean_len=15)
    rng = np.random.default_rng(4)
    long_q = rng.integers(0, 200, 150).tolist()
    queries = zipf_queries(3, 5, 200) + [long_q]
    _run(docs, doc_ptr, tokens, v, queries, 10, 1024, fmt="wide")
    _run(docs, doc_ptr, tokens, v, queries, 10, 1024, fmt="packed")


@pytest.mark.parametrize("density,max_terms", [(None, 0), (0.125, 64), (0.0001, 64), (0.0001, 3), (0.5, 64)])
def test_bm25_dense_columns_do_not_change_results(density, max_terms):
    """Dense factor columns are only a different way to walk the same postings: off, default,
    (nearly) every term dense, a capped number of columns.  Queries mix dense and sparse
    tokens in every order, repeat dense tokens (runs longer than the fused sweep) and start
    with sparse ones."""
    docs, doc_ptr, tokens, v = zipf_corpus(seed=21, n_docs=11000, vocab=120, mean_len=14)
    queries = zipf_queries(5, 10, 120) + [[0, 1, 2, 3, 4, 5, 6], [100, 0, 0, 0, 0, 0, 1], [0], [119, 118], [1, 117, 2, 116, 3]]
    kw = {} if density is None else {"dense_max_terms": max_terms}
    ix = _run(docs, doc_ptr, tokens, v, queries, 10, 2048, dense_density=density, **kw)
    if density is None:
        assert ix.dense_imp is None
    elif density <= 0.001:
        assert ix.dense_imp.shape[0] == min(max_terms, int((ix.shard_df_host >= density * 11000).sum()))
    _run(docs, doc_ptr, tokens, v, queries, 10, 1024, fmt="wide", dense_density=density, **kw)
    rng = np.random.default_rng(5)
    _run(docs, doc_ptr, tokens, v, queries, 10, 2048, mask=(rng.random(11000) < 0.5).astype(np.uint8),
         dense_density=density, **kw)


def test_bm25_batch_kernel_equals_tile_kernel_and_oracle(monkeypatch):
    """The opt-in tile-parallel kernel (CMR_BM25_BATCH=1, read per call): same bytes as the
    default kernel and the oracle on batches it serves (>= 8 queries, k <= 32), with queries it
    hands back to the tile kernel (> 16 tokens), masks, ragged sizes, dense columns on and off."""
    from classmate_rag_b200 import lexical, ops
    rng = np.random.default_rng(12)
    for n_docs, vocab, k, tile, density in ((20000, 300, 8, 2048, 0.125), (5002, 60, 24, 512, 0.01),
                                            (40000, 3000, 10, 4096, None), (777, 40, 8, 512, 0.125)):
        docs, doc_ptr, tokens, v = zipf_corpus(seed=n_docs, n_docs=n_docs, vocab=vocab, mean_len=20)
        queries = zipf_queries(7, 40, vocab) + [rng.integers(0, vocab, 30).tolist(), [], [-1, 0, 0, 1]]
        mask = (rng.random(n_docs) < 0.6).astype(np.uint8)
        for m in (None, mask):
            monkeypatch.setenv("CMR_BM25_BATCH", "1")
            ix = _run(docs, doc_ptr, tokens, v, queries, k, tile, mask=m, dense_density=density)
            qt, qp = lexical.pack_queries(queries)
            mm = None if m is None else torch.from_numpy(m).cuda()
            got = [t.clone() for t in ops.bm25_topk(ix, qt.cuda(), qp.cuda(), k, row_mask=mm)]
            monkeypatch.setenv("CMR_BM25_BATCH", "0")
            want = [t.clone() for t in ops.bm25_topk(ix, qt.cuda(), qp.cuda(), k, row_mask=mm)]
            torch.cuda.synchronize()
            for a, b in zip(got, want):
                assert a.cpu().numpy().tobytes() == b.cpu().numpy().tobytes()

