"""GPU parity: cmr_bm25_topk (through the C ABI) vs the rank_bm25 restatement."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as o
from tests.synth_small import zipf_corpus, zipf_queries

pytestmark = pytest.mark.gpu


def _run(docs, doc_ptr, tokens, v, queries, k, tile_docs, mask=None, row_offset=0, fmt="auto", algo="auto",
         allow_flags=False, ix=None, **build_kw):
    from classmate_rag_b200 import lexical, ops
    if ix is None:
        ix = lexical.build_lexical_index(doc_ptr, tokens, v, device="cuda", tile_docs=tile_docs, fmt=fmt, **build_kw)
        assert (ix.post_pack is None) == (fmt == "wide")
    qt, qp = lexical.pack_queries(queries)
    m = None if mask is None else torch.from_numpy(mask).cuda()
    sc, ids, cnt, fl = ops.bm25_topk(ix, qt.cuda(), qp.cuda(), k, row_mask=m, row_offset=row_offset, algo=algo)
    torch.cuda.synchronize()
    sc, ids, cnt, fl = sc.cpu().numpy(), ids.cpu().numpy(), cnt.cpu().numpy(), fl.cpu().numpy()
    ix.last_flags = fl
    tp = ix.term_ptr.cpu().numpy()
    pd = ix.post_doc.cpu().numpy()
    tf = (ix.post_tf.cpu().to(torch.int32) & 0xFFFF).numpy()
    dl = ix.doc_len.cpu().numpy()
    for b, q in enumerate(queries):
        if allow_flags and fl[b] != 0:
            continue      # head_nofallback: an uncertified query's output is unspecified
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


# ---- the batched path: head-term matrix on the tensor cores + bucketed sparse postings + exact rescoring ----

@pytest.mark.parametrize("n_docs,vocab,k,tile,nq", [(20000, 300, 8, 2048, 12), (5002, 60, 24, 512, 40),
                                                    (40000, 3000, 10, 2048, 70), (777, 40, 8, 512, 9),
                                                    (130, 25, 10, 512, 33), (3000, 50, 100, 1024, 16),
                                                    (66000, 30000, 10, 2048, 32), (25000, 500, 10, 2048, 130),
                                                    (9000, 200, 100, 1024, 100)])
def test_bm25_head_path_equals_oracle_and_exact_kernel(n_docs, vocab, k, tile, nq):
    """CMR_BM25_HEAD (with its exact re-run of uncertified queries) returns the oracle's bytes: ragged sizes, more
    than one block of 32 queries (groups of up to 4 blocks share one pass over the head matrix; 130 queries =
    a full group + a single block), tiny indexes without an admission bound, k = 100 (KP = 128), queries it
    must hand back (empty, unknown-only, > 16 tokens), repeated tokens."""
    from classmate_rag_b200 import lexical, ops
    rng = np.random.default_rng(n_docs)
    docs, doc_ptr, tokens, v = zipf_corpus(seed=n_docs, n_docs=n_docs, vocab=vocab, mean_len=20)
    queries = zipf_queries(7, nq - 3, vocab) + [rng.integers(0, vocab, 30).tolist(), [-1, -1], [0, 0, 1, -1, 0]]
    ix = _run(docs, doc_ptr, tokens, v, queries, k, tile, algo="head")
    assert ix.head_mat is not None and ix.head_mat.shape == (n_docs, 64)
    qt, qp = lexical.pack_queries(queries)
    got = [t.clone() for t in ops.bm25_topk(ix, qt.cuda(), qp.cuda(), k, algo="head")]
    want = [t.clone() for t in ops.bm25_topk(ix, qt.cuda(), qp.cuda(), k, algo="exact")]
    torch.cuda.synchronize()
    for a, b in zip(got, want):
        assert a.cpu().numpy().tobytes() == b.cpu().numpy().tobytes()


@pytest.mark.parametrize("rows", [0, 3, 40])
def test_bm25_skip_table_budget_does_not_change_results(rows):
    """The skip table holds rows for the most frequent terms only (a memory budget); every other term's tile
    slices are bisected on post_doc.  No row at all, three rows, forty rows: same bytes from both kernels."""
    n_docs, vocab = 30000, 900
    docs, doc_ptr, tokens, v = zipf_corpus(seed=202, n_docs=n_docs, vocab=vocab, mean_len=18)
    n_tiles = (n_docs + 2047) // 2048
    budget = max(1, rows) * (n_tiles + 1) * 4
    queries = zipf_queries(13, 20, vocab) + [[vocab - 1, vocab - 2, 0], [5, vocab - 7, vocab - 7], list(range(200, 230))]
    for algo in ("exact", "head"):
        ix = _run(docs, doc_ptr, tokens, v, queries, 10, 2048, algo=algo, skip_budget_bytes=budget)
        assert ix.tile_skip.shape[0] == max(1, rows) and int((ix.skip_row >= 0).sum()) == max(1, rows)
    _run(docs, doc_ptr, tokens, v, queries[:3], 10, 2048, algo="exact", skip_budget_bytes=budget)
    rng = np.random.default_rng(3)
    _run(docs, doc_ptr, tokens, v, queries, 10, 2048, mask=(rng.random(n_docs) < 0.5).astype(np.uint8),
         skip_budget_bytes=budget)


def test_bm25_head_path_certifies_normal_queries():
    """Without the re-run (head_nofallback) the flags say which queries were certified; every certified query is
    bit-exact, and on a Zipf corpus with 6-token queries (nearly) all of them are."""
    docs, doc_ptr, tokens, v = zipf_corpus(seed=77, n_docs=50000, vocab=5000, mean_len=30, empty_every=0)
    queries = [q for q in zipf_queries(11, 64, 5000) if q]
    ix = _run(docs, doc_ptr, tokens, v, queries, 10, 2048, algo="head_nofallback", allow_flags=True)
    flagged = int((ix.last_flags != 0).sum())
    assert flagged <= len(queries) // 8, ix.last_flags.tolist()
    # the same queries, one block at a time and alone: identical bytes (batch independence)
    _run(docs, doc_ptr, tokens, v, queries[:9], 10, 2048, algo="head", ix=ix)
    # reasons are reported in the upper bits: unknown-only query -> not certified (all-zero ties), long query -> bit 1
    _run(docs, doc_ptr, tokens, v, [[-1], list(range(40))] + queries[:8], 10, 2048, algo="head_nofallback",
         allow_flags=True, ix=ix)
    assert ix.last_flags[0] != 0 and (ix.last_flags[1] & 2)


def test_bm25_head_path_needs_its_index():
    from classmate_rag_b200 import lexical, ops
    docs, doc_ptr, tokens, v = zipf_corpus(seed=5, n_docs=4000, vocab=100, mean_len=10)
    plain = lexical.build_lexical_index(doc_ptr, tokens, v, device="cuda", tile_docs=2048, head_terms=0)
    assert plain.head_mat is None
    qt, qp = lexical.pack_queries(zipf_queries(1, 9, 100))
    with pytest.raises(RuntimeError):
        ops.bm25_topk(plain, qt.cuda(), qp.cuda(), 8, algo="head")
    big = lexical.build_lexical_index(doc_ptr, tokens, v, device="cuda", tile_docs=4096)
    with pytest.raises(RuntimeError):
        ops.bm25_topk(big, qt.cuda(), qp.cuda(), 8, algo="head")
    # auto falls back to the exact kernel for both
    _run(docs, doc_ptr, tokens, v, zipf_queries(1, 9, 100), 8, 4096)


def test_bm25_head_path_repeated_runs_identical():
    from classmate_rag_b200 import lexical, ops
    docs, doc_ptr, tokens, v = zipf_corpus(seed=31, n_docs=30000, vocab=800, mean_len=25)
    ix = lexical.build_lexical_index(doc_ptr, tokens, v, device="cuda", tile_docs=2048)
    qt, qp = lexical.pack_queries(zipf_queries(3, 48, 800))
    qt, qp = qt.cuda(), qp.cuda()
    first = None
    for _ in range(20):
        out = [t.cpu().numpy().tobytes() for t in ops.bm25_topk(ix, qt, qp, 10, algo="head")]
        first = first or out
        assert out == first


@pytest.mark.parametrize("n_docs,vocab,keep_every", [(5000, 300, 3), (70000, 20000, 16), (3000, 50, 1), (3000, 50, 0)])
def test_masked_df_matches_segmented_sum(n_docs, vocab, keep_every):
    """cmr_masked_df (subset document frequencies + first passing posting of every term, one pass over
    the CSR) against a cumulative-sum restatement; covers empty terms, an all-pass and an empty mask."""
    from classmate_rag_b200 import lexical, ops, synth
    doc_ptr, tokens = synth.lexical_corpus(n_docs, vocab, 24, "cuda")
    lex = lexical.build_lexical_index(doc_ptr, tokens, vocab)
    g = torch.Generator(device="cpu").manual_seed(n_docs + keep_every)
    if keep_every == 0:
        mask = torch.zeros(n_docs, dtype=torch.uint8)
    elif keep_every == 1:
        mask = torch.ones(n_docs, dtype=torch.uint8)
    else:
        mask = (torch.randint(0, keep_every, (n_docs,), generator=g) == 0).to(torch.uint8)
    df, first = ops.masked_df(lex.term_ptr, lex.post_doc, mask.cuda())
    torch.cuda.synchronize()
    tp = lex.term_ptr.cpu().numpy()
    hit = mask.numpy()[lex.post_doc.cpu().numpy()].astype(np.int64)
    cs = np.concatenate([[0], np.cumsum(hit)])
    want_df = cs[tp[1:]] - cs[tp[:-1]]
    assert np.array_equal(df.cpu().numpy().astype(np.int64), want_df)
    got_first = first.cpu().numpy()
    for t in np.nonzero(want_df > 0)[0][:2000]:
        seg = hit[tp[t]:tp[t + 1]]
        assert got_first[t] == tp[t] + int(np.argmax(seg))
    assert np.all(got_first[want_df == 0] >= 0x7F7F7F7F)
