"""CPU: the CSR index builder reproduces rank_bm25's statistics bit for bit."""
import numpy as np
import torch

from oracle import np_oracle as o
from classmate_rag_b200 import lexical
from tests.synth_small import zipf_corpus


def test_index_statistics_match_bm25okapi():
    docs, doc_ptr, tokens, v = zipf_corpus(seed=4, n_docs=1700, vocab=150, mean_len=12)
    ix = lexical.build_lexical_index(doc_ptr, tokens, v, device="cpu", tile_docs=512)
    bm = o.BM25Okapi([[f"w{t}" for t in d] for d in docs])
    assert ix.avgdl == bm.avgdl
    for t in range(v):
        assert ix.idf_host[t] == (bm.idf.get(f"w{t}") or 0.0)
    assert (ix.idf_host < 0).sum() == 0 or True
    # CSR content
    tp = ix.term_ptr.numpy()
    for t in (0, 1, 7, v - 1):
        lo, hi = tp[t], tp[t + 1]
        want = sorted((i, d.count(t)) for i, d in enumerate(docs) if t in d)
        got = list(zip(ix.post_doc[lo:hi].tolist(), (ix.post_tf[lo:hi].to(torch.int32) & 0xFFFF).tolist()))
        assert got == want
    # float64 impacts: bitwise the rank_bm25 expression
    tf = (ix.post_tf.to(torch.int32) & 0xFFFF).numpy().astype(np.int64)
    dl = ix.doc_len.numpy()[ix.post_doc.numpy()]
    want_imp = tf * (1.5 + 1) / (tf + 1.5 * (1 - 0.75 + 0.75 * dl / bm.avgdl))
    code = (ix.post_pack.numpy().view(np.uint32) >> 16).astype(np.int64)
    assert np.array_equal(ix.imp_table.numpy()[code], want_imp)              # packed: table lookup
    assert np.array_equal(ix.post_pack.numpy().view(np.uint32) & 0xFFFF, ix.post_doc.numpy() % 512)
    wide = lexical.build_lexical_index(doc_ptr, tokens, v, device="cpu", tile_docs=512, fmt="wide")
    assert wide.post_pack is None and np.array_equal(wide.post_imp.numpy(), want_imp)
    # skip table: tile slices partition each posting list by document range
    sk = ix.tile_skip.numpy()
    assert sk.shape == (v, ix.n_tiles + 1)
    for t in (0, 3, 50):
        lo = tp[t]
        for tile in range(ix.n_tiles):
            seg = ix.post_doc[lo + sk[t, tile]: lo + sk[t, tile + 1]].numpy()
            assert ((seg >= tile * 512) & (seg < (tile + 1) * 512)).all()
        assert sk[t, -1] == tp[t + 1] - tp[t]
    # oracle CSR scorer on the built arrays == dict-form scorer
    q = [0, 5, 5, 149, -1]
    want = bm.get_scores([f"w{t}" if t >= 0 else "zzz" for t in q])
    got = o.bm25_scores_csr(tp, ix.post_doc.numpy(), (ix.post_tf.to(torch.int32) & 0xFFFF).numpy(),
                            ix.doc_len.numpy(), ix.idf_host, ix.avgdl, q)
    assert np.array_equal(got, want)


def test_sharded_stats_equal_global():
    docs, doc_ptr, tokens, v = zipf_corpus(seed=5, n_docs=300, vocab=80, mean_len=9)
    full = lexical.build_lexical_index(doc_ptr, tokens, v, device="cpu", tile_docs=512)
    stats = lexical.corpus_stats(doc_ptr, tokens, v)
    half = 150
    lo_ptr = doc_ptr[: half + 1]
    sh0 = lexical.build_lexical_index(lo_ptr, tokens[: int(lo_ptr[-1])], v, device="cpu", tile_docs=512, stats=stats)
    hi_ptr = doc_ptr[half:] - doc_ptr[half]
    sh1 = lexical.build_lexical_index(hi_ptr, tokens[int(doc_ptr[half]):], v, device="cpu", tile_docs=512, stats=stats)
    assert np.array_equal(sh0.idf_host, full.idf_host) and np.array_equal(sh1.idf_host, full.idf_host)
    assert sh0.avgdl == full.avgdl == sh1.avgdl
    assert sh0.n_postings + sh1.n_postings == full.n_postings


def test_empty_corpus_builds():
    ix = lexical.build_lexical_index(torch.zeros(1, dtype=torch.int64), torch.zeros(0, dtype=torch.int32), 5,
                                     device="cpu", tile_docs=512)
    assert ix.n_docs == 0 and ix.n_postings == 0 and ix.n_tiles == 1


def test_dense_columns_hold_the_posting_factors():
    docs, doc_ptr, tokens, v = zipf_corpus(seed=6, n_docs=900, vocab=60, mean_len=12)
    ix = lexical.build_lexical_index(doc_ptr, tokens, v, device="cpu", tile_docs=512, dense_density=0.2)
    assert ix.dense_imp is not None and ix.dense_imp.shape[1] == 900
    df = ix.shard_df_host
    assert sorted(ix.dense_terms.tolist()) == [t for t in range(v) if df[t] >= 0.2 * 900]
    slot = ix.dense_slot.numpy()
    assert (slot >= 0).sum() == len(ix.dense_terms)
    tp = ix.term_ptr.numpy()
    code = (ix.post_pack.numpy().view(np.uint32) >> 16).astype(np.int64)
    for t in ix.dense_terms.tolist():
        col = ix.dense_imp[slot[t]].numpy()
        lo, hi = tp[t], tp[t + 1]
        d = ix.post_doc[lo:hi].numpy()
        assert np.array_equal(col[d], ix.imp_table.numpy()[code[lo:hi]])
        rest = np.ones(900, dtype=bool)
        rest[d] = False
        assert (col[rest] == 0.0).all()
    st = ix.struct()
    assert st.n_dense == len(ix.dense_terms) and st.dense_imp == ix.dense_imp.data_ptr()
    off = lexical.build_lexical_index(doc_ptr, tokens, v, device="cpu", tile_docs=512, dense_density=None)
    assert off.dense_imp is None and off.struct().n_dense == 0
    # algorithmic bytes: a dense token streams its 8-byte column, a sparse one its 4-byte postings
    t_dense, t_sparse = int(ix.dense_terms[0]), int(np.argmin(np.where(df > 0, df, 1 << 30)))
    assert ix.posting_bytes([t_dense, t_sparse, -1]) == 8 * 900 + 4 * int(df[t_sparse])


def test_snapshot_round_trip(tmp_path):
    """N2: save -> load gives the same arrays, scalars (bit for bit) and C struct."""
    docs, doc_ptr, tokens, v = zipf_corpus(seed=7, n_docs=1300, vocab=80, mean_len=10)
    for kw in ({}, {"fmt": "wide", "dense_density": None}):
        ix = lexical.build_lexical_index(doc_ptr, tokens, v, device="cpu", tile_docs=512, **kw)
        lexical.save_lexical_index(ix, tmp_path / "snap")
        back = lexical.load_lexical_index(tmp_path / "snap", "cpu")
        assert (back.n_docs, back.n_terms, back.tile_docs, back.n_tiles) == (ix.n_docs, ix.n_terms, ix.tile_docs, ix.n_tiles)
        assert back.avgdl == ix.avgdl and back.k1 == ix.k1 and back.b == ix.b
        for name in lexical._SNAPSHOT_TENSORS:
            a, b = getattr(ix, name), getattr(back, name)
            assert (a is None) == (b is None), name
            if a is not None:
                assert a.dtype == b.dtype and torch.equal(a, b), name
        assert np.array_equal(back.idf_host, ix.idf_host) and np.array_equal(back.shard_df_host, ix.shard_df_host)
        assert back.struct().n_dense == ix.struct().n_dense and back.struct().n_codes == ix.struct().n_codes
        assert back.posting_bytes([0, 1, 5]) == ix.posting_bytes([0, 1, 5])


def test_pack_queries_shapes_and_edge_cases():
    """pack_queries: CSR of the term lists (duplicates and -1 kept, empty queries allowed, an all-empty
    batch still hands back a valid one-element term array)."""
    from classmate_rag_b200 import lexical
    flat, ptr = lexical.pack_queries([[5, 5, -1], [], [7], (1, 2)])
    assert flat.dtype == torch.int32 and ptr.dtype == torch.int32
    assert flat.tolist() == [5, 5, -1, 7, 1, 2] and ptr.tolist() == [0, 3, 3, 4, 6]
    flat, ptr = lexical.pack_queries([[], []])
    assert ptr.tolist() == [0, 0, 0] and flat.numel() == 1
    flat, ptr = lexical.pack_queries([np.array([3, 4], dtype=np.int64)])
    assert flat.tolist() == [3, 4] and ptr.tolist() == [0, 2]
