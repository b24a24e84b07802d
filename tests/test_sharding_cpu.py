"""CPU (gloo, world_size 2): the shard exchange protocol and the distributed
index statistics give exactly the unsharded result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.synth_small import zipf_corpus


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _np_merge(scores, ids, counts):
    g, b, k = scores.shape
    out_s = torch.zeros((b, k), dtype=torch.float64)
    out_i = torch.full((b, k), -1, dtype=torch.int64)
    out_c = torch.zeros((b,), dtype=torch.int32)
    for q in range(b):
        items = [(float(scores[p, q, r]), int(ids[p, q, r])) for p in range(g) for r in range(int(counts[p, q]))]
        items.sort(key=lambda x: (-x[0], x[1]))
        items = items[:k]
        out_c[q] = len(items)
        for r, (s, i) in enumerate(items):
            out_s[q, r], out_i[q, r] = s, i
    return out_s, out_i, out_c


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from classmate_rag_b200 import lexical
    from classmate_rag_b200.sharding import ShardComm, global_corpus_stats, shard_range
    try:
        # ---- exchange protocol: per-shard exact top-k -> merged == unsharded top-k
        rng = np.random.default_rng(0)
        n, b, k = 1000, 3, 10
        scores_all = np.round(rng.random((b, n)), 2)          # coarse: ties across shards
        lo, hi = shard_range(n, rank, world)
        loc_s = torch.zeros((b, k), dtype=torch.float64)
        loc_i = torch.full((b, k), -1, dtype=torch.int64)
        loc_c = torch.zeros((b,), dtype=torch.int32)
        for q in range(b):
            order = np.lexsort((np.arange(lo, hi), -scores_all[q, lo:hi]))[:k]
            loc_c[q] = len(order)
            loc_s[q, :len(order)] = torch.from_numpy(scores_all[q, lo:hi][order])
            loc_i[q, :len(order)] = torch.from_numpy(order + lo)
        comm = ShardComm(merge_fn=_np_merge)
        m_s, m_i, m_c, m_f = comm.merge_topk(loc_s, loc_i, loc_c, torch.zeros(b, dtype=torch.int32))
        for q in range(b):
            want = np.lexsort((np.arange(n), -scores_all[q]))[:k]
            assert m_i[q].tolist() == want.tolist()
            assert m_s[q].tolist() == scores_all[q][want].tolist()
        # ---- the single-exchange all-gather of packed messages: [G, B, bytes], rank-major
        msg = torch.full((3, 48), rank + 1, dtype=torch.uint8)
        got = comm.all_gather_bytes(msg)
        assert got.shape == (world, 3, 48) and all(int(got[r].min()) == r + 1 == int(got[r].max()) for r in range(world))
        # ---- pool rows: exactly one rank holds each row
        rows = torch.zeros((4, 16), dtype=torch.bfloat16)
        full = (torch.arange(64, dtype=torch.float32).reshape(4, 16) / 7).to(torch.bfloat16)
        rows[rank::world] = full[rank::world]
        got = comm.sum_rows(rows.clone())
        assert torch.equal(got.view(torch.int16), full.view(torch.int16))
        # ---- distributed index statistics == unsharded statistics
        docs, doc_ptr, tokens, v = zipf_corpus(seed=12, n_docs=600, vocab=90, mean_len=8)
        dlo, dhi = shard_range(600, rank, world)
        t_lo, t_hi = int(doc_ptr[dlo]), int(doc_ptr[dhi])
        st = global_corpus_stats(doc_ptr[dlo:dhi + 1] - doc_ptr[dlo], tokens[t_lo:t_hi], v, doc_lo=dlo,
                                 n_docs_total=600, token_offset=t_lo)
        ref = lexical.corpus_stats(doc_ptr, tokens, v)
        assert st.n_docs == ref.n_docs and st.total_tokens == ref.total_tokens
        assert np.array_equal(st.df, ref.df) and np.array_equal(st.vocab_order, ref.vocab_order)
        sh = lexical.build_lexical_index(doc_ptr[dlo:dhi + 1] - doc_ptr[dlo], tokens[t_lo:t_hi], v, device="cpu",
                                         tile_docs=512, stats=st)
        full_ix = lexical.build_lexical_index(doc_ptr, tokens, v, device="cpu", tile_docs=512)
        assert np.array_equal(sh.idf_host, full_ix.idf_host) and sh.avgdl == full_ix.avgdl
        ret[rank] = "ok"
    except Exception as exc:  # surface the failure to the parent
        import traceback
        ret[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gloo_exchange_and_stats():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        for r in range(world):
            assert ret.get(r) == "ok", ret.get(r)


def test_shard_range_covers_everything():
    from classmate_rag_b200.sharding import shard_range
    for n in (0, 1, 15, 16, 17, 1000, 10_000_000):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            assert all(lo % 16 == 0 for lo, _ in spans if lo < n)
