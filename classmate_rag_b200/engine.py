"""Device-level hybrid search engine: the launch sequence of the hot path.

    queries (bf16) --cmr_dense_topk--> pool --cmr_gather_rows/cmr_mmr_select--> dense list
    query terms    --cmr_bm25_topk---> lexical list
    both           --cmr_hybrid_fuse-> (ids, fused, vector_distance, bm25_score)

This is what HybridRetriever.retrieve (reference rag/retrieval/fusion.py:108-167)
drives per question; the engine exposes it batched and asynchronous on the
current CUDA stream, optionally replayed from a CUDA graph so that one query is
one graph launch.  On a row-sharded corpus (one process per GPU) the per-shard
dense pool and BM25 list are packed into ONE message per query, exchanged once (stores into
every rank's receive buffer over NVLink + epoch flags, or an NCCL all-gather) and merged by
cmr_shard_merge (classmate_rag_b200.sharding).

Queries whose dense result could not be certified (CMR_FLAG_UNCERTIFIED) are reported in
``HybridEngine.last_dense_flags``; GraphedSearch / PipelinedSearch read them back with the
results and re-run such a batch on the exhaustive float64 scan before handing it out.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from .lexical import LexicalIndex, pack_queries


@dataclass
class SearchParams:
    """Mirrors the knobs of HybridRetriever (fusion.py:64-78)."""
    top_k: int = 8
    k_vector: int = 8
    k_bm25: int = 8
    rrf_k: int = 60
    weight_vector: float = 1.0
    weight_bm25: float = 1.0
    use_mmr: bool = True
    mmr_lambda: float = 0.5
    mmr_max_pool: int = 24
    hybrid: bool = True

    @property
    def pool(self) -> int:
        return max(self.k_vector, self.mmr_max_pool) if self.use_mmr else self.k_vector


class HybridEngine:
    """One shard (or the whole corpus) resident on one GPU."""

    def __init__(self, emb: torch.Tensor, lex: Optional[LexicalIndex], *, row_offset: int = 0,
                 max_row_norm: float = 1.0, comm=None, overlap: Optional[bool] = None):
        if not emb.is_cuda or emb.dtype != torch.bfloat16:
            raise RuntimeError("emb must be a CUDA bfloat16 matrix (no CPU path)")
        self.emb = emb
        self.lex = lex
        self.row_offset = int(row_offset)
        self.max_row_norm = float(max_row_norm)
        self.comm = comm  # classmate_rag_b200.sharding.ShardComm or None
        self.device = emb.device
        self._dense_ws = {}
        self._bm_buf = {}
        # overlap: for batches on the tcgen05 path (> 8 queries) the BM25 kernels run on a side
        # stream, launched first; the dense pass starts on the SMs BM25 has already left, so the
        # tail of one overlaps the head of the other (10M x 768, batch 32: 3.73 -> 3.56 ms per
        # step).  Truly concurrent execution does not pay: both want the SM's shared memory and
        # registers (tools/overlap_sweep.py: 5.0 ms with BM25 held to 3 CTAs per SM), and for a
        # single query the persistent scan kernel loses more than BM25 gains (p50 2.5 -> 3.0 ms),
        # so small batches stay serial.  CMRAG_OVERLAP=0 serialises everything.
        if overlap is None:
            overlap = os.environ.get("CMRAG_OVERLAP", "1") != "0"
        self.overlap = bool(overlap)
        self.last_dense_flags = None
        self.last_exchange_timeout = None
        self.bm25_first = os.environ.get("CMRAG_BM25_ORDER", "first") != "after"
        self._side = {}

    def _fork_lexical(self, q_terms, q_ptr, k, lex_mask, slot=0):
        """BM25 top-k on the side stream, forked from the current stream at the point of this
        call; returns (launch, join): launch() enqueues the kernels on the side stream and
        returns their result, join() makes the current stream wait for them.  Works eagerly and
        under CUDA-graph capture (the fork/join become graph edges)."""
        cur = torch.cuda.current_stream(self.device)
        side = self._side.get(slot)
        if side is None:
            side = self._side[slot] = torch.cuda.Stream(device=self.device)
        side.wait_stream(cur)

        def launch():
            with torch.cuda.stream(side):
                return self.lexical_topk(q_terms, q_ptr, k, lex_mask, slot)
        return launch, (lambda: cur.wait_stream(side))

    def _dense_and_lexical(self, q_bf16, pool, dense_mask, hybrid, q_terms, q_ptr, k_bm25, lex_mask, dense_algo="auto",
                           slot=0):
        """The two retrievers of a step; BM25 on the side stream when overlap applies."""
        if not hybrid:
            return self.dense_pool(q_bf16, pool, dense_mask, dense_algo, slot), None, None
        if not (self.overlap and q_bf16.shape[0] > 8):
            return (self.dense_pool(q_bf16, pool, dense_mask, dense_algo, slot), None,
                    lambda: self.lexical_topk(q_terms, q_ptr, k_bm25, lex_mask, slot))
        launch, join = self._fork_lexical(q_terms, q_ptr, k_bm25, lex_mask, slot)
        if self.bm25_first:
            bm = launch()
            dense = self.dense_pool(q_bf16, pool, dense_mask, dense_algo, slot)
        else:
            dense = self.dense_pool(q_bf16, pool, dense_mask, dense_algo, slot)
            bm = launch()
        return dense, join, lambda: bm

    # -- stage helpers -------------------------------------------------------
    def _cert_eps(self, dim: int) -> float:
        return ops.dense_cert_eps(dim, 1.0, self.max_row_norm)

    def dense_pool(self, q_bf16: torch.Tensor, k: int, row_mask: Optional[torch.Tensor] = None, algo: str = "auto",
                   slot: int = 0):
        """``slot``: which set of scratch / result buffers to use -- calls with different slots may be
        in flight at the same time on different streams (PipelinedSearch)."""
        n, d = self.emb.shape
        b = q_bf16.shape[0]
        key = (n, d, b, k, slot)
        ws = self._dense_ws.get(key)
        if ws is None:
            ws = self._dense_ws[key] = ops.DenseWorkspace(n, d, b, k, self.device)
        if algo == "wide":
            # the widest over-selection one pass certifies (top-120, KP = 128), cut back to k: serves
            # queries whose top k sits inside a cluster of exact duplicates (tools/dup_heavy.py)
            kw = max(k, ops.WIDE_K)
            wkey = (n, d, b, kw, slot)
            wws = self._dense_ws.get(wkey)
            if wws is None:
                wws = self._dense_ws[wkey] = ops.DenseWorkspace(n, d, b, kw, self.device)
            s_w, i_w, c_w, f_w = ops.dense_topk(self.emb, q_bf16, kw, row_mask=row_mask, row_offset=self.row_offset,
                                                cert_eps=self._cert_eps(d), workspace=wws)
            out = (s_w[:, :k].contiguous(), i_w[:, :k].contiguous(), torch.clamp(c_w, max=k), f_w)
        else:
            out = ops.dense_topk(self.emb, q_bf16, k, row_mask=row_mask, row_offset=self.row_offset,
                                 cert_eps=self._cert_eps(d), workspace=ws, algo=algo)
        if self.comm is not None:
            out = self.comm.merge_topk(*out)
        return out

    def lexical_topk(self, q_terms: torch.Tensor, q_ptr: torch.Tensor, k: int,
                     row_mask: Optional[torch.Tensor] = None, slot: int = 0):
        b = q_ptr.numel() - 1
        key = (b, k, slot)
        buf = self._bm_buf.get(key)
        if buf is None:
            import ctypes as C
            from . import _lib
            st = self.lex.struct()
            with torch.cuda.device(self.device):
                nbytes = _lib.load().cmr_bm25_workspace_bytes(C.byref(st), b, k)
            if nbytes == 0:
                raise ValueError("unsupported bm25 shape: " + _lib.last_error())
            buf = self._bm_buf[key] = ops.TopkBuffers(b, k, nbytes, self.device)
        out = ops.bm25_topk(self.lex, q_terms, q_ptr, k, row_mask=row_mask, row_offset=self.row_offset,
                            buffers=buf)
        if self.comm is not None:
            out = self.comm.merge_topk(*out)
        return out

    def pool_rows(self, ids: torch.Tensor) -> torch.Tensor:
        rows = ops.gather_rows(self.emb, ids, row_offset=self.row_offset)
        if self.comm is not None:
            rows = self.comm.sum_rows(rows)
        return rows

    # -- the hot path --------------------------------------------------------
    def search(self, q_bf16: torch.Tensor, q_terms: Optional[torch.Tensor], q_ptr: Optional[torch.Tensor],
               p: SearchParams, *, dense_mask: Optional[torch.Tensor] = None,
               lex_mask: Optional[torch.Tensor] = None, stage_events: Optional[list] = None,
               dense_algo: str = "auto", slot: int = 0):
        """Returns device tensors (ids i64 [B,top_k], fused f64, vector_distance f64
        (NaN = None), bm25_score f64 (NaN = None), counts i32 [B]); nothing is
        synchronised.  ``self.last_dense_flags`` (int32 [B], device) is non-zero for queries
        whose dense pool could not be certified: re-run those with ``dense_algo="wide"`` (the
        same kernels over-selecting 128 candidates) and, if still flagged, ``"exact"`` (the
        exhaustive float64 scan); GraphedSearch.result() does.  ``stage_events``: a list that
        receives timing events at the stage boundaries (start, dense done, MMR done, BM25
        done, fused) -- single-shard path only."""
        def mark():
            if stage_events is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                stage_events.append(e)
        if q_bf16.dim() == 1:
            q_bf16 = q_bf16[None]
        hybrid, k_vec, pool = self._shape(p, q_terms)
        if self.comm is not None:
            return self._search_sharded(q_bf16, q_terms, q_ptr, p, hybrid, k_vec, pool, dense_mask, lex_mask, dense_algo,
                                        slot)
        mark()
        (scores, ids, counts, flags), join, lexical = self._dense_and_lexical(
            q_bf16, pool, dense_mask, hybrid, q_terms, q_ptr, p.k_bm25, lex_mask, dense_algo, slot)
        self.last_dense_flags = flags
        mark()
        if p.use_mmr:
            rows = self.pool_rows(ids)
            v_ids, v_sims, v_cnt = ops.mmr_select(rows, scores, ids, counts, min(k_vec, pool), p.mmr_lambda)
        else:
            v_ids, v_sims, v_cnt = ids, scores, counts
        mark()
        bm = None
        if hybrid:
            if join is not None:
                join()
            b_sc, b_ids, b_cnt, _ = lexical()
            bm = (b_ids, b_sc, b_cnt)
        mark()
        out = ops.hybrid_fuse((v_ids, v_sims, v_cnt), bm, top_k=p.top_k, rrf_k=p.rrf_k,
                              w_vec=p.weight_vector if hybrid else 1.0, w_bm=p.weight_bm25)
        mark()
        return out


    # -- the same step in two halves (PipelinedSearch runs them on different streams) ----------
    def _shape(self, p: SearchParams, q_terms):
        hybrid = p.hybrid and self.lex is not None and q_terms is not None
        k_vec = p.k_vector if hybrid else max(p.top_k, p.k_vector)
        pool = max(k_vec, p.mmr_max_pool) if p.use_mmr else k_vec
        pool = min(pool, 64) if p.use_mmr else pool
        return hybrid, k_vec, pool

    def search_heavy(self, q_bf16: torch.Tensor, q_terms: Optional[torch.Tensor], q_ptr: Optional[torch.Tensor],
                     p: SearchParams, *, dense_mask: Optional[torch.Tensor] = None,
                     lex_mask: Optional[torch.Tensor] = None, dense_algo: str = "auto", slot: int = 0):
        """First half of search(): the two scans over this shard (dense pool + BM25 list, BM25 on its
        side stream), nothing exchanged.  Returns the state search_tail() continues from; its tensors
        live in the slot's buffers until the slot's next search_heavy()."""
        if q_bf16.dim() == 1:
            q_bf16 = q_bf16[None]
        hybrid, k_vec, pool = self._shape(p, q_terms)
        comm, self.comm = self.comm, None       # local lists only: the tail does the one exchange
        try:
            dense, join, lexical = self._dense_and_lexical(q_bf16, pool, dense_mask, hybrid, q_terms, q_ptr,
                                                           p.k_bm25, lex_mask, dense_algo, slot)
            bm_local = None
            if hybrid:
                if join is not None:
                    join()
                b_sc, b_ids, b_cnt, _ = lexical()
                bm_local = (b_sc, b_ids, b_cnt)
        finally:
            self.comm = comm
        return {"dense": dense, "bm": bm_local, "hybrid": hybrid, "k_vec": k_vec, "pool": pool,
                "b": q_bf16.shape[0], "slot": slot}

    def search_tail(self, state, p: SearchParams):
        """Second half of search(): (exchange + merge,) MMR, fusion.  Same results as search()."""
        dense, bm_local, hybrid, k_vec, pool = state["dense"], state["bm"], state["hybrid"], state["k_vec"], state["pool"]
        if self.comm is not None:
            return self._sharded_tail(dense, bm_local, p, hybrid, k_vec, pool, state["b"], state["slot"])
        scores, ids, counts, flags = dense
        self.last_dense_flags = flags
        if p.use_mmr:
            rows = self.pool_rows(ids)
            v_ids, v_sims, v_cnt = ops.mmr_select(rows, scores, ids, counts, min(k_vec, pool), p.mmr_lambda)
        else:
            v_ids, v_sims, v_cnt = ids, scores, counts
        bm = (bm_local[1], bm_local[0], bm_local[2]) if hybrid else None
        return ops.hybrid_fuse((v_ids, v_sims, v_cnt), bm, top_k=p.top_k, rrf_k=p.rrf_k,
                               w_vec=p.weight_vector if hybrid else 1.0, w_bm=p.weight_bm25)

    def _search_sharded(self, q_bf16, q_terms, q_ptr, p, hybrid, k_vec, pool, dense_mask, lex_mask, dense_algo="auto",
                        slot=0):
        """Row-sharded step with ONE collective: local dense pool + local BM25 list ->
        cmr_shard_pack -> all-gather -> cmr_shard_merge -> MMR -> fuse."""
        comm, self.comm = self.comm, None       # the stage helpers must not exchange on their own
        try:
            dense, join, lexical = self._dense_and_lexical(q_bf16, pool, dense_mask, hybrid, q_terms, q_ptr,
                                                           p.k_bm25, lex_mask, dense_algo, slot)
            bm_local = None
            if hybrid:
                if join is not None:
                    join()
                b_sc, b_ids, b_cnt, _ = lexical()
                bm_local = (b_sc, b_ids, b_cnt)
        finally:
            self.comm = comm
        return self._sharded_tail(dense, bm_local, p, hybrid, k_vec, pool, q_bf16.shape[0], slot)

    def _sharded_tail(self, dense, bm_local, p, hybrid, k_vec, pool, b, slot):
        """Exchange + merge + MMR + fuse of a row-sharded step (everything after the two local scans)."""
        comm = self.comm
        dim = self.emb.shape[1] if p.use_mmr else 0
        kb = p.k_bm25 if hybrid else 0
        peer = comm.peer_exchange(self.device, b * ops.shard_msg_bytes(pool, kb, dim), self.emb, self.row_offset, slot)
        if peer is not None:
            # stores into every rank's receive buffer over NVLink + flags: no collective launch
            # with every shard's matrix in shared memory (sharding.shared_rows) no rows travel: the
            # merge pulls the merged pool's rows from their owners
            px = peer.struct_for(self.emb)
            ops.shard_exchange_pack(dense, bm_local, self.emb if p.use_mmr else None, px, row_offset=self.row_offset)
            d_s, d_i, d_c, d_f, rows, g_bs, g_bi, g_bc = ops.shard_exchange_merge(
                peer.recv, peer.flags, px, peer.timeout, b, pool, kb, dim)
            self.last_exchange_timeout = peer.timeout   # int32 [1], device: non-zero = a rank never arrived
        else:
            msg = ops.shard_pack(dense, bm_local, self.emb if p.use_mmr else None, row_offset=self.row_offset)
            gathered = comm.all_gather_bytes(msg)
            d_s, d_i, d_c, d_f, rows, g_bs, g_bi, g_bc = ops.shard_merge(gathered, pool, kb, dim)
        self.last_dense_flags = d_f
        if p.use_mmr:
            v_ids, v_sims, v_cnt = ops.mmr_select(rows, d_s, d_i, d_c, min(k_vec, pool), p.mmr_lambda)
        else:
            v_ids, v_sims, v_cnt = d_i, d_s, d_c
        bm = (g_bi, g_bs, g_bc) if hybrid else None
        return ops.hybrid_fuse((v_ids, v_sims, v_cnt), bm, top_k=p.top_k, rrf_k=p.rrf_k,
                               w_vec=p.weight_vector if hybrid else 1.0, w_bm=p.weight_bm25)


class GraphedSearch:
    """One fixed-shape hybrid search captured in a CUDA graph: host inputs are
    copied into static device buffers, the whole kernel sequence replays as one
    graph launch, results land in static pinned host buffers.

    This is the end-to-end call with HOST buffers that bench.py's ``e2e`` times."""

    def __init__(self, engine: HybridEngine, p: SearchParams, n_queries: int, max_terms: int = 64,
                 graph_collectives: bool = True, stream: Optional[torch.cuda.Stream] = None, slot: int = 0,
                 heavy_stream: Optional[torch.cuda.Stream] = None):
        self.engine, self.p, self.b = engine, p, n_queries
        self.slot = slot   # the engine's buffer set this object runs on (see PipelinedSearch)
        # split form (PipelinedSearch): the two scans are one graph replayed on ``heavy_stream``, everything
        # after them (exchange, merge, MMR, fusion, result copies) a second graph on this object's own stream
        self.heavy_stream = heavy_stream
        self.graph_h = None
        self.heavy_done = torch.cuda.Event()
        self.graph_collectives = graph_collectives
        dev = engine.device
        d = engine.emb.shape[1]
        self.hybrid = p.hybrid and engine.lex is not None
        self.max_terms = max_terms
        # static device inputs and their pinned host staging: [queries f32 | term ids i32 | offsets i32] in ONE
        # buffer on each side, so a launch is one host-to-device copy
        n_q, n_t, n_p = n_queries * d, (max(1, n_queries * max_terms) + 3) // 4 * 4, n_queries + 1
        assert n_q % 4 == 0   # dim is a multiple of 8: every part starts 16-byte aligned
        self.d_in = torch.zeros((n_q + n_t + n_p,), dtype=torch.int32, device=dev)
        self.h_in = torch.zeros((n_q + n_t + n_p,), dtype=torch.int32).pin_memory()

        def carve(buf):
            return (buf[:n_q].view(torch.float32).view(n_queries, d), buf[n_q:n_q + n_t], buf[n_q + n_t:])
        self.q_f32, self.q_terms, self.q_ptr = carve(self.d_in)
        self.h_q, self.h_terms, self.h_ptr = carve(self.h_in)
        self.q_f32[:, 0] = 1.0   # capture warm-ups run on this buffer: a unit vector, not the all-ties zero query
        self.q_terms.fill_(-1)
        self.h_terms.fill_(-1)
        # several GraphedSearch objects may share one stream (they share the engine's scratch
        # buffers, so their device work must be serialised anyway): see PipelinedSearch
        self.stream = stream if stream is not None else torch.cuda.Stream(device=dev)
        self.done = torch.cuda.Event()
        self.graph = None
        self._capture()

    def _run(self):
        q_bf16 = ops.f32_to_bf16(self.q_f32)
        return self.engine.search(q_bf16, self.q_terms if self.hybrid else None,
                                  self.q_ptr if self.hybrid else None, self.p, slot=self.slot)

    def _capture(self):
        with torch.cuda.device(self.engine.device), torch.cuda.stream(self.stream):
            for _ in range(2):  # warm-up: one-time attribute / occupancy calls, allocations
                self.out = self._run()
            self.stream.synchronize()
            if self.heavy_stream is not None:
                q_terms, q_ptr = (self.q_terms, self.q_ptr) if self.hybrid else (None, None)
                gh = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gh, stream=self.stream):
                    self._state = self.engine.search_heavy(ops.f32_to_bf16(self.q_f32), q_terms, q_ptr, self.p,
                                                           slot=self.slot)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.stream, pool=gh.pool()):
                    self.out = self.engine.search_tail(self._state, self.p)
                self.graph_h, self.graph = gh, g
                self._peer_ref = self._current_peer()
            elif self.engine.comm is None or self.graph_collectives:
                # with a communicator the graph holds the step's single NCCL all-gather too
                # (every rank captures and replays the same sequence)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.stream):
                    self.out = self._run()
                self.graph = g
                self._peer_ref = self._current_peer()
            ids, fused, vd, bm, cnt = self.out
            # dense certificate flags (+ the peer exchange's timeout word) travel with the results
            self.flags = self.engine.last_dense_flags
            self.timeout = self.engine.last_exchange_timeout
            self.h_flags = torch.zeros(self.flags.shape, dtype=self.flags.dtype).pin_memory()
            self.h_timeout = torch.zeros((1,), dtype=torch.int32).pin_memory()
            self.reruns = 0
            # hybrid_fuse hands back its five results as views of one allocation: one device-to-host copy
            self.d_res = ops.fused_result_buffer(self.out)
            if self.d_res is not None:
                self.h_res = torch.empty(self.d_res.shape, dtype=torch.uint8).pin_memory()
                n8 = ids.numel() * 8
                self.h_ids = self.h_res[:n8].view(ids.dtype).view(ids.shape)
                self.h_fused = self.h_res[n8:2 * n8].view(fused.dtype).view(fused.shape)
                self.h_vd = self.h_res[2 * n8:3 * n8].view(vd.dtype).view(vd.shape)
                self.h_bm = self.h_res[3 * n8:4 * n8].view(bm.dtype).view(bm.shape)
                self.h_cnt = self.h_res[4 * n8:].view(cnt.dtype)
            else:
                self.h_ids = torch.empty(ids.shape, dtype=ids.dtype).pin_memory()
                self.h_fused = torch.empty(fused.shape, dtype=fused.dtype).pin_memory()
                self.h_vd = torch.empty(vd.shape, dtype=vd.dtype).pin_memory()
                self.h_bm = torch.empty(bm.shape, dtype=bm.dtype).pin_memory()
                self.h_cnt = torch.empty(cnt.shape, dtype=cnt.dtype).pin_memory()

    @property
    def h2d_bytes(self) -> int:
        n = self.h_q.numel() * 4
        if self.hybrid:
            n += self.h_terms.numel() * 4 + self.h_ptr.numel() * 4
        return n

    @property
    def d2h_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.h_ids, self.h_fused, self.h_vd, self.h_bm, self.h_cnt,
                                                          self.h_flags))

    def set_queries(self, q_f32: np.ndarray, term_lists: Optional[Sequence[Sequence[int]]]):
        """Stage host inputs into the pinned buffers (not part of the device work)."""
        self.h_q.copy_(torch.from_numpy(np.ascontiguousarray(q_f32, dtype=np.float32)).reshape(self.h_q.shape))
        if self.hybrid:
            flat, ptr = pack_queries(term_lists)
            if int(ptr[-1]) > self.h_terms.numel():
                raise ValueError("too many query tokens for this graph (raise max_terms)")
            self.h_terms[: int(ptr[-1])].copy_(flat[: int(ptr[-1])])
            self.h_ptr.copy_(ptr)

    def _current_peer(self):
        peers = getattr(self.engine.comm, "peers", None)
        return None if not peers else peers.get(self.slot)

    def _check_graph(self):
        if self.graph is not None and self._current_peer() is not getattr(self, "_peer_ref", None):
            raise RuntimeError("the peer-exchange buffers were re-allocated after this graph was captured; "
                               "build a new GraphedSearch")

    def launch(self):
        """H2D copies + graph replay + D2H copies on the engine's stream (asynchronous)."""
        self._check_graph()
        with torch.cuda.device(self.engine.device), torch.cuda.stream(self._input_stream()):
            if self.hybrid:
                self.d_in.copy_(self.h_in, non_blocking=True)
            else:
                self.q_f32.copy_(self.h_q, non_blocking=True)
            self._replay_heavy()
        with torch.cuda.device(self.engine.device), torch.cuda.stream(self.stream):
            self._replay_rest()
            self._copy_out(self.out)
            self.h_flags.copy_(self.flags, non_blocking=True)
            if self.timeout is not None:
                self.h_timeout.copy_(self.timeout, non_blocking=True)
            self.done.record(self.stream)

    def _input_stream(self) -> torch.cuda.Stream:
        """The stream the inputs are written on: the heavy stream in the split form.  There the slot's
        previous tail must have finished first (it reads the lists the scans are about to overwrite)."""
        if self.graph_h is None:
            return self.stream
        self.heavy_stream.wait_event(self.done)
        return self.heavy_stream

    def _replay_heavy(self):
        """(current stream = _input_stream())  Split form: the scans, then the hand-over event."""
        if self.graph_h is not None:
            self.graph_h.replay()
            self.heavy_done.record(self.heavy_stream)

    def _replay_rest(self):
        """(current stream = self.stream)  The whole step, or in the split form everything after the scans."""
        if self.graph_h is not None:
            self.stream.wait_event(self.heavy_done)
        if self.graph is not None:
            self.graph.replay()
        else:
            self.out = self._run()

    def _copy_out(self, out):
        res = ops.fused_result_buffer(out) if self.d_res is not None else None
        if res is not None and res.shape == self.h_res.shape:
            self.h_res.copy_(res, non_blocking=True)
            return
        for h, t in zip((self.h_ids, self.h_fused, self.h_vd, self.h_bm, self.h_cnt), out):
            h.copy_(t, non_blocking=True)

    def _rerun_certified(self):
        """A query of the batch could not be certified by the fp32 pass (a cluster of exact duplicates
        around rank k): the batch is run again, eagerly, with the widest over-selection, and if that
        still leaves a flag, with the exhaustive float64 scan as the dense stage (same inputs -- they are
        still in the static buffers -- same exchange on every rank, since the merged flags are identical
        everywhere)."""
        for algo in ("wide", "exact"):
            self.reruns += 1
            with torch.cuda.device(self.engine.device), torch.cuda.stream(self.stream):
                q_bf16 = ops.f32_to_bf16(self.q_f32)
                out = self.engine.search(q_bf16, self.q_terms if self.hybrid else None,
                                         self.q_ptr if self.hybrid else None, self.p, dense_algo=algo, slot=self.slot)
                self._copy_out(out)
                self.h_flags.copy_(self.engine.last_dense_flags, non_blocking=True)
                self.stream.synchronize()
            if int(self.h_flags.sum()) == 0:
                return
        raise RuntimeError("the exhaustive dense scan returned an uncertified result")

    def launch_resident(self, q_f32: torch.Tensor, q_terms: Optional[torch.Tensor] = None,
                        q_ptr: Optional[torch.Tensor] = None):
        """Replay with inputs that already live in HBM (device tensors of the captured shapes:
        q_f32 [B, dim] float32, q_terms int32 (at most B * max_terms), q_ptr int32 [B + 1]).
        Results stay on the device (``self.out``); nothing is copied to the host."""
        self._check_graph()
        # the inputs were produced on the caller's stream and may be temporaries: order this
        # stream after it and keep their memory from being reused while the copies are pending
        cur = torch.cuda.current_stream(self.engine.device)
        ins = self._input_stream()
        if cur != ins:
            ins.wait_stream(cur)
            for t in (q_f32, q_terms, q_ptr):
                if t is not None:
                    t.record_stream(ins)
        with torch.cuda.device(self.engine.device), torch.cuda.stream(ins):
            if q_f32.data_ptr() != self.q_f32.data_ptr():   # an encoder may have written the buffer itself
                self.q_f32.copy_(q_f32, non_blocking=True)
            if self.hybrid:
                n = q_terms.numel()
                if n > self.q_terms.numel():
                    raise ValueError("too many query tokens for this graph (raise max_terms)")
                self.q_terms[:n].copy_(q_terms, non_blocking=True)
                self.q_ptr.copy_(q_ptr, non_blocking=True)
            self._replay_heavy()
        with torch.cuda.device(self.engine.device), torch.cuda.stream(self.stream):
            self._replay_rest()
            self.done.record(self.stream)
        return self.out

    def result(self):
        """Wait for the last launch() of THIS object and hand back its pinned result buffers.
        Fails loudly when a rank never arrived at the peer exchange; re-runs the batch (wider
        over-selection, then the exhaustive scan) when the dense certificate failed for a query."""
        self.done.synchronize()
        if int(self.h_timeout[0]) != 0:
            raise RuntimeError("shard exchange timed out: a rank did not deliver its message (results are invalid)")
        if int(self.h_flags.sum()) != 0:
            self._rerun_certified()
        return self.h_ids.numpy(), self.h_fused.numpy(), self.h_vd.numpy(), self.h_bm.numpy(), self.h_cnt.numpy()

    def __call__(self, q_f32: np.ndarray, term_lists=None):
        self.set_queries(q_f32, term_lists)
        self.launch()
        return self.result()


class PipelinedSearch:
    """Throughput form of the search call: two GraphedSearch objects in their split form, each with
    its own set of engine buffers (slot 0 / 1), used alternately.  While the device works on
    batch i the host stages batch i+1 into the other object's pinned buffers and enqueues it.  On
    the device the scan graphs of consecutive batches replay in order on ONE stream (on a sharded
    engine: the same kernel order on every rank) and the latency-bound tail of a batch (candidate
    merge, shard exchange and the wait for the slowest rank, MMR, fusion, result copy -- small
    grids) replays on the slot's own high-priority stream beside the scans of the next batch
    instead of leaving the SMs idle.  Events hand over: scans -> tail, tail -> the slot's next scans.

        ps = PipelinedSearch(engine, params, batch)
        for q, terms in batches:
            prev = ps.submit(q, terms)        # results of the batch submitted before (or None)
        last = ps.drain()

    ``launch_resident`` is the same rotation for inputs that already live in HBM.  With an NCCL
    (not peer-memory) shard exchange both objects share one stream and one buffer set: collectives
    of two steps must not be in flight at once."""

    def __init__(self, engine: HybridEngine, p: SearchParams, n_queries: int, max_terms: int = 64):
        comm = engine.comm
        self.independent = (comm is None or bool(getattr(comm, "peer_memory", False))) and \
            os.environ.get("CMRAG_PIPELINE", "1") != "0"
        self.heavy_stream = None
        if self.independent:
            # the scans of consecutive steps run in order on ONE stream (the same order on every rank of a
            # sharded engine); each slot's tail has its own high-priority stream, so its small grids take
            # the SMs a scan kernel frees first
            self.heavy_stream = torch.cuda.Stream(device=engine.device)
            self.slots = [GraphedSearch(engine, p, n_queries, max_terms, slot=i, heavy_stream=self.heavy_stream,
                                        stream=torch.cuda.Stream(device=engine.device, priority=-1))
                          for i in range(2)]
            if comm is not None and not comm.peer_memory:
                # the peer exchange turned out to be unavailable while capturing: fall back to one stream
                self.independent = False
        if not self.independent:
            stream = torch.cuda.Stream(device=engine.device)
            self.slots = [GraphedSearch(engine, p, n_queries, max_terms, stream=stream) for _ in range(2)]
        self.turn = 0
        self.pending: Optional[GraphedSearch] = None
        self.last: Optional[GraphedSearch] = None

    @property
    def h2d_bytes(self) -> int:
        return self.slots[0].h2d_bytes

    @property
    def d2h_bytes(self) -> int:
        return self.slots[0].d2h_bytes

    def submit(self, q_f32: np.ndarray, term_lists=None):
        g = self.slots[self.turn]
        self.turn ^= 1
        g.set_queries(q_f32, term_lists)   # g's previous results were handed out two submits ago
        g.launch()
        prev, self.pending, self.last = self.pending, g, g
        return None if prev is None else tuple(a.copy() for a in prev.result())

    def drain(self):
        prev, self.pending = self.pending, None
        return None if prev is None else tuple(a.copy() for a in prev.result())

    def launch_resident(self, q_f32: torch.Tensor, q_terms: Optional[torch.Tensor] = None,
                        q_ptr: Optional[torch.Tensor] = None):
        """Device-resident rotation: returns the output tensors of this launch (valid until the
        same slot is launched again, i.e. for one more call)."""
        g = self.slots[self.turn]
        self.turn ^= 1
        self.last = g
        return g.launch_resident(q_f32, q_terms, q_ptr)

    def wait(self, stream: Optional[torch.cuda.Stream] = None):
        """Order ``stream`` (default: the current one) after everything launched so far."""
        cur = stream if stream is not None else torch.cuda.current_stream(self.slots[0].engine.device)
        if self.heavy_stream is not None:
            cur.wait_stream(self.heavy_stream)
        for g in self.slots:
            cur.wait_stream(g.stream)
