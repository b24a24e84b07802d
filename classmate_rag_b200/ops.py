"""Tensor-level calls into libcmrag.so.  torch is plumbing here: device
memory, streams.  All compute is in the CUDA library."""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: classmate_rag_b200 has no CPU path")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


def device_info() -> Tuple[int, int, int]:
    import ctypes as C
    lib = _lib.load()
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    _lib.check(lib.cmr_device_info(C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


_DENSE_ALGO = {"auto": _lib.CMR_DENSE_AUTO, "scan": _lib.CMR_DENSE_SCAN, "mma": _lib.CMR_DENSE_MMA,
                "exact": _lib.CMR_DENSE_EXACT}


def f32_to_bf16(src: torch.Tensor) -> torch.Tensor:
    """fp32 -> bf16 bits (round to nearest even) with the library's kernel."""
    _require_cuda(src, "src")
    if src.dtype != torch.float32:
        raise ValueError("src must be float32")
    dst = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    with torch.cuda.device(src.device):
        _lib.check(_lib.load().cmr_f32_to_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()))
    return dst


class DenseWorkspace:
    """Reusable workspace + output buffers for cmr_dense_topk."""

    def __init__(self, n_rows: int, dim: int, n_queries: int, k: int, device):
        lib = _lib.load()
        with torch.cuda.device(device):
            nbytes = lib.cmr_dense_workspace_bytes(n_rows, dim, n_queries, k)
        if nbytes == 0:
            raise ValueError(f"unsupported dense shape n_rows={n_rows} dim={dim} B={n_queries} k={k}: "
                             + _lib.last_error())
        self.key = (n_rows, dim, n_queries, k)
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.scores = torch.empty((n_queries, k), dtype=torch.float64, device=device)
        self.ids = torch.empty((n_queries, k), dtype=torch.int64, device=device)
        self.counts = torch.empty((n_queries,), dtype=torch.int32, device=device)
        self.flags = torch.empty((n_queries,), dtype=torch.int32, device=device)


def dense_cert_eps(dim: int, q_norm: float = 1.0, max_row_norm: float = 1.0) -> float:
    """The one bound every caller of the fp32 tensor-pipe scores uses:
    |fp32 score - exact| <= dim * 2^-22 * |q| * max|c| (accumulation on the tensor pipe is not
    round-to-nearest, so 2^-24 per addition would not be a bound), with 1 % slack."""
    return dim * 2.0 ** -22 * 1.01 * float(q_norm) * float(max_row_norm)


def dense_topk(emb: torch.Tensor, queries: torch.Tensor, k: int, *, row_mask: Optional[torch.Tensor] = None,
               row_offset: int = 0, cert_eps: Optional[float] = None,
               workspace: Optional[DenseWorkspace] = None, algo: str = "auto"):
    """Exact top-k of queries (bf16 [B, D]) against emb (bf16 [N, D]).

    algo: "auto" | "scan" (HBM-streaming mma.sync scan) | "mma" (tcgen05/TMA GEMM
    with the top-k epilogue) | "exact" (exhaustive float64 scan, the fallback for
    flagged queries); see cmr_dense_topk_ex in include/cmrag.h.

    Returns (scores f64 [B,k], ids i64 [B,k], counts i32 [B], flags i32 [B]) on
    the device, enqueued on the current stream (no synchronisation)."""
    _require_cuda(emb, "emb")
    _require_cuda(queries, "queries")
    if emb.dtype != torch.bfloat16 or queries.dtype != torch.bfloat16:
        raise ValueError("emb and queries must be bfloat16")
    if queries.dim() == 1:
        queries = queries[None, :]
    n_rows, dim = emb.shape
    b = queries.shape[0]
    if queries.shape[1] != dim:
        raise ValueError("query dim mismatch")
    if row_mask is not None:
        _require_cuda(row_mask, "row_mask")
        if row_mask.dtype != torch.uint8 or row_mask.numel() != n_rows:
            raise ValueError("row_mask must be uint8 [n_rows]")
    if cert_eps is None:   # unit rows and queries (the E5 contract); other callers pass their norms' bound
        cert_eps = dense_cert_eps(dim)
    if workspace is None or workspace.key != (n_rows, dim, b, k):
        workspace = DenseWorkspace(n_rows, dim, b, k, emb.device)
    lib = _lib.load()
    with torch.cuda.device(emb.device):
        rc = lib.cmr_dense_topk_ex(emb.data_ptr(), n_rows, dim, queries.data_ptr(), b, k, _ptr(row_mask),
                                   row_offset, float(cert_eps), workspace.scores.data_ptr(),
                                   workspace.ids.data_ptr(), workspace.counts.data_ptr(),
                                   workspace.flags.data_ptr(), workspace.ws.data_ptr(), workspace.ws.numel(),
                                   _stream(), _DENSE_ALGO[algo])
    _lib.check(rc)
    return workspace.scores, workspace.ids, workspace.counts, workspace.flags


class TopkBuffers:
    """Output buffers shared by the top-k entry points."""

    def __init__(self, n_queries: int, k: int, ws_bytes: int, device):
        self.ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=device)
        self.scores = torch.empty((n_queries, k), dtype=torch.float64, device=device)
        self.ids = torch.empty((n_queries, k), dtype=torch.int64, device=device)
        self.counts = torch.empty((n_queries,), dtype=torch.int32, device=device)
        self.flags = torch.empty((n_queries,), dtype=torch.int32, device=device)


_BM25_ALGO = {"auto": _lib.CMR_BM25_AUTO, "exact": _lib.CMR_BM25_EXACT, "head": _lib.CMR_BM25_HEAD,
              "head_nofallback": _lib.CMR_BM25_HEAD_NOFALLBACK}


def bm25_topk(index, q_terms: torch.Tensor, q_ptr: torch.Tensor, k: int, *,
              row_mask: Optional[torch.Tensor] = None, row_offset: int = 0,
              buffers: Optional[TopkBuffers] = None, algo: str = "auto"):
    """Exact BM25 top-k for a batch of tokenised queries (device tensors:
    q_terms int32, q_ptr int32 [B+1]).  Returns (scores f64 [B,k], ids i64 [B,k],
    counts, flags) on the device, enqueued on the current stream.

    algo: "auto" | "exact" (per-query float64 tile kernel) | "head" (head-term matrix on the
    tensor cores + bucketed sparse postings + exact rescoring, flagged queries re-run by the
    exact kernel) | "head_nofallback" (flags left for the caller); see cmr_bm25_topk_ex."""
    import ctypes as C
    for name, t in (("q_terms", q_terms), ("q_ptr", q_ptr)):
        _require_cuda(t, name)
        if t.dtype != torch.int32:
            raise ValueError(f"{name} must be int32")
    b = q_ptr.numel() - 1
    if row_mask is not None:
        _require_cuda(row_mask, "row_mask")
        if row_mask.dtype != torch.uint8 or row_mask.numel() != index.n_docs:
            raise ValueError("row_mask must be uint8 [n_docs]")
    lib = _lib.load()
    st = index.struct()
    with torch.cuda.device(index.device):
        if buffers is None:
            nbytes = lib.cmr_bm25_workspace_bytes(C.byref(st), b, k)
            if nbytes == 0:
                raise ValueError(f"unsupported bm25 shape B={b} k={k}: " + _lib.last_error())
            buffers = TopkBuffers(b, k, nbytes, index.device)
        rc = lib.cmr_bm25_topk_ex(C.byref(st), q_terms.data_ptr(), q_ptr.data_ptr(), b, k,
                                  _ptr(row_mask), row_offset, buffers.scores.data_ptr(), buffers.ids.data_ptr(),
                                  buffers.counts.data_ptr(), buffers.flags.data_ptr(), buffers.ws.data_ptr(),
                                  buffers.ws.numel(), _stream(), _BM25_ALGO[algo])
    _lib.check(rc)
    return buffers.scores, buffers.ids, buffers.counts, buffers.flags


def gather_rows(emb: torch.Tensor, ids: torch.Tensor, *, row_offset: int = 0) -> torch.Tensor:
    """Rows ``ids - row_offset`` of emb (bf16 [N, D]); ids outside this shard
    (or -1) give zero rows.  ids: int64 [...]; returns bf16 [..., D]."""
    _require_cuda(emb, "emb")
    _require_cuda(ids, "ids")
    if ids.dtype != torch.int64:
        raise ValueError("ids must be int64")
    n_rows, dim = emb.shape
    out = torch.empty((*ids.shape, dim), dtype=torch.bfloat16, device=emb.device)
    with torch.cuda.device(emb.device):
        _lib.check(_lib.load().cmr_gather_rows(emb.data_ptr(), n_rows, dim, row_offset, ids.data_ptr(),
                                               ids.numel(), out.data_ptr(), _stream()))
    return out


def mmr_select(cand_rows: torch.Tensor, cand_sims: torch.Tensor, cand_ids: torch.Tensor,
               cand_counts: torch.Tensor, k: int, lambd: float = 0.5):
    """Greedy MMR over each query's pool.  cand_rows bf16 [B, pool, D], cand_sims
    f64 [B, pool], cand_ids i64 [B, pool], cand_counts i32 [B].  Returns
    (ids i64 [B,k], sims f64 [B,k], counts i32 [B])."""
    for name, t in (("cand_rows", cand_rows), ("cand_sims", cand_sims), ("cand_ids", cand_ids),
                    ("cand_counts", cand_counts)):
        _require_cuda(t, name)
    b, pool, dim = cand_rows.shape
    k = min(k, pool)
    dev = cand_rows.device
    out_ids = torch.empty((b, k), dtype=torch.int64, device=dev)
    out_sims = torch.empty((b, k), dtype=torch.float64, device=dev)
    out_counts = torch.empty((b,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().cmr_mmr_select(cand_rows.data_ptr(), cand_sims.data_ptr(), cand_ids.data_ptr(),
                                              cand_counts.data_ptr(), b, pool, dim, k, float(lambd),
                                              out_ids.data_ptr(), out_sims.data_ptr(), out_counts.data_ptr(),
                                              _stream()))
    return out_ids, out_sims, out_counts


def hybrid_fuse(vec, bm, *, top_k: int, rrf_k: int = 60, w_vec: float = 1.0, w_bm: float = 1.0):
    """RRF + merge + final order.  vec = (ids i64 [B,kv], sims f64 [B,kv], counts i32 [B]);
    bm = (ids, scores, counts) or None for the non-hybrid path.  Returns
    (ids [B,top_k], fused, vector_distance (NaN = None), bm25_score (NaN = None), counts)."""
    v_ids, v_sims, v_cnt = vec
    for name, t in (("vec ids", v_ids), ("vec sims", v_sims), ("vec counts", v_cnt)):
        _require_cuda(t, name)
    b, kv = v_ids.shape
    dev = v_ids.device
    if bm is not None:
        b_ids, b_sc, b_cnt = bm
        for name, t in (("bm ids", b_ids), ("bm scores", b_sc), ("bm counts", b_cnt)):
            _require_cuda(t, name)
        kb = b_ids.shape[1]
        bp = (b_ids.data_ptr(), b_sc.data_ptr(), b_cnt.data_ptr())
    else:
        kb, bp = 0, (None, None, None)
    # the five results are views of ONE allocation ([ids | fused | vector_distance | bm25 | counts]), so a caller
    # that wants them on the host can fetch them with a single copy (fused_result_buffer)
    n8 = b * top_k * 8
    buf = torch.empty((4 * n8 + b * 4,), dtype=torch.uint8, device=dev)
    out_ids = buf[:n8].view(torch.int64).view(b, top_k)
    out_fused = buf[n8:2 * n8].view(torch.float64).view(b, top_k)
    out_vd = buf[2 * n8:3 * n8].view(torch.float64).view(b, top_k)
    out_bm = buf[3 * n8:4 * n8].view(torch.float64).view(b, top_k)
    out_cnt = buf[4 * n8:].view(torch.int32)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().cmr_hybrid_fuse(v_ids.data_ptr(), v_sims.data_ptr(), v_cnt.data_ptr(), kv, *bp, kb, b,
                                               float(w_vec), float(w_bm), int(rrf_k), int(top_k),
                                               out_ids.data_ptr(), out_fused.data_ptr(), out_vd.data_ptr(),
                                               out_bm.data_ptr(), out_cnt.data_ptr(), _stream()))
    return out_ids, out_fused, out_vd, out_bm, out_cnt


def fused_result_buffer(out) -> Optional[torch.Tensor]:
    """The single uint8 allocation behind hybrid_fuse's five results (None if ``out`` is not such a tuple)."""
    ids, fused, vd, bm, cnt = out
    base = ids._base
    if base is None:
        return None
    while base._base is not None:
        base = base._base
    b, k = ids.shape
    ok = (base.dtype == torch.uint8 and base.numel() == 4 * b * k * 8 + b * 4 and ids.data_ptr() == base.data_ptr()
          and cnt.data_ptr() == base.data_ptr() + 4 * b * k * 8)
    return base if ok else None


def topk_merge(scores: torch.Tensor, ids: torch.Tensor, counts: torch.Tensor):
    """Merge per-shard lists: scores f64 [G,B,k], ids i64 [G,B,k], counts i32 [G,B]
    -> (scores [B,k], ids [B,k], counts [B]) ordered by (score desc, id asc)."""
    for name, t in (("scores", scores), ("ids", ids), ("counts", counts)):
        _require_cuda(t, name)
    g, b, k = scores.shape
    dev = scores.device
    out_s = torch.empty((b, k), dtype=torch.float64, device=dev)
    out_i = torch.empty((b, k), dtype=torch.int64, device=dev)
    out_c = torch.empty((b,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().cmr_topk_merge(scores.data_ptr(), ids.data_ptr(), counts.data_ptr(), g, b, k,
                                              out_s.data_ptr(), out_i.data_ptr(), out_c.data_ptr(), _stream()))
    return out_s, out_i, out_c


def rrf_fuse_lists(list_ids: torch.Tensor, list_counts: torch.Tensor, weights: torch.Tensor, rrf_k: int = 60):
    """rrf_fuse over L rank lists.  list_ids i64 [L, max_len], list_counts i32 [L],
    weights f64 [L] (device).  Returns (ids i64 [L*max_len], scores f64 [L*max_len],
    count i32 [1]): distinct ids in first-appearance order."""
    for name, t in (("list_ids", list_ids), ("list_counts", list_counts), ("weights", weights)):
        _require_cuda(t, name)
    if list_ids.dtype != torch.int64 or list_counts.dtype != torch.int32 or weights.dtype != torch.float64:
        raise ValueError("rrf_fuse_lists: list_ids int64, list_counts int32, weights float64")
    n_lists, max_len = list_ids.shape
    dev = list_ids.device
    out_ids = torch.empty((n_lists * max_len,), dtype=torch.int64, device=dev)
    out_scores = torch.empty((n_lists * max_len,), dtype=torch.float64, device=dev)
    out_count = torch.zeros((1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().cmr_rrf_fuse(list_ids.data_ptr(), list_counts.data_ptr(), n_lists, max_len,
                                            weights.data_ptr(), int(rrf_k), out_ids.data_ptr(),
                                            out_scores.data_ptr(), out_count.data_ptr(), _stream()))
    return out_ids, out_scores, out_count


def filter_mask(field_codes: torch.Tensor, clause_field: torch.Tensor, clause_code: torch.Tensor,
                alive: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """uint8 [n_rows] mask of the rows whose dictionary-coded metadata satisfy every
    (field, code) clause (cmr_filter_mask).  field_codes int32 [n_fields, n_rows]."""
    _require_cuda(field_codes, "field_codes")
    if field_codes.dtype != torch.int32 or field_codes.dim() != 2:
        raise ValueError("field_codes must be int32 [n_fields, n_rows]")
    n_fields, n_rows = field_codes.shape
    n_clauses = int(clause_field.numel())
    if n_clauses:
        _require_cuda(clause_field, "clause_field")
        _require_cuda(clause_code, "clause_code")
    if alive is not None:
        _require_cuda(alive, "alive")
    if out is None:
        out = torch.empty((n_rows,), dtype=torch.uint8, device=field_codes.device)
    with torch.cuda.device(field_codes.device):
        _lib.check(_lib.load().cmr_filter_mask(field_codes.data_ptr(), n_rows, n_fields,
                                               clause_field.data_ptr() if n_clauses else None,
                                               clause_code.data_ptr() if n_clauses else None, n_clauses,
                                               _ptr(alive), out.data_ptr(), _stream()))
    return out


def masked_df(term_ptr: torch.Tensor, post_doc: torch.Tensor, row_mask: torch.Tensor):
    """Document frequency of every term over the documents passing ``row_mask`` (uint8 [n_docs]) and the
    index of each term's first passing posting (cmr_masked_df).  Returns (df int32 [V], first int32 [V];
    first >= 0x7F7F7F7F where df == 0)."""
    for name, t in (("term_ptr", term_ptr), ("post_doc", post_doc), ("row_mask", row_mask)):
        _require_cuda(t, name)
    if term_ptr.dtype != torch.int64 or post_doc.dtype != torch.int32 or row_mask.dtype != torch.uint8:
        raise ValueError("term_ptr int64, post_doc int32, row_mask uint8 expected")
    n_terms = term_ptr.numel() - 1
    dev = post_doc.device
    df = torch.empty((n_terms,), dtype=torch.int32, device=dev)
    first = torch.empty((n_terms,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().cmr_masked_df(term_ptr.data_ptr(), post_doc.data_ptr(), n_terms, post_doc.numel(),
                                             row_mask.data_ptr(), df.data_ptr(), first.data_ptr(), _stream()))
    return df, first


def shard_msg_bytes(pool: int, kb: int, dim: int) -> int:
    n = _lib.load().cmr_shard_msg_bytes(pool, kb, dim)
    if n == 0:
        raise ValueError("bad shard message shape")
    return int(n)


def shard_pack(dense, bm, emb: Optional[torch.Tensor], *, row_offset: int = 0,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One rank's message of a sharded step (cmr_shard_pack).  dense = (scores f64 [B,pool],
    ids i64 [B,pool], counts i32 [B], flags i32 [B] or None); bm = (scores, ids, counts) or
    None; emb: this shard's matrix (its rows are copied for the dense candidates) or None
    when the step does not need rows.  Returns uint8 [B, msg_bytes]."""
    d_s, d_i, d_c, d_f = dense
    b, pool = d_s.shape
    kb = 0 if bm is None else bm[0].shape[1]
    dim = 0 if emb is None else emb.shape[1]
    nbytes = shard_msg_bytes(pool, kb, dim)
    if out is None:
        out = torch.empty((b, nbytes), dtype=torch.uint8, device=d_s.device)
    bp = (None, None, None) if bm is None else (bm[0].data_ptr(), bm[1].data_ptr(), bm[2].data_ptr())
    with torch.cuda.device(d_s.device):
        _lib.check(_lib.load().cmr_shard_pack(d_s.data_ptr(), d_i.data_ptr(), d_c.data_ptr(), _ptr(d_f), pool, *bp, kb,
                                              _ptr(emb), 0 if emb is None else emb.shape[0], dim, row_offset, b,
                                              out.data_ptr(), _stream()))
    return out


def shard_merge(gathered: torch.Tensor, pool: int, kb: int, dim: int):
    """Merge the all-gathered messages uint8 [G, B, msg_bytes] (cmr_shard_merge).  Returns
    (d_scores [B,pool], d_ids, d_counts, d_flags, d_rows bf16 [B,pool,dim] or None,
    b_scores [B,kb] or None, b_ids, b_counts)."""
    _require_cuda(gathered, "gathered")
    g, b, nbytes = gathered.shape
    if nbytes != shard_msg_bytes(pool, kb, dim):
        raise ValueError("gathered message size does not match (pool, kb, dim)")
    dev = gathered.device
    d_s = torch.empty((b, pool), dtype=torch.float64, device=dev)
    d_i = torch.empty((b, pool), dtype=torch.int64, device=dev)
    d_c = torch.empty((b,), dtype=torch.int32, device=dev)
    d_f = torch.empty((b,), dtype=torch.int32, device=dev)
    rows = torch.empty((b, pool, dim), dtype=torch.bfloat16, device=dev) if dim else None
    b_s = torch.empty((b, kb), dtype=torch.float64, device=dev) if kb else None
    b_i = torch.empty((b, kb), dtype=torch.int64, device=dev) if kb else None
    b_c = torch.empty((b,), dtype=torch.int32, device=dev) if kb else None
    with torch.cuda.device(dev):
        _lib.check(_lib.load().cmr_shard_merge(gathered.data_ptr(), g, b, pool, kb, dim, d_s.data_ptr(), d_i.data_ptr(),
                                               d_c.data_ptr(), d_f.data_ptr(), _ptr(rows), _ptr(b_s), _ptr(b_i),
                                               _ptr(b_c), _stream()))
    return d_s, d_i, d_c, d_f, rows, b_s, b_i, b_c


WIDE_K = _lib.CMR_MAX_K   # the widest over-selection one pass can certify (KP = 128)


def dense_topk_certified(emb: torch.Tensor, queries: torch.Tensor, k: int, *, row_mask: Optional[torch.Tensor] = None,
                         row_offset: int = 0, cert_eps: Optional[float] = None,
                         workspace: Optional[DenseWorkspace] = None, algo: str = "auto"):
    """dense_topk, then the queries whose result could not be certified are served by a ladder and
    patched in place: (1) the same fast kernels with the widest over-selection (top-120, KP = 128:
    certifies through clusters of ~100 exact duplicates around rank k -- duplicated course chunks --
    for the price of one more pass over the matrix); (2) what is still flagged goes to the exhaustive
    float64 scan, which ranks on the exact score of every row.  Synchronises (it reads the flags)."""
    scores, ids, counts, flags = dense_topk(emb, queries, k, row_mask=row_mask, row_offset=row_offset,
                                            cert_eps=cert_eps, workspace=workspace, algo=algo)
    bad = torch.nonzero(flags).flatten()
    if bad.numel():
        q = queries[None, :] if queries.dim() == 1 else queries
        if k < WIDE_K and algo != "exact":
            qb = q[bad].contiguous()
            s2, i2, c2, f2 = dense_topk(emb, qb, WIDE_K, row_mask=row_mask, row_offset=row_offset, cert_eps=cert_eps,
                                        algo=algo)
            ok = torch.nonzero(f2 == 0).flatten()
            if ok.numel():
                dst = bad[ok]
                scores[dst], ids[dst] = s2[ok, :k], i2[ok, :k]
                counts[dst], flags[dst] = torch.clamp(c2[ok], max=k), f2[ok]
            bad = bad[torch.nonzero(f2).flatten()]
    if bad.numel():
        s2, i2, c2, f2 = dense_topk(emb, q[bad].contiguous(), k, row_mask=row_mask, row_offset=row_offset,
                                    algo="exact")
        scores[bad], ids[bad], counts[bad], flags[bad] = s2, i2, c2, f2
    return scores, ids, counts, flags


class ShardP2PStruct(ctypes.Structure):
    """Mirror of ``cmr_shard_p2p`` (include/cmrag.h)."""
    _fields_ = [("peer_recv", ctypes.c_void_p), ("peer_flags", ctypes.c_void_p), ("state", ctypes.c_void_p),
                ("n_parts", ctypes.c_int32), ("my_rank", ctypes.c_int32), ("slot_stride", ctypes.c_uint64),
                ("parity_stride", ctypes.c_uint64), ("peer_rows", ctypes.c_void_p), ("peer_row_lo", ctypes.c_void_p)]


def shard_exchange_pack(dense, bm, emb: Optional[torch.Tensor], x: "ShardP2PStruct", *, row_offset: int = 0) -> None:
    """cmr_shard_exchange_pack: like shard_pack, but the message is stored directly into every
    rank's receive buffer over NVLink and the epoch flags are raised (no collective)."""
    import ctypes as C
    d_s, d_i, d_c, d_f = dense
    b, pool = d_s.shape
    kb = 0 if bm is None else bm[0].shape[1]
    dim = 0 if emb is None else emb.shape[1]
    bp = (None, None, None) if bm is None else (bm[0].data_ptr(), bm[1].data_ptr(), bm[2].data_ptr())
    with torch.cuda.device(d_s.device):
        _lib.check(_lib.load().cmr_shard_exchange_pack(d_s.data_ptr(), d_i.data_ptr(), d_c.data_ptr(), _ptr(d_f), pool,
                                                       *bp, kb, _ptr(emb), 0 if emb is None else emb.shape[0], dim,
                                                       row_offset, b, C.byref(x), _stream()))


def shard_exchange_merge(local_recv: torch.Tensor, local_flags: torch.Tensor, x: "ShardP2PStruct",
                         timeout_flag: torch.Tensor, n_queries: int, pool: int, kb: int, dim: int):
    """cmr_shard_exchange_merge: wait for every rank's flag of this epoch, then merge.  Same
    outputs as shard_merge."""
    import ctypes as C
    dev = local_recv.device
    b = n_queries
    d_s = torch.empty((b, pool), dtype=torch.float64, device=dev)
    d_i = torch.empty((b, pool), dtype=torch.int64, device=dev)
    d_c = torch.empty((b,), dtype=torch.int32, device=dev)
    d_f = torch.empty((b,), dtype=torch.int32, device=dev)
    rows = torch.empty((b, pool, dim), dtype=torch.bfloat16, device=dev) if dim else None
    b_s = torch.empty((b, kb), dtype=torch.float64, device=dev) if kb else None
    b_i = torch.empty((b, kb), dtype=torch.int64, device=dev) if kb else None
    b_c = torch.empty((b,), dtype=torch.int32, device=dev) if kb else None
    with torch.cuda.device(dev):
        _lib.check(_lib.load().cmr_shard_exchange_merge(local_recv.data_ptr(), local_flags.data_ptr(), C.byref(x),
                                                        timeout_flag.data_ptr(), b, pool, kb, dim, d_s.data_ptr(),
                                                        d_i.data_ptr(), d_c.data_ptr(), d_f.data_ptr(), _ptr(rows),
                                                        _ptr(b_s), _ptr(b_i), _ptr(b_c), _stream()))
    return d_s, d_i, d_c, d_f, rows, b_s, b_i, b_c
