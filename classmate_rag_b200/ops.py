"""Tensor-level calls into libcmrag.so.  torch is plumbing here: device
memory, streams.  All compute is in the CUDA library."""
from __future__ import annotations

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


def dense_topk(emb: torch.Tensor, queries: torch.Tensor, k: int, *, row_mask: Optional[torch.Tensor] = None,
               row_offset: int = 0, cert_eps: Optional[float] = None,
               workspace: Optional[DenseWorkspace] = None):
    """Exact top-k of queries (bf16 [B, D]) against emb (bf16 [N, D]).

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
    if cert_eps is None:
        cert_eps = dim * 2.0 ** -24 * 1.02
    if workspace is None or workspace.key != (n_rows, dim, b, k):
        workspace = DenseWorkspace(n_rows, dim, b, k, emb.device)
    lib = _lib.load()
    with torch.cuda.device(emb.device):
        rc = lib.cmr_dense_topk(emb.data_ptr(), n_rows, dim, queries.data_ptr(), b, k, _ptr(row_mask),
                                row_offset, float(cert_eps), workspace.scores.data_ptr(),
                                workspace.ids.data_ptr(), workspace.counts.data_ptr(),
                                workspace.flags.data_ptr(), workspace.ws.data_ptr(), workspace.ws.numel(),
                                _stream())
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


def bm25_topk(index, q_terms: torch.Tensor, q_ptr: torch.Tensor, k: int, *,
              row_mask: Optional[torch.Tensor] = None, row_offset: int = 0,
              buffers: Optional[TopkBuffers] = None):
    """Exact BM25 top-k for a batch of tokenised queries (device tensors:
    q_terms int32, q_ptr int32 [B+1]).  Returns (scores f64 [B,k], ids i64 [B,k],
    counts, flags) on the device, enqueued on the current stream."""
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
        rc = lib.cmr_bm25_topk(C.byref(st), q_terms.data_ptr(), q_ptr.data_ptr(), b, k,
                               _ptr(row_mask), row_offset, buffers.scores.data_ptr(), buffers.ids.data_ptr(),
                               buffers.counts.data_ptr(), buffers.flags.data_ptr(), buffers.ws.data_ptr(),
                               buffers.ws.numel(), _stream())
    _lib.check(rc)
    return buffers.scores, buffers.ids, buffers.counts, buffers.flags
