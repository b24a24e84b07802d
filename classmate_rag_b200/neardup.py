"""Near-duplicate cosine filter over a chunk-embedding matrix.

No reference symbol exists for it; the semantics follow the greedy keep-first rule of the
reference's text dedup (rag/utils/dedup.py:40-55, comparison ``>=`` at :50) applied to
cosine similarity, at the place `rag rebuild` re-embeds the corpus
(rag/admin/backup.py:226-233): row i is kept iff no previously KEPT row j < i has
q.c >= threshold.

The N x N similarity work runs on the tcgen05 tensor cores (``cmr_neardup_edges``), the
borderline pairs are rescored exactly in float64 (``cmr_neardup_rescore``), so the decision
is bit-identical to the oracle (oracle/np_oracle.py:neardup_keep_mask).  Sharded over G
ranks the 128-row blocks of the lower triangle are dealt round-robin (rank r takes blocks
r, r+G, ...: the triangle is balanced), the few surviving edges are all-gathered and every
rank resolves the same keep mask.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def candidate_bound(threshold: float, dim: int, max_row_norm: float = 1.0) -> float:
    """fp32 admission bound of the tensor-core pass: threshold minus the worst-case error of
    an fp32-accumulated bf16 dot (dim * 2^-22 * |a| * |b|), rounded down to float32."""
    b = np.float32(threshold - dim * 2.0 ** -22 * 1.01 * max_row_norm * max_row_norm)
    return float(np.nextafter(b, np.float32(-np.inf)))


def neardup_edges(emb: torch.Tensor, threshold: float = 0.95, *, rank: int = 0, world: int = 1,
                  max_row_norm: float = 1.0) -> torch.Tensor:
    """Exact edges (i << 32 | j, j < i, q.c >= threshold) of this rank's share of the lower
    triangle, unordered, as an int64 device tensor."""
    if not emb.is_cuda or emb.dtype != torch.bfloat16 or not emb.is_contiguous():
        raise RuntimeError("emb must be a contiguous CUDA bfloat16 matrix (no CPU path)")
    n, d = emb.shape
    lib = _lib.load()
    dev = emb.device
    if n < 2:
        return torch.zeros((0,), dtype=torch.int64, device=dev)
    bound = candidate_bound(threshold, d, max_row_norm)
    cap = max(1 << 16, 4 * n // max(world, 1))
    count = torch.zeros((1,), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        while True:
            edges = torch.empty((cap,), dtype=torch.int64, device=dev)
            _lib.check(lib.cmr_neardup_edges(emb.data_ptr(), n, d, bound, rank, world, edges.data_ptr(), cap,
                                             count.data_ptr(), _stream()))
            found = int(count.item())
            if found <= cap:
                break
            cap = found + found // 8 + 1024   # the buffer overflowed: exact size is now known
        exact = torch.empty((max(found, 1),), dtype=torch.int64, device=dev)
        _lib.check(lib.cmr_neardup_rescore(emb.data_ptr(), d, edges.data_ptr(), found, float(threshold),
                                           exact.data_ptr(), count.data_ptr(), _stream()))
        kept = int(count.item())
    return exact[:kept]


def resolve(edges: torch.Tensor, n_rows: int) -> torch.Tensor:
    """uint8 [n_rows] keep mask from ALL exact edges (any order): greedy, ascending rows."""
    dev = edges.device
    keep = torch.empty((n_rows,), dtype=torch.uint8, device=dev)
    srt = torch.sort(edges).values.contiguous() if edges.numel() else edges
    with torch.cuda.device(dev):
        _lib.check(_lib.load().cmr_neardup_resolve(srt.data_ptr() if srt.numel() else None, srt.numel(), n_rows,
                                                   keep.data_ptr(), _stream()))
    return keep


def neardup_keep_mask(emb: torch.Tensor, threshold: float = 0.95, *, max_row_norm: float = 1.0,
                      group=None) -> torch.Tensor:
    """keep[i] (uint8, device).  With a torch.distributed ``group`` (or the default group
    when initialised and ``group`` is True) every rank holds the whole matrix and computes a
    1/G share of the triangle; edges are exchanged with one all-gather."""
    import torch.distributed as dist
    world, rank = 1, 0
    pg = None
    if group is not None and dist.is_available() and dist.is_initialized():
        pg = None if group is True else group
        world, rank = dist.get_world_size(pg), dist.get_rank(pg)
    mine = neardup_edges(emb, threshold, rank=rank, world=world, max_row_norm=max_row_norm)
    if world > 1:
        mine = allgather_edges(mine, pg)
    return resolve(mine, emb.shape[0])


def allgather_edges(mine: torch.Tensor, pg=None) -> torch.Tensor:
    """Concatenation of every rank's edge list (ragged: counts first, then padded payloads)."""
    import torch.distributed as dist
    world = dist.get_world_size(pg)
    counts = torch.zeros((world,), dtype=torch.int64, device=mine.device)
    counts[dist.get_rank(pg)] = mine.numel()
    dist.all_reduce(counts, group=pg)
    width = int(counts.max().item())
    if width == 0:
        return mine
    padded = torch.full((width,), -1, dtype=torch.int64, device=mine.device)
    padded[: mine.numel()] = mine
    gathered = torch.empty((world, width), dtype=torch.int64, device=mine.device)
    dist.all_gather_into_tensor(gathered, padded, group=pg)
    return torch.cat([gathered[r, : int(counts[r])] for r in range(world)])
