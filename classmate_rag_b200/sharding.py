"""Row-sharding of the corpus over the GPUs of one box (one process per GPU).

The dense matrix and the inverted index are partitioned by contiguous row
ranges; every rank scores the same queries against its shard and only the
per-shard top-k candidates travel.  The hot path (engine.HybridEngine.search) packs a
rank's dense pool (scores, ids and the embedding rows MMR needs) and its BM25
list into ONE message (cmr_shard_pack), all-gathers the messages (the only
collective of a step) and merges them with cmr_shard_merge under the total order
(score desc, id asc), so the result is identical for any number of shards.  The
per-retriever exchange (merge_topk: one all-gather of B x k x (score, id);
sum_rows: one all-reduce of the pool rows) remains for callers that run a single
retriever.  Index build needs one all-reduce of df[V] / token totals / first
positions so that every shard scores with the corpus-wide idf and avgdl.
The reference has no distributed code; this is new.
"""
from __future__ import annotations

import os
from typing import Callable, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .lexical import GlobalStats


def shard_range(n: int, rank: int, world: int, align: int = 16) -> Tuple[int, int]:
    """Contiguous, `align`-row aligned split of n rows."""
    per = (n + world - 1) // world
    per = (per + align - 1) // align * align
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


# bf16 row matrices allocated by shared_rows(): data_ptr -> (symmetric-memory handle, base tensor)
_SHARED_ROWS: dict = {}


def shared_rows(n_rows: int, dim: int, device, group=None) -> torch.Tensor:
    """A bf16 [n_rows, dim] matrix every rank of the box can read (torch symmetric memory: one
    cudaMalloc per rank, handles exchanged once).  With every shard's matrix allocated this way
    the peer exchange carries no embedding rows: the merge reads the merged pool's rows from
    their owners over NVLink (``cmr_shard_p2p.peer_rows``).  Collective: every rank calls it;
    ``n_rows`` may differ per rank (the allocation is the maximum).  Raises when symmetric
    memory is unavailable -- callers fall back to an ordinary tensor, whose rows then travel
    inside the messages."""
    import torch.distributed._symmetric_memory as symm_mem
    group = group if group is not None else dist.group.WORLD
    most = torch.tensor([int(n_rows)], dtype=torch.int64, device=device)
    dist.all_reduce(most, op=dist.ReduceOp.MAX, group=group)
    with torch.cuda.device(device):
        base = symm_mem.empty((max(int(most.item()), 1), dim), dtype=torch.bfloat16, device=device)
        handle = symm_mem.rendezvous(base, group)
    _SHARED_ROWS[base.data_ptr()] = (handle, base)
    return base[:n_rows]


def _shared_row_pointers(emb: Optional[torch.Tensor]):
    ent = None if emb is None else _SHARED_ROWS.get(emb.data_ptr())
    return None if ent is None else [int(p) for p in ent[0].buffer_ptrs]


class PeerExchange:
    """Receive buffers + flag words of the peer-memory exchange (cmr_shard_exchange_*), mapped
    into every rank of the box through torch's symmetric memory (cudaMalloc'ed buffers whose
    handles are exchanged once; afterwards the kernels store into peers over NVLink)."""

    def __init__(self, group, device, slot_bytes: int):
        import torch.distributed._symmetric_memory as symm_mem
        from . import ops
        self.group = group if group is not None else dist.group.WORLD
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.slot = (int(slot_bytes) + 255) // 256 * 256
        self.parity = self.slot * world
        with torch.cuda.device(device):
            self.recv = symm_mem.empty(2 * self.parity, dtype=torch.uint8, device=device)
            self.flags = symm_mem.empty(max(2 * world, 64), dtype=torch.int32, device=device)
            self.recv.zero_()
            self.flags.zero_()
            h_recv = symm_mem.rendezvous(self.recv, self.group)
            h_flags = symm_mem.rendezvous(self.flags, self.group)
        self._handles = (h_recv, h_flags)
        self.peer_recv = torch.tensor([int(p) for p in h_recv.buffer_ptrs], dtype=torch.int64, device=device)
        self.peer_flags = torch.tensor([int(p) for p in h_flags.buffer_ptrs], dtype=torch.int64, device=device)
        self.state = torch.zeros(2, dtype=torch.int32, device=device)
        self.timeout = torch.zeros(1, dtype=torch.int32, device=device)
        self.struct = ops.ShardP2PStruct(self.peer_recv.data_ptr(), self.peer_flags.data_ptr(), self.state.data_ptr(),
                                         world, rank, self.slot, self.parity)
        self.pull_rows = False
        torch.cuda.synchronize(device)
        dist.barrier(group=self.group)     # every rank's buffers are zeroed before anyone stores into them

    def attach_rows(self, emb: Optional[torch.Tensor], row_offset: int) -> bool:
        """Collective.  If EVERY rank's matrix came from shared_rows(), the exchange switches to its
        pull form (no rows in the messages); otherwise nothing changes."""
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        ptrs = _shared_row_pointers(emb)
        dev = self.state.device
        info = torch.zeros((world, 2), dtype=torch.int64, device=dev)
        info[rank, 0] = 1 if ptrs is not None and len(ptrs) == world else 0
        info[rank, 1] = int(row_offset)
        dist.all_reduce(info, group=self.group)
        if int(info[:, 0].min()) == 0:
            return False
        from . import ops
        self.peer_rows = torch.tensor(ptrs, dtype=torch.int64, device=dev)
        self.peer_row_lo = info[:, 1].contiguous()
        s = self.struct
        self.struct_pull = ops.ShardP2PStruct(s.peer_recv, s.peer_flags, s.state, s.n_parts, s.my_rank, s.slot_stride,
                                              s.parity_stride, self.peer_rows.data_ptr(), self.peer_row_lo.data_ptr())
        self._pull_ptr = emb.data_ptr()
        self.pull_rows = True
        return True

    def struct_for(self, emb: Optional[torch.Tensor]):
        """The descriptor a step over ``emb`` uses: the pull form only for the matrix it was set up with."""
        if self.pull_rows and emb is not None and emb.data_ptr() == self._pull_ptr:
            return self.struct_pull
        return self.struct


class ShardComm:
    """The exchanges of the sharded hot path."""

    def __init__(self, group=None, merge_fn: Optional[Callable] = None, peer_memory: Optional[bool] = None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._merge_fn = merge_fn  # injectable for the CPU (gloo) tests
        # peer-memory exchange: on by default with NCCL (CMRAG_P2P=0 turns it off); any failure
        # to set it up falls back to the NCCL all-gather of the same messages
        if peer_memory is None:
            peer_memory = os.environ.get("CMRAG_P2P", "1") != "0" and dist.get_backend(group) == "nccl"
        self.peer_memory = bool(peer_memory)
        self.peers: dict = {}      # buffer-set slot -> PeerExchange (two steps may be in flight: PipelinedSearch)
        self.peer_error: Optional[str] = None
        self.peer_generation = 0
        self.min_slot_bytes = 16 << 20

    @property
    def peer(self) -> Optional[PeerExchange]:
        return self.peers.get(0) if self.peers else None

    def peer_exchange(self, device, slot_bytes: int, rows: Optional[torch.Tensor] = None,
                      row_offset: int = 0, slot: int = 0) -> Optional[PeerExchange]:
        """The PeerExchange big enough for ``slot_bytes`` per rank (collective on first use /
        growth: every rank calls it with the same size), or None when unavailable.  ``rows``:
        this rank's matrix; if every rank's is a shared_rows() allocation the exchange pulls the
        pool rows from their owners instead of shipping them."""
        if not self.peer_memory:
            return None
        cur = self.peers.get(slot)
        if cur is None or cur.slot < slot_bytes:
            try:
                # generous first allocation: a later, larger message would need new buffers, and
                # CUDA graphs captured before hold the old addresses (see GraphedSearch.launch)
                cur = self.peers[slot] = PeerExchange(self.group, device, max(int(slot_bytes), self.min_slot_bytes))
                self.peer_generation += 1
                if os.environ.get("CMRAG_PULL_ROWS", "1") != "0":
                    cur.attach_rows(rows, row_offset)
            except Exception as exc:  # no symmetric memory on this build / topology
                self.peer_memory, self.peer_error = False, repr(exc)
                self.peers.clear()
                import warnings
                warnings.warn(f"peer-memory shard exchange unavailable ({exc!r}): using the NCCL all-gather "
                              "(set CMRAG_P2P=0 to silence)", RuntimeWarning, stacklevel=2)
        return self.peers.get(slot)

    def _merge(self, scores, ids, counts):
        if self._merge_fn is not None:
            return self._merge_fn(scores, ids, counts)
        from . import ops
        return ops.topk_merge(scores, ids, counts)

    def merge_topk(self, scores: torch.Tensor, ids: torch.Tensor, counts: torch.Tensor, flags: torch.Tensor):
        """All-gather the per-shard lists (one collective) and merge them."""
        b, k = scores.shape
        # one packed float64 message per rank: [scores | ids (bit-cast) | count | flag]
        msg = torch.empty((b, 2 * k + 2), dtype=torch.float64, device=scores.device)
        msg[:, :k] = scores
        msg[:, k:2 * k] = ids.view(torch.float64)
        msg[:, 2 * k] = counts.to(torch.float64)
        msg[:, 2 * k + 1] = flags.to(torch.float64)
        flat = torch.empty((self.world * b, 2 * k + 2), dtype=torch.float64, device=scores.device)
        dist.all_gather_into_tensor(flat, msg, group=self.group)
        out = flat.view(self.world, b, 2 * k + 2)
        g_scores = out[:, :, :k].contiguous()
        g_ids = out[:, :, k:2 * k].contiguous().view(torch.int64)
        g_counts = out[:, :, 2 * k].to(torch.int32).contiguous()
        m_scores, m_ids, m_counts = self._merge(g_scores, g_ids, g_counts)
        m_flags = out[:, :, 2 * k + 1].amax(dim=0).to(torch.int32)
        return m_scores, m_ids, m_counts, m_flags

    def all_gather_bytes(self, msg: torch.Tensor) -> torch.Tensor:
        """The single collective of a sharded step: every rank's packed message
        (uint8 [B, msg_bytes], see cmr_shard_pack) -> uint8 [G, B, msg_bytes]."""
        out = torch.empty((self.world, *msg.shape), dtype=msg.dtype, device=msg.device)
        dist.all_gather_into_tensor(out.view(self.world * msg.shape[0], *msg.shape[1:]), msg.contiguous(),
                                    group=self.group)
        return out

    def sum_rows(self, rows: torch.Tensor) -> torch.Tensor:
        """Reassemble gathered bf16 rows: exactly one rank holds each row, the
        others hold zeros, so an integer SUM is exact."""
        buf = rows.contiguous().view(torch.int32)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        return buf.view(torch.bfloat16).view(rows.shape)


def global_corpus_stats(doc_ptr: torch.Tensor, tokens: torch.Tensor, n_terms: int, *, doc_lo: int,
                        n_docs_total: int, token_offset: int, group=None) -> GlobalStats:
    """Corpus-wide (N, total tokens, df, first-appearance order) from shard-local
    token arrays: three all-reduces.  `token_offset` = number of tokens in all
    earlier shards (so first positions are comparable across shards)."""
    dev = tokens.device
    tok = tokens.long()
    n_local = doc_ptr.numel() - 1
    total_local = int(tok.numel())
    big = torch.iinfo(torch.int64).max
    first_pos = torch.full((n_terms,), big, dtype=torch.int64, device=dev)
    df = torch.zeros(n_terms, dtype=torch.int64, device=dev)
    if total_local:
        doc_len = (doc_ptr[1:] - doc_ptr[:-1]).long()
        doc_of = torch.repeat_interleave(torch.arange(n_local, device=dev), doc_len)
        key = torch.unique(tok * max(n_local, 1) + doc_of)
        df = torch.bincount(key // max(n_local, 1), minlength=n_terms)
        first_pos.scatter_reduce_(0, tok, torch.arange(total_local, device=dev) + token_offset, reduce="amin")
    total = torch.tensor([total_local], dtype=torch.int64, device=dev)
    dist.all_reduce(df, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(first_pos, op=dist.ReduceOp.MIN, group=group)
    order = torch.argsort(first_pos, stable=True)
    n_seen = int((df > 0).sum())
    return GlobalStats(n_docs_total, int(total.item()), df.cpu().numpy().astype(np.int64), order[:n_seen].cpu().numpy())
