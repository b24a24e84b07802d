"""Deterministic synthetic corpora and queries.

Rows / documents are generated in fixed blocks of 65 536 with one seed per
block, so any row-sharding of the corpus sees exactly the same data: shard
results can be compared bit for bit with the unsharded run.  Generation runs
with torch on whatever device is given (the CUDA and CPU generators produce
different streams; a corpus is only ever compared with itself).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

BLOCK = 65536


def _gen(device, seed: int) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def dense_block(b: int, d: int, device, seed: int = 0xC0FFEE) -> torch.Tensor:
    """Block b of the corpus: unit-norm gaussian rows rounded to bf16, with exact
    duplicates planted (row i with i % 1000 == 999 copies row i - 500)."""
    x = torch.randn((BLOCK, d), generator=_gen(device, seed + b), device=device, dtype=torch.float32)
    x = torch.nn.functional.normalize(x, dim=1).to(torch.bfloat16)
    idx = torch.arange(BLOCK, device=device)
    dup = idx[(idx % 1000) == 999]
    x[dup] = x[dup - 500]
    return x


def dense_corpus(n: int, d: int, device, *, row_lo: int = 0, row_hi: int = None, seed: int = 0xC0FFEE) -> torch.Tensor:
    """Rows [row_lo, row_hi) of the n-row corpus as a bf16 [rows, d] tensor."""
    row_hi = n if row_hi is None else row_hi
    out = torch.empty((row_hi - row_lo, d), dtype=torch.bfloat16, device=device)
    b = row_lo // BLOCK
    while b * BLOCK < row_hi:
        lo, hi = max(row_lo, b * BLOCK), min(row_hi, (b + 1) * BLOCK, n)
        if hi > lo:
            blk = dense_block(b, d, device, seed)
            out[lo - row_lo: hi - row_lo] = blk[lo - b * BLOCK: hi - b * BLOCK]
        b += 1
    return out


def dense_queries(n: int, d: int, n_queries: int, device, *, seed: int = 0xBEEF,
                  corpus_seed: int = 0xC0FFEE) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp32 unit queries q = normalize(c_r + 0.5 * noise), |noise| ~ 1, so the
    planted row r scores ~0.89 and is the true top-1.  Returns (q [B,d] fp32 on
    `device`, planted rows int64 [B])."""
    g = _gen("cpu", seed)
    rows = torch.randint(0, n, (n_queries,), generator=g)
    noise = torch.randn((n_queries, d), generator=g) / d ** 0.5
    base = torch.empty((n_queries, d), dtype=torch.float32)
    cache = {}
    for i, r in enumerate(rows.tolist()):
        b = r // BLOCK
        if b not in cache:
            cache.clear()
            cache[b] = dense_block(b, d, device, corpus_seed)
        base[i] = cache[b][r - b * BLOCK].float().cpu()
    q = torch.nn.functional.normalize(base + 0.5 * noise, dim=1)
    return q.to(device), rows


def _zipf_cdf(vocab: int, s: float, device) -> torch.Tensor:
    p = 1.0 / torch.arange(1, vocab + 1, dtype=torch.float64, device=device) ** s
    return torch.cumsum(p / p.sum(), 0)


def lexical_block(b: int, vocab: int, mean_len: float, device, seed: int = 0x1E81CA1, s: float = 1.07):
    """Documents of block b: (lengths int64 [BLOCK], tokens int32 [sum])."""
    g = _gen(device, seed + b)
    ln = torch.exp(torch.randn(BLOCK, generator=g, device=device) * 0.5 + float(np.log(mean_len)))
    lens = ln.clamp(1, 4 * mean_len).long()
    total = int(lens.sum())
    u = torch.rand(total, generator=g, device=device, dtype=torch.float64)
    tokens = torch.searchsorted(_zipf_cdf(vocab, s, device), u).clamp_(max=vocab - 1).to(torch.int32)
    return lens, tokens


def lexical_corpus(n_docs: int, vocab: int, mean_len: float, device, *, doc_lo: int = 0, doc_hi: int = None,
                   seed: int = 0x1E81CA1):
    """Documents [doc_lo, doc_hi): (doc_ptr int64 [docs+1], tokens int32)."""
    doc_hi = n_docs if doc_hi is None else doc_hi
    lens_parts, tok_parts = [], []
    b = doc_lo // BLOCK
    while b * BLOCK < doc_hi:
        lo, hi = max(doc_lo, b * BLOCK), min(doc_hi, (b + 1) * BLOCK, n_docs)
        if hi > lo:
            lens, tokens = lexical_block(b, vocab, mean_len, device, seed)
            ptr = torch.zeros(BLOCK + 1, dtype=torch.int64, device=device)
            ptr[1:] = torch.cumsum(lens, 0)
            a, z = lo - b * BLOCK, hi - b * BLOCK
            lens_parts.append(lens[a:z])
            tok_parts.append(tokens[int(ptr[a]): int(ptr[z])])
        b += 1
    lens = torch.cat(lens_parts) if lens_parts else torch.zeros(0, dtype=torch.int64, device=device)
    doc_ptr = torch.zeros(lens.numel() + 1, dtype=torch.int64, device=device)
    doc_ptr[1:] = torch.cumsum(lens, 0)
    tokens = torch.cat(tok_parts) if tok_parts else torch.zeros(0, dtype=torch.int32, device=device)
    return doc_ptr, tokens


def lexical_queries(n_queries: int, vocab: int, *, n_tokens: int = 6, seed: int = 0xFACE, s: float = 1.07) -> List[List[int]]:
    """Term-id queries: Zipf draws, 5 % with a repeated token, 2 % with an
    out-of-vocabulary token (-1)."""
    rng = np.random.default_rng(seed)
    p = 1.0 / np.arange(1, vocab + 1) ** s
    p /= p.sum()
    draws = rng.choice(vocab, size=(n_queries, n_tokens), p=p)
    rep = rng.random(n_queries) < 0.05
    oov = rng.random(n_queries) < 0.02
    out = []
    for i in range(n_queries):
        q = draws[i].tolist()
        if rep[i]:
            q[-1] = q[0]
        if oov[i]:
            q[n_tokens // 2] = -1
        out.append(q)
    return out
