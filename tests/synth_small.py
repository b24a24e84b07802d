"""Small deterministic synthetic corpora for the tests (NumPy, CPU)."""
from __future__ import annotations

import numpy as np
import torch


def zipf_corpus(seed: int, n_docs: int, vocab: int, mean_len: float, s: float = 1.07, empty_every: int = 97):
    """Token-id documents with Zipf term popularity and ragged lengths
    (including empty documents).  Returns (docs list, doc_ptr, tokens, vocab)."""
    rng = np.random.default_rng(seed)
    p = 1.0 / np.arange(1, vocab + 1) ** s
    p /= p.sum()
    lens = np.clip(rng.lognormal(np.log(mean_len), 0.5, n_docs).astype(int), 1, 4 * int(mean_len))
    if empty_every:
        lens[empty_every - 1::empty_every] = 0
    docs = [rng.choice(vocab, size=int(l), p=p).tolist() for l in lens]
    doc_ptr = np.zeros(n_docs + 1, dtype=np.int64)
    doc_ptr[1:] = np.cumsum(lens)
    tokens = np.fromiter((t for d in docs for t in d), dtype=np.int32, count=int(doc_ptr[-1]))
    return docs, torch.from_numpy(doc_ptr), torch.from_numpy(tokens), vocab


def zipf_queries(seed: int, n_queries: int, vocab: int, n_tokens: int = 6, s: float = 1.07):
    """Queries of term ids: Zipf draws, 1 in 5 with a repeated token, 1 in 7
    with an unknown token (-1), plus one empty query."""
    rng = np.random.default_rng(seed)
    p = 1.0 / np.arange(1, vocab + 1) ** s
    p /= p.sum()
    out = []
    for i in range(n_queries):
        q = rng.choice(vocab, size=n_tokens, p=p).tolist()
        if i % 5 == 1:
            q[-1] = q[0]
        if i % 7 == 2:
            q[1] = -1
        out.append(q)
    if n_queries > 3:
        out[3] = []
    return out
