"""Determinism / race check of every kernel family: the same call repeated must give the same
bytes every time (and the batched dense path the exhaustive scan's answer)."""
import sys
import numpy as np
import torch
from classmate_rag_b200 import lexical, neardup, ops, synth
from classmate_rag_b200.engine import HybridEngine, SearchParams

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 25
bad = 0


def same(name, fn):
    global bad
    ref = None
    for r in range(reps):
        out = fn()
        torch.cuda.synchronize()
        b = [t.cpu().numpy().tobytes() for t in out]
        if ref is None:
            ref = b
        elif b != ref:
            bad += 1
            print("NONDETERMINISTIC", name, "rep", r, flush=True)
            return
    print("ok", name, flush=True)


n, d, vocab = 200_000, 768, 5000
emb = synth.dense_corpus(n, d, "cuda")
doc_ptr, tokens = synth.lexical_corpus(n, vocab, 48, "cuda")
lex = lexical.build_lexical_index(doc_ptr, tokens, vocab)
eng = HybridEngine(emb, lex)
for b in (1, 4, 8, 12, 32, 150):
    q, _ = synth.dense_queries(n, d, b, "cuda")
    qb = ops.f32_to_bf16(q)
    terms = synth.lexical_queries(b, vocab)
    qt, qp = [t.cuda() for t in lexical.pack_queries(terms)]
    for algo in ("scan", "mma", "exact"):
        if algo == "scan" and b > 32:
            continue
        same(f"dense {algo} b={b}", lambda: [t.clone() for t in ops.dense_topk(emb, qb, 24, algo=algo)])
    same(f"bm25 b={b} k=8", lambda: [t.clone() for t in ops.bm25_topk(lex, qt, qp, 8)])
    same(f"bm25 b={b} k=100", lambda: [t.clone() for t in ops.bm25_topk(lex, qt, qp, 100)])
    for ov in (False, True):
        eng.overlap = ov
        same(f"hybrid b={b} overlap={ov}", lambda: [t.clone() for t in eng.search(qb, qt, qp, SearchParams(top_k=10))])
nd = synth.dense_corpus(30_000, 256, "cuda")
nd[1::7] = nd[0::7][: nd[1::7].shape[0]]
same("neardup", lambda: [neardup.neardup_keep_mask(nd, 0.95).clone()])
print("done,", bad, "bad")
