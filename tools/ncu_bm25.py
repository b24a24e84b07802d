"""BM25-only workload for ncu / quick timing: N docs, B queries, k."""
import sys, json, time
import torch
from classmate_rag_b200 import lexical, ops, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
b = int(sys.argv[2]) if len(sys.argv) > 2 else 32
k = int(sys.argv[3]) if len(sys.argv) > 3 else 8
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
doc_ptr, tokens = synth.lexical_corpus(n, 30000, 64, "cuda")
import os
lex = lexical.build_lexical_index(doc_ptr, tokens, 30000, tile_docs=int(os.environ.get("TILE", str(lexical.DEFAULT_TILE_DOCS))))
del doc_ptr, tokens
terms = synth.lexical_queries(b, 30000)
qt, qp = lexical.pack_queries(terms)
qt, qp = qt.cuda(), qp.cuda()
for _ in range(3):
    out = ops.bm25_topk(lex, qt, qp, k)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
ev[0].record()
for i in range(iters):
    ops.bm25_topk(lex, qt, qp, k)
    ev[i + 1].record()
torch.cuda.synchronize()
ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
byts = sum(lex.posting_bytes(t) for t in terms)
print(json.dumps({"n": n, "b": b, "k": k, "ms_med": ts[len(ts) // 2], "us_per_query": ts[len(ts) // 2] * 1e3 / b,
                  "postings_per_query": byts / 4 / b, "GBps": byts / (ts[len(ts) // 2] * 1e-3) / 1e9, "n_tiles": lex.n_tiles}))
