"""BM25-only workload for ncu / quick timing: N docs, B queries, k, iterations, algo.
   python tools/ncu_bm25.py 10000000 32 10 5 head [check]
`check`: also run the exact kernel and compare the bytes; prints flagged queries of the head path."""
import json
import os
import sys

import torch

from classmate_rag_b200 import lexical, ops, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
b = int(sys.argv[2]) if len(sys.argv) > 2 else 32
k = int(sys.argv[3]) if len(sys.argv) > 3 else 8
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
algo = sys.argv[5] if len(sys.argv) > 5 else "auto"
check = len(sys.argv) > 6 and sys.argv[6] == "check"
doc_ptr, tokens = synth.lexical_corpus(n, 30000, 64, "cuda")
lex = lexical.build_lexical_index(doc_ptr, tokens, 30000, tile_docs=int(os.environ.get("TILE", str(lexical.DEFAULT_TILE_DOCS))))
del doc_ptr, tokens
torch.cuda.empty_cache()
n_sets = 4   # fresh queries every iteration (cold skip rows / posting slices, as inside a real step)
terms_all = synth.lexical_queries(b * n_sets, 30000)
sets = []
for s in range(n_sets):
    qt, qp = lexical.pack_queries(terms_all[s * b:(s + 1) * b])
    sets.append((qt.cuda(), qp.cuda()))
out = {"n": n, "b": b, "k": k, "algo": algo, "n_tiles": lex.n_tiles, "n_head": 0 if lex.head_terms is None else len(lex.head_terms)}
if check:
    bad = 0
    for qt, qp in sets:
        want = [t.clone() for t in ops.bm25_topk(lex, qt, qp, k, algo="exact")]
        got = [t.clone() for t in ops.bm25_topk(lex, qt, qp, k, algo=algo)]
        torch.cuda.synchronize()
        bad += sum(int(a.cpu().numpy().tobytes() != w.cpu().numpy().tobytes()) for a, w in zip(got, want))
    out["mismatching_tensors_vs_exact"] = bad
    if algo.startswith("head"):
        fl = ops.bm25_topk(lex, *sets[0], k, algo="head_nofallback")[3]
        torch.cuda.synchronize()
        out["flags_nofallback"] = [int(x) for x in fl.cpu().tolist() if x]
buf = None
for i in range(3):
    ops.bm25_topk(lex, *sets[i % n_sets], k, algo=algo)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
ev[0].record()
for i in range(iters):
    ops.bm25_topk(lex, *sets[i % n_sets], k, algo=algo)
    ev[i + 1].record()
torch.cuda.synchronize()
ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
byts = sum(lex.posting_bytes(t) for t in terms_all[:b])
out.update({"ms_med": ts[len(ts) // 2], "ms_min": ts[0], "us_per_query": ts[len(ts) // 2] * 1e3 / b,
            "postings_per_query": sum(int(lex.shard_df_host[t]) for q in terms_all[:b] for t in q if t >= 0) / b})
print(json.dumps(out))
