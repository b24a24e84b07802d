"""One hybrid step on a single shard for ncu launch lists: python tools/ncu_step.py rows batch."""
import sys
import torch
from classmate_rag_b200 import lexical, ops, synth
from classmate_rag_b200.engine import HybridEngine, SearchParams
n, b = int(sys.argv[1]), int(sys.argv[2])
emb = synth.dense_corpus(n, 768, "cuda")
doc_ptr, tokens = synth.lexical_corpus(n, 30000, 64, "cuda")
lex = lexical.build_lexical_index(doc_ptr, tokens, 30000)
eng = HybridEngine(emb, lex)
q, _ = synth.dense_queries(n, 768, b, "cuda")
qb = ops.f32_to_bf16(q)
qt, qp = lexical.pack_queries(synth.lexical_queries(b, 30000))
qt, qp = qt.cuda(), qp.cuda()
p = SearchParams(top_k=10)
for _ in range(4):
    out = eng.search(qb, qt, qp, p)
torch.cuda.synchronize()
print("ok", int(out[4].sum()))
