"""One batched dense call for ncu launch lists: python tools/ncu_dense_b.py n d b."""
import sys
import torch
from classmate_rag_b200 import ops, synth
n, d, b = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
emb = synth.dense_corpus(n, d, "cuda")
q, _ = synth.dense_queries(n, d, b, "cuda")
qb = ops.f32_to_bf16(q)
ws = ops.DenseWorkspace(n, d, b, 10, emb.device)
for _ in range(3):
    out = ops.dense_topk(emb, qb, 10, workspace=ws)
torch.cuda.synchronize()
print("ok", int(out[3].sum()))
