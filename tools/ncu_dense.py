"""Minimal dense workload for ncu (1M x 768, B=1, k=10): 3 warm-up + 3 launches."""
import sys
import torch
from classmate_rag_b200 import ops
n, d, k = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 768, 10
g = torch.Generator(device="cuda").manual_seed(1)
emb = torch.nn.functional.normalize(torch.randn((n, d), generator=g, device="cuda"), dim=1).to(torch.bfloat16)
q = emb[777:778].clone()
ws = ops.DenseWorkspace(n, d, 1, k, emb.device)
for _ in range(6):
    ops.dense_topk(emb, q, k, workspace=ws)
torch.cuda.synchronize()
print("ok", ws.ids[0, :3].tolist())
