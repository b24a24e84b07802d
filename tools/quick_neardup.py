"""Timing of the near-duplicate pair scan (development aid)."""
import json
import sys
import time

import torch

from classmate_rag_b200 import neardup

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 768
g = torch.Generator(device="cuda").manual_seed(1)
emb = torch.empty((n, d), dtype=torch.bfloat16, device="cuda")
step = 1 << 18
for lo in range(0, n, step):
    x = torch.randn((min(step, n - lo), d), generator=g, device="cuda")
    emb[lo:lo + x.shape[0]] = torch.nn.functional.normalize(x, dim=1).to(torch.bfloat16)
emb[1000:2000] = emb[0:1000]
for it in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    edges = neardup.neardup_edges(emb, 0.95)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    keep = neardup.resolve(edges, n)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(json.dumps({"n": n, "d": d, "edges": int(edges.numel()), "dropped": int(n - keep.sum()),
                      "scan_s": round(t1 - t0, 4), "resolve_s": round(t2 - t1, 4),
                      "TFLOPs": round(n * n * d / (t1 - t0) / 1e12, 1)}), flush=True)
