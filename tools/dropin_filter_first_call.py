"""First call under each of several NEW filters through HybridRetriever.retrieve (1M rows): separates the
process's one-time warm-up from the per-filter cost.  python tools/dropin_filter_first_call.py [rows]"""
import os, sys, time
from types import SimpleNamespace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
a = SimpleNamespace(rows=10_000_000, dim=768)
captured = {}
orig = bench.HybridRetriever if hasattr(bench, "HybridRetriever") else None
# reuse bench.dropin_block's setup by monkeypatching its measuring loop: simplest is to copy its outputs
import classmate_rag_b200.retrieval.fusion as fusion
_real = fusion.HybridRetriever.retrieve
calls = []
def timed(self, **kw):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = _real(self, **kw)
    calls.append((dict(kw.get("filters") or {}), (time.perf_counter() - t0) * 1e3))
    captured["hr"], captured["kw"] = self, kw
    return r
fusion.HybridRetriever.retrieve = timed
out = bench.dropin_block(a, torch.device("cuda"), rows, iters=5)
print({k: out[k] for k in ("no_filter", "course_filter")})
hr, kw = captured["hr"], captured["kw"]
for c in ("C5", "C7", "C9"):
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hr.retrieve(question=kw["question"], filters={"course": c}, top_k=8)
        print(f"filter course={c} call {rep}: {(time.perf_counter() - t0) * 1e3:.2f} ms", flush=True)
