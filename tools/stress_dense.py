"""Repeat the batched (tcgen05) dense top-k on shapes with several query blocks and compare it
with the exhaustive float64 scan on the same inputs: any run-to-run difference is a race."""
import sys
import numpy as np
import torch
from classmate_rag_b200 import ops

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
torch.manual_seed(0)
bad = 0
shapes = [(9000, 256, 300, 10), (9000, 256, 130, 10), (40000, 128, 300, 10), (9000, 768, 520, 24)]
for it in range(iters):
    n, d, b, k = shapes[it % len(shapes)]
    g = torch.Generator(device="cuda").manual_seed(it)
    x = torch.randn((n, d), device="cuda", generator=g)
    x = torch.nn.functional.normalize(x, dim=1).to(torch.bfloat16)
    x[63::64] = x[31::64][: x[63::64].shape[0]]
    rows = torch.randint(0, n, (b,), device="cuda", generator=g)
    q = x[rows].float() + 0.5 * torch.randn((b, d), device="cuda", generator=g) / d ** 0.5
    q = torch.nn.functional.normalize(q, dim=1).to(torch.bfloat16)
    ref = [t.clone() for t in ops.dense_topk(x, q, k, algo="exact")]
    for rep in range(6):
        out = [t.clone() for t in ops.dense_topk(x, q, k, algo="mma")]
        torch.cuda.synchronize()
        flagged = out[3] != 0
        same = torch.equal(out[1][~flagged], ref[1][~flagged]) and torch.equal(out[0][~flagged], ref[0][~flagged])
        if not same or bool(flagged.any()):
            bad += 1
            diff = ((out[1] != ref[1]).any(dim=1) & ~flagged).nonzero().flatten().tolist()
            print("MISMATCH", it, rep, (n, d, b, k), "queries", diff[:10], "flagged", int(flagged.sum()), flush=True)
            for qq in diff[:2]:
                print("  got ", out[1][qq].tolist(), "\n  want", ref[1][qq].tolist())
print("done", iters, "iterations,", bad, "bad")
