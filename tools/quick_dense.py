"""Quick device timing of the dense paths (development aid, not the bench).

    python tools/quick_dense.py [n_rows] [dim]      # table of (algo, batch) timings
"""
import json
import sys

import torch

from classmate_rag_b200 import ops


def corpus(n, d):
    g = torch.Generator(device="cuda").manual_seed(1)
    emb = torch.empty((n, d), dtype=torch.bfloat16, device="cuda")
    step = 1 << 18
    for lo in range(0, n, step):
        x = torch.randn((min(step, n - lo), d), generator=g, device="cuda")
        emb[lo:lo + x.shape[0]] = torch.nn.functional.normalize(x, dim=1).to(torch.bfloat16)
    return emb


def run(emb, b, k, algo, iters=10):
    n, d = emb.shape
    rows = torch.randint(0, n, (b,), device="cuda")
    q = torch.nn.functional.normalize(emb[rows].float() + 0.5 * torch.randn(b, d, device="cuda") / d ** 0.5, dim=1).to(torch.bfloat16)
    ws = ops.DenseWorkspace(n, d, b, k, emb.device)
    for _ in range(3):
        out = ops.dense_topk(emb, q, k, workspace=ws, algo=algo)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        ops.dense_topk(emb, q, k, workspace=ws, algo=algo)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    med = ts[len(ts) // 2]
    hit = float((out[1][:, 0] == rows).float().mean())
    print(json.dumps({"algo": algo, "n": n, "d": d, "b": b, "k": k, "ms_med": round(med, 4), "ms_min": round(ts[0], 4),
                      "GBps_matrix_once": round(n * d * 2 / (med * 1e-3) / 1e9, 1),
                      "TFLOPs": round(2.0 * b * n * d / (med * 1e-3) / 1e12, 1), "qps": round(b / (med * 1e-3)),
                      "top1_is_planted": hit, "flags": int(out[3].sum())}), flush=True)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 768
    print(ops.device_info(), flush=True)
    emb = corpus(n, d)
    for b, algo in [(1, "scan"), (8, "scan"), (16, "scan"), (32, "scan"), (9, "mma"), (32, "mma"), (128, "mma"),
                    (256, "mma"), (1024, "mma")]:
        run(emb, b, 10, algo)
