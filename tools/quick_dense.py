"""Quick device timing of the dense scan (development aid, not the bench)."""
import sys, json
import torch
from classmate_rag_b200 import ops

def run(n, d, k, iters=20):
    g = torch.Generator(device="cuda").manual_seed(1)
    emb = torch.empty((n, d), dtype=torch.bfloat16, device="cuda")
    step = 1 << 18
    for lo in range(0, n, step):
        x = torch.randn((min(step, n - lo), d), generator=g, device="cuda")
        emb[lo:lo + x.shape[0]] = torch.nn.functional.normalize(x, dim=1).to(torch.bfloat16)
    q = torch.nn.functional.normalize(emb[12345].float() + 0.5 * torch.randn(d, device="cuda") / d ** 0.5, dim=0).to(torch.bfloat16)[None]
    ws = ops.DenseWorkspace(n, d, 1, k, emb.device)
    for _ in range(3):
        out = ops.dense_topk(emb, q, k, workspace=ws)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        ops.dense_topk(emb, q, k, workspace=ws)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    med = ts[len(ts) // 2]
    gbs = n * d * 2 / (med * 1e-3) / 1e9
    print(json.dumps({"n": n, "d": d, "k": k, "ms_med": med, "ms_min": ts[0], "GBps": gbs, "ids": out[1][0, :5].tolist(), "flags": out[3].tolist()}))

if __name__ == "__main__":
    print(ops.device_info())
    for n, d, k in [(1_000_000, 768, 10), (1_000_000, 768, 24), (10_000_000, 768, 10), (4_000_000, 1024, 10), (1_000_000, 768, 100)]:
        run(n, d, k)
