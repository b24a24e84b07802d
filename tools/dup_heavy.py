"""Uncertified-query fallback on a duplicate-heavy corpus.
   python tools/dup_heavy.py ROWS [DIM] [CLUSTER] [DUP_FRACTION]
5 % of the rows (default) are exact copies in clusters of 50; half of the queries are aimed at a
cluster (their top-50 are exact ties), half at ordinary rows.  Reports, for the pool size the hybrid
step uses (24) and k = 10: the share of queries the fp32 pass flags, what the widest over-selection
(k = 120) still flags, and the time of each rung of the ladder (fast pass, wide pass, exhaustive scan)."""
import json
import sys

import numpy as np
import torch

from classmate_rag_b200 import ops, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 768
cluster = int(sys.argv[3]) if len(sys.argv) > 3 else 50
frac = float(sys.argv[4]) if len(sys.argv) > 4 else 0.05
dev = "cuda"
emb = synth.dense_corpus(n, d, dev)
g = torch.Generator(device="cpu").manual_seed(7)
n_clusters = int(n * frac / cluster)
members = torch.randperm(n, generator=g)[: n_clusters * cluster].view(n_clusters, cluster).to(dev)
emb[members[:, 1:].reshape(-1)] = emb[members[:, :1].expand(-1, cluster - 1).reshape(-1)]
nq = 64
seeds = torch.cat([members[torch.randperm(n_clusters, generator=g)[: nq // 2], 0].cpu(),
                   torch.randint(0, n, (nq // 2,), generator=g)])
q = torch.nn.functional.normalize(emb[seeds.to(dev)].float() + 0.5 * torch.randn(nq, d, device=dev) / d ** 0.5, dim=1)
qb = ops.f32_to_bf16(q)


def timed(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


out = {"rows": n, "dim": d, "cluster": cluster, "dup_fraction": frac, "queries": nq}
for k in (24, 10):
    for algo, batch in (("mma", 32), ("scan", 1)):
        fl = []
        for lo in range(0, nq, batch):
            f = ops.dense_topk(emb, qb[lo:lo + batch], k, algo=algo)[3]
            fl.append(f.clone())
        fl = torch.cat(fl).cpu().numpy()
        fw = []
        for lo in range(0, nq, batch):
            fw.append(ops.dense_topk(emb, qb[lo:lo + batch], ops.WIDE_K, algo=algo)[3].clone())
        fw = torch.cat(fw).cpu().numpy()
        key = f"k{k}_{algo}_b{batch}"
        out[key] = {"flagged_cluster_queries": float((fl[: nq // 2] != 0).mean()),
                    "flagged_ordinary_queries": float((fl[nq // 2:] != 0).mean()),
                    "still_flagged_at_k120": float((fw != 0).mean()),
                    "fast_ms": timed(lambda: ops.dense_topk(emb, qb[:batch], k, algo=algo)),
                    "wide_ms": timed(lambda: ops.dense_topk(emb, qb[:batch], ops.WIDE_K, algo=algo)),
                    "exhaustive_ms": timed(lambda: ops.dense_topk(emb, qb[:batch], k, algo="exact"), iters=2)}
# the ladder end to end == the exhaustive scan, bit for bit
s1, i1, c1, f1 = [t.clone() for t in ops.dense_topk_certified(emb, qb[:32], 24, algo="mma")]
s2, i2, c2, f2 = [t.clone() for t in ops.dense_topk(emb, qb[:32], 24, algo="exact")]
torch.cuda.synchronize()
out["ladder_equals_exhaustive"] = bool(torch.equal(i1, i2) and s1.cpu().numpy().tobytes() == s2.cpu().numpy().tobytes()
                                       and int(f1.sum()) == 0)
print(json.dumps(out))
