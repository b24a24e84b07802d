"""Soak of PipelinedSearch: many steps with different batches through the two-slot pipeline (host-buffer
and resident forms) against the eager engine, byte for byte.  python tools/pipeline_soak.py [steps]"""
import sys
import numpy as np
import torch
from classmate_rag_b200 import lexical, ops, synth
from classmate_rag_b200.engine import HybridEngine, PipelinedSearch, SearchParams

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n, d, vocab, b = 400_000, 768, 5000, 32
emb = synth.dense_corpus(n, d, "cuda")
doc_ptr, tokens = synth.lexical_corpus(n, vocab, 32, "cuda")
lex = lexical.build_lexical_index(doc_ptr, tokens, vocab)
eng = HybridEngine(emb, lex)
p = SearchParams(top_k=10)
n_sets = 8
q, _ = synth.dense_queries(n, d, b * n_sets, "cuda")
terms = synth.lexical_queries(b * n_sets, vocab)
want, dev_in = [], []
for s in range(n_sets):
    qt, qp = lexical.pack_queries(terms[s * b:(s + 1) * b])
    qt, qp = qt.cuda(), qp.cuda()
    out = eng.search(ops.f32_to_bf16(q[s * b:(s + 1) * b]), qt, qp, p)
    torch.cuda.synchronize()
    want.append([t.cpu().numpy().tobytes() for t in out])
    dev_in.append((q[s * b:(s + 1) * b].contiguous(), qt, qp))
qh = q.cpu().numpy()
ps = PipelinedSearch(eng, p, b, max_terms=16)
bad = 0
order = np.random.default_rng(0).integers(0, n_sets, size=steps)
pending = []
for i, s in enumerate(order):
    r = ps.submit(qh[s * b:(s + 1) * b], terms[s * b:(s + 1) * b])
    pending.append(s)
    if r is not None:
        w = want[pending.pop(0)]
        bad += [a.tobytes() for a in r] != w
r = ps.drain()
bad += [a.tobytes() for a in r] != want[pending.pop(0)]
print("host-buffer form:", steps, "steps,", bad, "mismatching")
bad2 = 0
prev = None
for i, s in enumerate(order):
    out = ps.launch_resident(*dev_in[s])
    if prev is not None and i % 7 == 0:      # spot checks need a sync; most steps stay back to back
        ps.wait(); torch.cuda.synchronize()
        bad2 += [t.cpu().numpy().tobytes() for t in prev[0]] != want[prev[1]]
    prev = (out, s)
ps.wait(); torch.cuda.synchronize()
bad2 += [t.cpu().numpy().tobytes() for t in prev[0]] != want[prev[1]]
print("resident form:", steps, "steps,", bad2, "mismatching")
sys.exit(1 if bad or bad2 else 0)
