"""Where the host time of PipelinedSearch.submit goes (set_queries / launch / result), per step.
   python tools/host_path_profile.py [rows] [batch]"""
import sys, time
import numpy as np
import torch
from classmate_rag_b200 import lexical, synth
from classmate_rag_b200.engine import HybridEngine, PipelinedSearch, SearchParams

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
b = int(sys.argv[2]) if len(sys.argv) > 2 else 128
emb = synth.dense_corpus(n, 768, "cuda")
doc_ptr, tokens = synth.lexical_corpus(n, 30000, 64, "cuda")
lex = lexical.build_lexical_index(doc_ptr, tokens, 30000)
eng = HybridEngine(emb, lex)
steps = 40
q, _ = synth.dense_queries(n, 768, b * steps, "cuda")
qh = q.cpu().numpy()
terms = synth.lexical_queries(b * steps, 30000)
ps = PipelinedSearch(eng, SearchParams(top_k=10), b, max_terms=16)
t = {"set_queries": 0.0, "launch": 0.0, "result": 0.0, "copy": 0.0}
for s in range(steps):
    g = ps.slots[s & 1]
    if s >= 2:
        t0 = time.perf_counter(); r = g.result(); t1 = time.perf_counter(); r = tuple(a.copy() for a in r); t2 = time.perf_counter()
        if s >= 10:
            t["result"] += t1 - t0; t["copy"] += t2 - t1
    t0 = time.perf_counter()
    g.set_queries(qh[s * b:(s + 1) * b], terms[s * b:(s + 1) * b])
    t1 = time.perf_counter()
    g.launch()
    t2 = time.perf_counter()
    if s >= 10:
        t["set_queries"] += t1 - t0; t["launch"] += t2 - t1
torch.cuda.synchronize()
print({k: round(v / (steps - 10) * 1e6, 1) for k, v in t.items()}, "us per step; torch threads", torch.get_num_threads())
