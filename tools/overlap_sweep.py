"""Step time of the graph-replayed hybrid search (10M x 768, batch B) for combinations of
overlap / dense ring depth / BM25 CTAs per SM.  One process: the knobs are environment
variables the library reads per call, the graph is re-captured per variant."""
import json, os, sys
import numpy as np
import torch
from classmate_rag_b200 import lexical, ops, synth
from classmate_rag_b200.engine import GraphedSearch, HybridEngine, SearchParams

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
b = int(sys.argv[2]) if len(sys.argv) > 2 else 32
variants = sys.argv[3:] or ["0:4:0", "1:4:0", "1:3:5", "1:3:4", "1:2:6", "1:2:5", "1:3:0"]
d, vocab, steps = 768, 30000, 12
emb = synth.dense_corpus(n, d, "cuda")
doc_ptr, tokens = synth.lexical_corpus(n, vocab, 64, "cuda")
lex = lexical.build_lexical_index(doc_ptr, tokens, vocab, tile_docs=int(os.environ.get("TILE", "2048")))
del doc_ptr, tokens
q, _ = synth.dense_queries(n, d, b * steps, "cuda")
terms = synth.lexical_queries(b * steps, vocab)
dev_terms = [tuple(t.cuda() for t in lexical.pack_queries(terms[s * b:(s + 1) * b])) for s in range(steps)]
p = SearchParams(top_k=10)
ref = None
for v in variants:
    ov, st, cap, *rest = v.split(":")
    os.environ["CMR_BM25_BATCH"] = rest[0] if rest else "0"
    os.environ["CMR_MM_STAGES"] = st
    os.environ["CMR_BM25_CTAS_PER_SM"] = cap
    eng = HybridEngine(emb, lex, overlap=ov in ("1", "2"))
    eng.bm25_first = ov != "2"     # overlap "2": dense pass enqueued first, BM25 after it
    g = GraphedSearch(eng, p, b, max_terms=16)
    for s in range(3):
        g.launch_resident(q[s * b:(s + 1) * b], *dev_terms[s])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(g.stream)
    for s in range(3, steps):
        out = g.launch_resident(q[s * b:(s + 1) * b], *dev_terms[s])
    e1.record(g.stream)
    torch.cuda.synchronize()
    res = [t.cpu().numpy().tobytes() for t in out]
    if ref is None:
        ref = res
    print(json.dumps({"overlap": ov, "stages": st, "bm25_ctas_per_sm": cap, "bm25_batch": os.environ["CMR_BM25_BATCH"], "tile": os.environ.get("TILE", "2048"), "ms_per_step": e0.elapsed_time(e1) / (steps - 3),
                      "same_as_first": res == ref}), flush=True)
    del g, eng
