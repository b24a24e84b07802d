"""Stage times of a new filter's BM25 statistics (_FilteredView) on a synthetic index:
   CMRAG_PROFILE_FILTER=1 python tools/filter_view_profile.py [docs] [keep_one_in]"""
import os, sys, time
from types import SimpleNamespace
import torch
os.environ.setdefault("CMRAG_PROFILE_FILTER", "1")
from classmate_rag_b200 import lexical, synth
from classmate_rag_b200.retrieval.bm25_store import _FilteredView

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
one_in = int(sys.argv[2]) if len(sys.argv) > 2 else 16
doc_ptr, tokens = synth.lexical_corpus(n, 30000, 64, "cuda")
lex = lexical.build_lexical_index(doc_ptr, tokens, 30000)
full = SimpleNamespace(lex=lex, rows=None)
for rep in range(3):
    mask = ((torch.arange(n, device="cuda") + rep) % one_in == 0).to(torch.uint8)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    v = _FilteredView(full, mask, doc_ptr, tokens)
    torch.cuda.synchronize()
    print(f"docs {n}, postings {lex.n_postings}, subset {v.n_docs}: {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
