"""GPU: the fixed-shape E5 query encoder (CUDA graph, device-resident output) in front of the
search graph (N4): same rows as the eager wrapper, and feeding GraphedSearch on the device gives
the same hits as the reference-style NumPy hand-over."""
import numpy as np
import pytest
import torch

from tests.test_embeddings_cpu import _Tok, _tiny

pytestmark = pytest.mark.gpu


def test_graphed_query_encoder_feeds_search_graph():
    from classmate_rag_b200 import lexical, synth
    from classmate_rag_b200.embeddings import E5MultilingualEmbedder, GraphedQueryEncoder
    from classmate_rag_b200.engine import GraphedSearch, HybridEngine, SearchParams
    emb = E5MultilingualEmbedder(model=_tiny(), tokenizer=_Tok(), device="cuda")
    n, d, vocab, b = 6000, 64, 300, 4
    corpus = synth.dense_corpus(n, d, "cuda")
    doc_ptr, tokens = synth.lexical_corpus(n, vocab, 24, "cuda")
    eng = HybridEngine(corpus, lexical.build_lexical_index(doc_ptr, tokens, vocab))
    gs = GraphedSearch(eng, SearchParams(top_k=10), b, max_terms=16)
    enc = GraphedQueryEncoder(emb, b, 16, out=gs.q_f32, stream=gs.stream)
    eager = GraphedQueryEncoder(emb, b, 16, use_graph=False)
    qs = ["what is a gradient", "kernel", "memory bandwidth of HBM", "posting list"]
    terms = [[1, 2, 3], [5], [7, 8, 250], [0, 0]]
    rows = enc(qs)
    gs.stream.synchronize()
    rows = rows.clone()
    assert np.allclose(rows.cpu().numpy(), emb.encode_queries(qs), atol=1e-4)
    assert np.allclose(rows.cpu().numpy(), eager(qs).cpu().numpy(), atol=1e-5)
    assert np.allclose(rows.norm(dim=1).cpu().numpy(), 1.0, atol=1e-5)
    # reference-style hand-over: NumPy rows through the host buffers
    want = [a.copy() for a in gs(rows.cpu().numpy(), terms)]
    # device hand-over: encoder graph writes gs.q_f32, the search graph replays on the same stream
    qt, qp = [t.cuda() for t in lexical.pack_queries(terms)]
    enc(qs)
    out = gs.launch_resident(gs.q_f32, qt, qp)
    gs.stream.synchronize()
    for a, w in zip(out, want):
        assert a.cpu().numpy().tobytes() == w.tobytes()
    # a second batch through the same graphs
    qs2 = ["entropy of a photon", "lattice", "compiler parser", "x"]
    enc(qs2)
    out2 = gs.launch_resident(gs.q_f32, qt, qp)
    gs.stream.synchronize()
    out2 = [t.clone() for t in out2]
    want2 = [a.copy() for a in gs(eager(qs2).cpu().numpy(), terms)]
    assert out2[0].cpu().numpy().tolist() == want2[0].tolist()
    print("encoder graph:", "captured" if enc.graph is not None else f"eager ({enc.capture_error})")
