"""CPU: the C restatement equals the NumPy restatement bit for bit."""
import numpy as np
import torch

from oracle import c_oracle, np_oracle as o
from classmate_rag_b200 import lexical
from tests.synth_small import zipf_corpus


def test_c_exact_dots_equal_numpy():
    rng = np.random.default_rng(0)
    for d in (8, 64, 264, 768, 1024):
        c = o.f32_to_bf16_bits(rng.standard_normal((300, d)).astype(np.float32))
        q = o.f32_to_bf16_bits(rng.standard_normal(d).astype(np.float32))
        assert c_oracle.exact_dots(q, c).tobytes() == o.exact_dots(q, c).tobytes()


def test_c_bm25_equals_numpy():
    docs, doc_ptr, tokens, v = zipf_corpus(seed=3, n_docs=2000, vocab=120, mean_len=10)
    ix = lexical.build_lexical_index(doc_ptr, tokens, v, device="cpu", tile_docs=512)
    tf = (ix.post_tf.to(torch.int32) & 0xFFFF).numpy()
    for q in ([0, 3, 3, 119, -1], [], [5]):
        a = c_oracle.bm25_scores(ix.term_ptr.numpy(), ix.post_doc.numpy(), tf, ix.doc_len.numpy(), ix.idf_host, ix.avgdl, q)
        b = o.bm25_scores_csr(ix.term_ptr.numpy(), ix.post_doc.numpy(), tf, ix.doc_len.numpy(), ix.idf_host, ix.avgdl, q)
        assert a.tobytes() == b.tobytes()
