"""CPU: the E5 encoder wrapper keeps the reference's contract (prefixes, masked mean pooling,
unit rows, float32 [B, D]); a tiny randomly initialised XLM-R stands in for the checkpoint."""
import numpy as np
import torch


class _Tok:
    """Whitespace tokenizer with the Hugging Face call signature the wrapper uses."""

    def __init__(self):
        self.seen = []

    def __call__(self, texts, padding=True, truncation=True, max_length=512, return_tensors="pt"):
        self.seen.extend(texts)
        ids = [[2] + [3 + (hash(w) % 900) for w in t.split()][: max_length - 2] + [1] for t in texts]
        width = max(len(x) for x in ids)
        input_ids = torch.tensor([x + [0] * (width - len(x)) for x in ids])
        mask = torch.tensor([[1] * len(x) + [0] * (width - len(x)) for x in ids])
        return {"input_ids": input_ids, "attention_mask": mask}


def _tiny():
    from transformers import XLMRobertaConfig, XLMRobertaModel
    torch.manual_seed(0)
    cfg = XLMRobertaConfig(vocab_size=1000, hidden_size=64, num_hidden_layers=2, num_attention_heads=4,
                           intermediate_size=128, max_position_embeddings=140)
    return XLMRobertaModel(cfg)


def test_contract_prefix_pooling_norm():
    from classmate_rag_b200.embeddings import E5MultilingualEmbedder
    tok = _Tok()
    emb = E5MultilingualEmbedder(model=_tiny(), tokenizer=tok, device="cpu", batch_size=2)
    q = emb.encode_queries(["what is a gradient", "kernel", "memory bandwidth of HBM"])
    assert q.dtype == np.float32 and q.shape == (3, 64)
    assert np.allclose(np.linalg.norm(q, axis=1), 1.0, atol=1e-5)
    assert tok.seen == ["query: what is a gradient", "query: kernel", "query: memory bandwidth of HBM"]
    p = emb.encode_passages(["kernel"])
    assert tok.seen[-1] == "passage: kernel" and not np.allclose(p[0], q[1])      # the prefix matters
    # padding must not change a row: encoded alone == encoded in a ragged batch
    alone = emb.encode_queries(["kernel"])
    assert np.allclose(alone[0], q[1], atol=1e-5)
    dev = emb.encode_queries_device(["kernel"])
    assert isinstance(dev, torch.Tensor) and np.allclose(dev.numpy(), alone, atol=1e-6)
    assert emb.encode_queries([]).shape == (0, 64)
    raw = E5MultilingualEmbedder(model=_tiny(), tokenizer=_Tok(), device="cpu", normalize=False).encode_queries(["kernel"])
    assert not np.isclose(np.linalg.norm(raw[0]), 1.0, atol=1e-3)


def test_fixed_shape_encoder_equals_wrapper_on_cpu():
    """GraphedQueryEncoder's host logic (padding to the captured width, empty rows, truncation)
    on the CPU in eager mode: same rows as the wrapper."""
    from classmate_rag_b200.embeddings import E5MultilingualEmbedder, GraphedQueryEncoder
    emb = E5MultilingualEmbedder(model=_tiny(), tokenizer=_Tok(), device="cpu")
    enc = GraphedQueryEncoder(emb, n_queries=4, max_tokens=16)
    assert enc.graph is None
    qs = ["what is a gradient", "kernel", "memory bandwidth of HBM"]
    got = enc(qs).numpy().copy()
    want = emb.encode_queries(qs)
    assert got.shape == (4, 64) and np.allclose(got[:3], want, atol=1e-5)
    again = enc(["kernel"]).numpy()
    assert np.allclose(again[0], want[1], atol=1e-5)
    long_q = " ".join(["token"] * 60)
    assert np.allclose(enc([long_q]).numpy()[0], E5MultilingualEmbedder(
        model=emb.model, tokenizer=_Tok(), device="cpu", max_length=16).encode_queries([long_q])[0], atol=1e-5)
    import pytest
    with pytest.raises(ValueError):
        enc(["a"] * 5)

