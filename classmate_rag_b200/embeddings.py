"""E5 query / passage encoder with the reference's contract (rag/embeddings/__init__.py:37-110).

The encoder itself stays PyTorch (north star: only the search path is hand-written CUDA).
What the retrieval path consumes is the CONTRACT: ``encode_queries(list[str]) -> float32
[B, D]`` with the "query: " prefix, masked mean pooling and L2 normalisation (D = 768 for
multilingual-e5-base, 1024 for -large); ``encode_passages`` with "passage: ".

Differences from the reference wrapper: it runs the Hugging Face ``transformers`` model
directly (sentence-transformers is not required), keeps the model resident on the GPU, and
offers ``encode_queries_device`` -- the unit-norm fp32 rows as a CUDA tensor, so a caller
that feeds ``HybridEngine`` / ``GraphedSearch`` avoids the device -> host -> device hop of the
reference (N4 of SURVEY.md section 8f).
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import numpy as np
import torch


def _resolve_cache_dir() -> Optional[str]:
    """SENTENCE_TRANSFORMERS_HOME, HUGGINGFACE_HUB_CACHE, HF_HOME -- first one set."""
    for key in ("SENTENCE_TRANSFORMERS_HOME", "HUGGINGFACE_HUB_CACHE", "HF_HOME"):
        v = os.getenv(key)
        if v and v.strip():
            return os.path.abspath(os.path.expanduser(v))
    return None


class E5MultilingualEmbedder:
    def __init__(self, model_name: str = "intfloat/multilingual-e5-base", device: Optional[str] = None,
                 normalize: bool = True, *, model=None, tokenizer=None, max_length: int = 512,
                 batch_size: int = 32) -> None:
        """``model`` / ``tokenizer``: inject already-built objects (tests, custom checkpoints);
        otherwise both are loaded with ``from_pretrained`` (local cache first: the reference
        resolves the same cache directories)."""
        self.normalize = bool(normalize)
        self.max_length, self.batch_size = int(max_length), int(batch_size)
        self.device = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
        if model is None or tokenizer is None:
            from transformers import AutoModel, AutoTokenizer
            kw = {"cache_dir": _resolve_cache_dir()} if _resolve_cache_dir() else {}
            token = os.getenv("HF_TOKEN") or None
            if token:
                kw["token"] = token
            tokenizer = tokenizer or AutoTokenizer.from_pretrained(model_name, **kw)
            model = model or AutoModel.from_pretrained(model_name, **kw)
        self.tokenizer = tokenizer
        self.model = model.to(self.device).eval()

    @staticmethod
    def _fmt_queries(queries: Iterable[str]) -> List[str]:
        return [f"query: {q}" for q in queries]

    @staticmethod
    def _fmt_passages(texts: Iterable[str]) -> List[str]:
        return [f"passage: {t}" for t in texts]

    @torch.no_grad()
    def _encode(self, texts: List[str]) -> torch.Tensor:
        out = []
        for lo in range(0, len(texts), self.batch_size):
            enc = self.tokenizer(texts[lo:lo + self.batch_size], padding=True, truncation=True,
                                 max_length=self.max_length, return_tensors="pt")
            enc = {k: v.to(self.device) for k, v in enc.items()}
            hidden = self.model(**enc).last_hidden_state.float()
            mask = enc["attention_mask"].unsqueeze(-1).float()
            pooled = (hidden * mask).sum(1) / mask.sum(1).clamp(min=1e-9)     # masked mean pooling
            if self.normalize:
                pooled = torch.nn.functional.normalize(pooled, p=2, dim=1)
            out.append(pooled)
        if not out:
            d = int(getattr(self.model.config, "hidden_size", 0))
            return torch.zeros((0, d), dtype=torch.float32, device=self.device)
        return torch.cat(out, 0)

    # -- reference API ---------------------------------------------------------------------
    def encode_queries(self, queries: Iterable[str]) -> np.ndarray:
        return self._encode(self._fmt_queries(queries)).cpu().numpy().astype("float32", copy=False)

    def encode_passages(self, texts: Iterable[str]) -> np.ndarray:
        return self._encode(self._fmt_passages(texts)).cpu().numpy().astype("float32", copy=False)

    # -- device-resident forms (no host hop) ---------------------------------------------------
    def encode_queries_device(self, queries: Iterable[str]) -> torch.Tensor:
        return self._encode(self._fmt_queries(queries))

    def encode_passages_device(self, texts: Iterable[str]) -> torch.Tensor:
        return self._encode(self._fmt_passages(texts))


__all__ = ["E5MultilingualEmbedder"]
