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


class GraphedQueryEncoder:
    """Fixed-shape query encode for the serving loop (N4 of SURVEY.md section 8f): the E5
    forward for ``n_queries`` x ``max_tokens`` token ids is captured in a CUDA graph next to the
    search graph, reads its ids from static device buffers and writes the unit-norm float32 rows
    into ``out`` on the device -- e.g. straight into ``GraphedSearch.q_f32`` -- so a question goes
    text -> ids (host tokeniser) -> one H2D copy of the ids -> encoder graph -> search graph
    without the device -> host -> device hop of the reference (rag/pipeline/rag.py:533-545 loads
    the model per ask and passes NumPy arrays).  The model stays PyTorch (north star).

    If the model's forward cannot be captured (a data-dependent host check inside the
    attention-mask preparation of some transformers versions), the encoder runs eagerly on the
    same buffers: ``graph`` is then None.  Results are identical either way."""

    def __init__(self, embedder: E5MultilingualEmbedder, n_queries: int, max_tokens: int = 64, *,
                 out: Optional[torch.Tensor] = None, stream=None, use_graph: bool = True) -> None:
        self.emb, self.b, self.l = embedder, int(n_queries), int(max_tokens)
        dev = embedder.device
        if dev.type == "cuda" and dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        d = int(embedder.model.config.hidden_size)
        pad_id = getattr(embedder.model.config, "pad_token_id", None)
        self.pad_id = 0 if pad_id is None else int(pad_id)
        self.ids = torch.full((self.b, self.l), self.pad_id, dtype=torch.long, device=dev)
        self.mask = torch.zeros((self.b, self.l), dtype=torch.long, device=dev)
        self.mask[:, 0] = 1                      # warm-up runs need a non-empty row
        self.h_ids = torch.full((self.b, self.l), self.pad_id, dtype=torch.long)
        self.h_mask = torch.zeros((self.b, self.l), dtype=torch.long)
        if dev.type == "cuda":
            self.h_ids, self.h_mask = self.h_ids.pin_memory(), self.h_mask.pin_memory()
        self.out = out if out is not None else torch.zeros((self.b, d), dtype=torch.float32, device=dev)
        if tuple(self.out.shape) != (self.b, d) or self.out.dtype != torch.float32 or self.out.device != dev:
            raise ValueError("out must be a float32 [n_queries, hidden_size] tensor on the encoder's device")
        self.stream = stream
        self.graph = None
        self.capture_error: Optional[str] = None
        if use_graph and dev.type == "cuda":
            self._capture()

    @torch.no_grad()
    def _forward(self) -> None:
        hidden = self.emb.model(input_ids=self.ids, attention_mask=self.mask).last_hidden_state.float()
        m = self.mask.unsqueeze(-1).float()
        pooled = (hidden * m).sum(1) / m.sum(1).clamp(min=1e-9)
        if self.emb.normalize:
            pooled = torch.nn.functional.normalize(pooled, p=2, dim=1)
        self.out.copy_(pooled)

    def _capture(self) -> None:
        if self.stream is None:
            self.stream = torch.cuda.Stream(device=self.device)
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            for _ in range(2):
                self._forward()
            self.stream.synchronize()
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.stream):
                    self._forward()
                self.graph = g
            except Exception as e:  # capture refused: stay eager (same buffers, same result)
                self.graph = None
                self.capture_error = f"{type(e).__name__}: {e}"
                torch.cuda.synchronize(self.device)

    def set_texts(self, queries: List[str]) -> None:
        """Tokenise ("query: " prefix, truncation to max_tokens) into the pinned host buffers;
        fewer than n_queries texts leave the remaining rows empty (their output rows are unused)."""
        if len(queries) > self.b:
            raise ValueError(f"{len(queries)} queries for an encoder captured for {self.b}")
        self.h_ids.fill_(self.pad_id)
        self.h_mask.zero_()
        self.h_mask[:, 0] = 1
        if queries:
            enc = self.emb.tokenizer(self.emb._fmt_queries(queries), padding=True, truncation=True,
                                     max_length=self.l, return_tensors="pt")
            ids, mask = enc["input_ids"][:, : self.l], enc["attention_mask"][:, : self.l]
            self.h_ids[: ids.shape[0], : ids.shape[1]] = ids
            self.h_mask[: ids.shape[0]] = 0
            self.h_mask[: ids.shape[0], : ids.shape[1]] = mask

    def launch(self) -> torch.Tensor:
        """H2D copy of the ids + the encoder (graph replay or eager) on ``stream``; returns
        ``out`` (valid in stream order, nothing is synchronised)."""
        if self.device.type != "cuda":
            self.ids.copy_(self.h_ids)
            self.mask.copy_(self.h_mask)
            self._forward()
            return self.out
        if self.stream is None:
            self.stream = torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            self.ids.copy_(self.h_ids, non_blocking=True)
            self.mask.copy_(self.h_mask, non_blocking=True)
            if self.graph is not None:
                self.graph.replay()
            else:
                self._forward()
        return self.out

    def __call__(self, queries: List[str]) -> torch.Tensor:
        self.set_texts(queries)
        return self.launch()


__all__ = ["E5MultilingualEmbedder", "GraphedQueryEncoder"]
