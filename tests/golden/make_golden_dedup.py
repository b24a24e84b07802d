#!/usr/bin/env python
"""Generate tests/golden/reference_dedup.json by running the LIVE reference text dedup
(/root/reference/rag/utils/dedup.py: dedup_text_blocks, Jaccard on token 5-gram shingles,
greedy keep-first) on synthetic chunk lists.  Build container only; the vectors are committed.

Each case stores the blocks and, per threshold, the INDICES of the blocks the reference kept
(its output is a subsequence of the input, so the indices are recovered by in-order matching).
"""
from __future__ import annotations

import importlib.util
import json
from pathlib import Path

import numpy as np

REF = Path("/root/reference/rag/utils/dedup.py")
spec = importlib.util.spec_from_file_location("ref_dedup", REF)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

WORDS = ("gradient descent matrix vector tensor kernel memory bandwidth cache latency pipeline fusion "
         "retrieval ranking lexical dense sparse token shard stripe posting index query chunk embedding "
         "cosine neighbor lecture exam course unit integral derivative theorem proof lemma corollary entropy "
         "compiler parser lattice algebra topology manifold quantum photon été naïve Hôpital straße "
         "x86 b2b 3x2 λ-calculus über").split()
PUNCT = [",", ".", ";", ":", "!", "?", " - ", "(", ")", "'", "\"", "\n", "\t", "  "]


def sentence(rng, n):
    out = []
    for _ in range(n):
        w = WORDS[int(rng.integers(0, len(WORDS)))]
        r = rng.random()
        if r < 0.1:
            w = w.upper()
        elif r < 0.2:
            w = w.capitalize()
        out.append(w)
        if rng.random() < 0.15:
            out.append(PUNCT[int(rng.integers(0, len(PUNCT)))])
    return " ".join(out)


def mutate(rng, text, n_edits):
    toks = text.split(" ")
    for _ in range(n_edits):
        i = int(rng.integers(0, len(toks)))
        r = rng.random()
        if r < 0.5:
            toks[i] = WORDS[int(rng.integers(0, len(WORDS)))]
        elif r < 0.75:
            toks.insert(i, WORDS[int(rng.integers(0, len(WORDS)))])
        elif len(toks) > 1:
            del toks[i]
    return " ".join(toks)


def make_blocks(rng, n, lo, hi):
    blocks = []
    for _ in range(n):
        r = rng.random()
        if blocks and r < 0.25:       # near duplicate of an earlier block
            src = blocks[int(rng.integers(0, len(blocks)))]
            blocks.append(mutate(rng, src, int(rng.integers(0, 4))) if src.strip() else src)
        elif blocks and r < 0.32:     # exact duplicate, possibly with other case / punctuation only
            src = blocks[int(rng.integers(0, len(blocks)))]
            blocks.append(src.upper() if rng.random() < 0.5 else src + " ...")
        elif r < 0.36:
            blocks.append(["", "   ", "...", "\n\t", "?!"][int(rng.integers(0, 5))])
        elif r < 0.42:                # fewer than 5 tokens: one short shingle
            blocks.append(sentence(rng, int(rng.integers(1, 5))))
        else:
            blocks.append(sentence(rng, int(rng.integers(lo, hi))))
    return blocks


def kept_indices(blocks, kept):
    idx, pos = [], 0
    for t in kept:
        while blocks[pos] != t:
            pos += 1
        idx.append(pos)
        pos += 1
    return idx


def main():
    rng = np.random.default_rng(20260118)
    cases = []
    for n, lo, hi in ((12, 6, 20), (40, 8, 40), (120, 10, 80), (300, 20, 120), (64, 5, 9)):
        blocks = make_blocks(rng, n, lo, hi)
        kept = {}
        for thr in (0.92, 0.5, 0.8, 1.0, 0.0, 1.5):
            kept[repr(thr)] = kept_indices(blocks, ref.dedup_text_blocks(list(blocks), jaccard_threshold=thr))
        cases.append({"blocks": blocks, "kept_by_threshold": kept})
    cases.append({"blocks": [], "kept_by_threshold": {"0.92": []}})
    toks = [{"text": t, "tokens": ref._norm_tokens(t)} for t in
            ["The Chain-rule's été, l'Hôpital 3x2!", "  A  b\tC\n", "", "snake_case under_score 42", "ÜBER straße λ-calculus"]]
    out = {"cases": cases, "norm_tokens": toks}
    path = Path(__file__).with_name("reference_dedup.json")
    path.write_text(json.dumps(out, ensure_ascii=False))
    print(path, len(cases), "corpora;", sum(len(c["blocks"]) - len(k) for c in cases for k in c["kept_by_threshold"].values()),
          "blocks dropped in total")


if __name__ == "__main__":
    main()
