"""N4: ingest-side text dedup (reference rag/utils/dedup.py:40-55).  CPU: the oracle and the host
shingling against golden vectors from the LIVE reference (tests/golden/make_golden_dedup.py).
GPU: cmr_jaccard_edges + cmr_neardup_resolve through the drop-in function."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import np_oracle as o

GOLD = json.loads((Path(__file__).parent / "golden" / "reference_dedup.json").read_text())


def test_oracle_matches_reference_golden():
    for case in GOLD["cases"]:
        for thr, kept in case["kept_by_threshold"].items():
            assert o.dedup_keep_indices(case["blocks"], float(thr)) == kept
    for t in GOLD["norm_tokens"]:
        assert o.dedup_norm_tokens(t["text"]) == t["tokens"]


def test_host_shingle_sets_equal_oracle_sets():
    from classmate_rag_b200.retrieval import dedup
    for t in GOLD["norm_tokens"]:
        assert dedup.norm_tokens(t["text"]) == t["tokens"]
    blocks = GOLD["cases"][2]["blocks"]
    ptr, items = dedup.shingle_sets(blocks)
    assert ptr.dtype == np.int32 and items.dtype == np.int32 and ptr[0] == 0 and ptr[-1] == items.size
    sets = [o.dedup_shingles(o.dedup_norm_tokens(b)) for b in blocks]
    for i, s in enumerate(sets):
        mine = items[ptr[i]:ptr[i + 1]]
        assert mine.size == len(s) and (np.diff(mine) > 0).all()
    # equal ids <=> equal shingles: intersections counted on ids equal set intersections
    for i, j in ((5, 3), (40, 11), (100, 99), (7, 7)):
        a, b = items[ptr[i]:ptr[i + 1]], items[ptr[j]:ptr[j + 1]]
        assert np.intersect1d(a, b).size == len(sets[i] & sets[j])


def test_id_sets_plus_pair_rule_reproduce_the_reference_on_random_text():
    """Host side end to end on the CPU: the id sets the GPU receives, pushed through the pair rule
    the kernel implements (integer intersection, float64 inter/union >= thr, empty-set conventions)
    and the greedy keep-first pass, give the oracle's kept indices on random punctuated /
    accented / repeated text."""
    from classmate_rag_b200.retrieval import dedup
    rng = np.random.default_rng(99)
    alphabet = ["alpha", "Beta", "GAMMA", "été", "naïve", "x1", "under_score", "l'Hôpital", "a-b", "42", "ß"]
    seps = [" ", "  ", ", ", ". ", "\n", " - ", "!", "\t"]
    for trial in range(20):
        blocks = []
        for i in range(int(rng.integers(2, 40))):
            if blocks and rng.random() < 0.4:
                toks = blocks[int(rng.integers(0, len(blocks)))].split(" ")
                if toks and rng.random() < 0.7:
                    toks[int(rng.integers(0, len(toks)))] = alphabet[int(rng.integers(0, len(alphabet)))]
                blocks.append(" ".join(toks))
            else:
                n = int(rng.integers(0, 25))
                blocks.append("".join(alphabet[int(rng.integers(0, len(alphabet)))] + seps[int(rng.integers(0, len(seps)))]
                                      for _ in range(n)))
        ptr, items = dedup.shingle_sets(blocks)
        sets = [items[ptr[i]:ptr[i + 1]] for i in range(len(blocks))]
        for thr in (0.92, 0.6, 1.0):
            keep = []
            for i, a in enumerate(sets):
                dup = False
                for j in keep:
                    b = sets[j]
                    if a.size == 0 and b.size == 0:
                        jac = 1.0
                    elif a.size == 0 or b.size == 0:
                        jac = 0.0
                    else:
                        inter = np.intersect1d(a, b).size
                        jac = float(inter) / float(a.size + b.size - inter)
                    if jac >= thr:
                        dup = True
                        break
                if not dup:
                    keep.append(i)
            assert keep == o.dedup_keep_indices(blocks, thr), (trial, thr)


@pytest.mark.gpu
def test_dedup_text_blocks_matches_reference_golden():
    from classmate_rag_b200.retrieval import dedup
    for case in GOLD["cases"]:
        blocks = case["blocks"]
        for thr, kept in case["kept_by_threshold"].items():
            got = dedup.dedup_text_blocks(list(blocks), jaccard_threshold=float(thr))
            assert got == [blocks[i] for i in kept], (len(blocks), thr)
            mask = dedup.dedup_keep_mask(blocks, jaccard_threshold=float(thr))
            assert np.flatnonzero(mask).tolist() == kept


@pytest.mark.gpu
def test_dedup_larger_file_and_long_chunks_vs_oracle():
    """Per-file scale upwards: 1500 chunks, some longer than the kernel's shared-memory staging
    (> 8192 shingles), chains of near-copies, thresholds around the copies' similarity."""
    from classmate_rag_b200.retrieval import dedup
    rng = np.random.default_rng(5)
    vocab = [f"w{i}" for i in range(400)]
    blocks = []
    for i in range(1500):
        r = rng.random()
        if blocks and r < 0.3:
            toks = blocks[int(rng.integers(0, len(blocks)))].split(" ")
            for _ in range(int(rng.integers(0, 3))):
                toks[int(rng.integers(0, len(toks)))] = vocab[int(rng.integers(0, 400))]
            blocks.append(" ".join(toks))
        else:
            n = 9000 if i in (10, 700) else int(rng.integers(3, 90))
            blocks.append(" ".join(vocab[int(t)] for t in rng.integers(0, 400, n)))
    blocks[701] = blocks[700]            # exact copy of a long chunk
    blocks[11] = blocks[10] + " w1 w2"   # near copy of a long chunk
    for thr in (0.92, 0.7):
        want = o.dedup_keep_indices(blocks, thr)
        got = np.flatnonzero(dedup.dedup_keep_mask(blocks, jaccard_threshold=thr)).tolist()
        assert got == want
        assert 701 not in got and 11 not in got
