"""GPU parity: near-duplicate filter (tcgen05 pair scan + exact rescoring + greedy resolve)
vs the oracle's sequential keep-first rule."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as o

pytestmark = pytest.mark.gpu


def _to_dev(bits):
    return torch.from_numpy(bits.view(np.int16)).cuda().view(torch.bfloat16)


def _corpus(seed, n, d, near=0.05, exact=0.01, noise=0.05):
    """unit rows; a share are noisy copies of EARLIER rows (cos ~ 0.99+) or exact copies;
    some copies are copies of copies (chains), and some noise levels straddle 0.95."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    for i in range(1, n):
        u = rng.random()
        if u < near:
            j = int(rng.integers(0, i))
            s = noise * (1 + 8 * (rng.random() < 0.3))      # 30 %: heavier noise, cos near the threshold
            v = x[j] + s * rng.standard_normal(d).astype(np.float32) / np.sqrt(d) * np.sqrt(d) * 0.2
            x[i] = v / np.linalg.norm(v)
        elif u < near + exact:
            x[i] = x[int(rng.integers(0, i))]
    return o.f32_to_bf16_bits(x)


@pytest.mark.parametrize("n,d,thr", [(3000, 64, 0.95), (1500, 128, 0.95), (700, 768, 0.95), (2100, 72, 0.9), (257, 64, 0.99)])
def test_neardup_matches_oracle(n, d, thr):
    from classmate_rag_b200 import neardup
    bits = _corpus(n + d, n, d)
    want = o.neardup_keep_mask(bits, thr)
    got = neardup.neardup_keep_mask(_to_dev(bits), thr).cpu().numpy().astype(bool)
    assert 0 < (~want).sum() < n, "corpus should contain duplicates"
    assert np.array_equal(got, want)


def test_neardup_block_shares_partition_the_triangle():
    """Ranks emulated one after the other on one GPU: the union of the shares' edges is the
    single-rank edge set, for any number of ranks, and resolves to the same mask."""
    from classmate_rag_b200 import neardup
    bits = _corpus(5, 2500, 64)
    emb = _to_dev(bits)
    base = torch.sort(neardup.neardup_edges(emb, 0.95)).values
    want = o.neardup_keep_mask(bits, 0.95)
    for world in (2, 3, 8):
        parts = [neardup.neardup_edges(emb, 0.95, rank=r, world=world) for r in range(world)]
        allp = torch.sort(torch.cat(parts)).values
        assert torch.equal(allp, base)
        assert np.array_equal(neardup.resolve(torch.cat(parts), 2500).cpu().numpy().astype(bool), want)


def test_neardup_edges_are_exact_pairs():
    from classmate_rag_b200 import neardup
    bits = _corpus(9, 1200, 64)
    edges = neardup.neardup_edges(_to_dev(bits), 0.95).cpu().numpy()
    got = {(int(e >> 32), int(e & 0xFFFFFFFF)) for e in edges}
    want = set()
    for i in range(1, 1200):
        s = o.exact_dots(bits[i], bits[:i])
        for j in np.nonzero(s >= 0.95)[0]:
            want.add((i, int(j)))
    assert got == want


def test_neardup_trivial_inputs():
    from classmate_rag_b200 import neardup
    one = _to_dev(_corpus(1, 1, 64))
    assert neardup.neardup_keep_mask(one).cpu().tolist() == [1]
    same = _to_dev(np.repeat(_corpus(2, 1, 64), 300, axis=0))
    assert neardup.neardup_keep_mask(same).cpu().tolist() == [1] + [0] * 299
    with pytest.raises(RuntimeError):
        neardup.neardup_keep_mask(one.cpu())
