"""GPU parity: cmr_dense_topk (through the C ABI) vs the exact CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as o

pytestmark = pytest.mark.gpu


def _corpus(rng, n, d, dup_every=0):
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    bits = o.f32_to_bf16_bits(x)
    if dup_every:
        for i in range(dup_every - 1, n, dup_every):
            bits[i] = bits[max(0, i - dup_every // 2)]
    return bits


def _queries(rng, bits, b):
    n, d = bits.shape
    rows = rng.integers(0, n, b)
    base = o.bf16_bits_to_f32(bits[rows])
    q = base + 0.5 * rng.standard_normal((b, d)).astype(np.float32) / np.sqrt(d)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return o.f32_to_bf16_bits(q)


def _to_dev(bits):
    return torch.from_numpy(bits.view(np.int16)).cuda().view(torch.bfloat16)


def _check(bits, qbits, k, mask=None, row_offset=0):
    from classmate_rag_b200 import ops
    emb = _to_dev(bits)
    q = _to_dev(qbits)
    m = None if mask is None else torch.from_numpy(mask).cuda()
    scores, ids, counts, flags = ops.dense_topk(emb, q, k, row_mask=m, row_offset=row_offset)
    torch.cuda.synchronize()
    scores, ids, counts, flags = scores.cpu().numpy(), ids.cpu().numpy(), counts.cpu().numpy(), flags.cpu().numpy()
    for b in range(qbits.shape[0]):
        want_ids, want_sc = o.dense_topk(qbits[b], bits, k, mask=mask, row_offset=row_offset)
        n = len(want_ids)
        assert counts[b] == n
        assert ids[b, :n].tolist() == want_ids.tolist()           # bit-exact ids / ranks
        assert scores[b, :n].tobytes() == want_sc.tobytes()       # bit-exact float64 scores
        assert (ids[b, n:] == -1).all()
        assert flags[b] == 0


@pytest.mark.parametrize("n,d,k", [(20000, 768, 10), (5000, 1024, 24), (3000, 384, 10), (4096, 64, 10),
                                   (2500, 8, 3), (1500, 2048, 10), (1111, 1536, 5)])
def test_dense_topk_matches_oracle(n, d, k):
    rng = np.random.default_rng(n + d)
    bits = _corpus(rng, n, d, dup_every=100)
    _check(bits, _queries(rng, bits, 3), k)


@pytest.mark.parametrize("k", [1, 24, 25, 56, 57, 100, 120])
def test_dense_topk_k_ranges(k):
    rng = np.random.default_rng(k)
    bits = _corpus(rng, 6000, 128, dup_every=50)
    _check(bits, _queries(rng, bits, 2), k)


def test_dense_tiny_and_empty():
    rng = np.random.default_rng(1)
    for n in (1, 5, 31, 32, 33, 40):
        bits = _corpus(rng, n, 64)
        _check(bits, _queries(rng, bits, 2), 10)
    from classmate_rag_b200 import ops
    emb = torch.empty((0, 64), dtype=torch.bfloat16, device="cuda")
    q = torch.zeros((1, 64), dtype=torch.bfloat16, device="cuda")
    s, i, c, f = ops.dense_topk(emb, q, 4)
    torch.cuda.synchronize()
    assert c.item() == 0 and (i.cpu().numpy() == -1).all()


def test_dense_mask_and_offset():
    rng = np.random.default_rng(2)
    bits = _corpus(rng, 9000, 256, dup_every=64)
    mask = (rng.random(9000) < 0.3).astype(np.uint8)
    _check(bits, _queries(rng, bits, 3), 10, mask=mask, row_offset=123456789012)
    mask[:] = 0
    mask[[5, 77]] = 1
    _check(bits, _queries(rng, bits, 1), 10, mask=mask)


def test_dense_all_duplicates_tie_break_is_row_ascending():
    rng = np.random.default_rng(3)
    bits = np.repeat(_corpus(rng, 1, 128), 5000, axis=0)
    from classmate_rag_b200 import ops
    s, i, c, f = ops.dense_topk(_to_dev(bits), _to_dev(bits[:1]), 10)
    torch.cuda.synchronize()
    assert i.cpu().numpy()[0].tolist() == list(range(10))
    # identical fp32 scores: the certificate cannot prove sufficiency and says so
    assert f.item() == 1


def test_bad_arguments_raise():
    from classmate_rag_b200 import ops
    emb = torch.zeros((16, 64), dtype=torch.bfloat16, device="cuda")
    q = torch.zeros((1, 64), dtype=torch.bfloat16, device="cuda")
    with pytest.raises(ValueError):
        ops.dense_topk(emb, q, 0)
    with pytest.raises(ValueError):
        ops.dense_topk(emb, q, 500)
    with pytest.raises(RuntimeError):
        ops.dense_topk(emb.cpu(), q, 4)


def test_f32_to_bf16_kernel_is_rne():
    from classmate_rag_b200 import ops
    x = torch.randn(100003, device="cuda")
    got = ops.f32_to_bf16(x)
    assert torch.equal(got.view(torch.int16), x.to(torch.bfloat16).view(torch.int16))
