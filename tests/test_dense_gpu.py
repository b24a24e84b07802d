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


# ---- batched path: tcgen05/TMA GEMM with the top-k epilogue (cmr_dense_topk_ex, CMR_DENSE_MMA) ----

def _check_mma(bits, qbits, k, row_offset=0, also_scan=True):
    from classmate_rag_b200 import ops
    emb, q = _to_dev(bits), _to_dev(qbits)
    s, i, c, f = [t.clone() for t in ops.dense_topk(emb, q, k, row_offset=row_offset, algo="mma")]
    torch.cuda.synchronize()
    scores, ids, counts, flags = s.cpu().numpy(), i.cpu().numpy(), c.cpu().numpy(), f.cpu().numpy()
    for b in range(qbits.shape[0]):
        want_ids, want_sc = o.dense_topk(qbits[b], bits, k, row_offset=row_offset)
        n = len(want_ids)
        assert counts[b] == n, (b, counts[b], n)
        assert ids[b, :n].tolist() == want_ids.tolist(), b
        assert scores[b, :n].tobytes() == want_sc.tobytes(), b
        assert (ids[b, n:] == -1).all()
        assert flags[b] == 0
    if also_scan:  # the two kernels must agree bit for bit
        s2, i2, c2, f2 = ops.dense_topk(emb, q, k, row_offset=row_offset, algo="scan")
        torch.cuda.synchronize()
        assert torch.equal(i, i2) and torch.equal(c, c2)
        assert s.cpu().numpy().tobytes() == s2.cpu().numpy().tobytes()


@pytest.mark.parametrize("n,d,k,b", [(20000, 768, 10, 32), (5000, 1024, 24, 7), (3000, 384, 10, 128),
                                      (4096, 64, 10, 33), (2600, 136, 3, 16), (1500, 2048, 10, 3),
                                      (1111, 1536, 5, 40), (256, 64, 10, 1), (300, 72, 40, 9)])
def test_dense_mma_matches_oracle(n, d, k, b):
    rng = np.random.default_rng(n + d + b)
    bits = _corpus(rng, n, d, dup_every=100)
    _check_mma(bits, _queries(rng, bits, b), k)


@pytest.mark.parametrize("k", [1, 24, 25, 56, 57, 120])
def test_dense_mma_k_ranges(k):
    rng = np.random.default_rng(100 + k)
    bits = _corpus(rng, 6000, 128, dup_every=50)
    _check_mma(bits, _queries(rng, bits, 12), k, row_offset=5_000_000_000)


def test_dense_mma_many_queries_multiple_blocks():
    """More than 128 queries: several query blocks per row tile, last block ragged."""
    rng = np.random.default_rng(77)
    bits = _corpus(rng, 9000, 256, dup_every=64)
    _check_mma(bits, _queries(rng, bits, 300), 10)


def test_dense_mma_repeated_runs_equal_exhaustive_scan():
    """Race check: the batched path, run repeatedly on shapes with several query blocks, must
    give the exhaustive float64 scan's answer every time (a missing barrier in the candidate
    finalize once dropped candidates in about 1 run in 25 -- tools/stress_dense.py)."""
    from classmate_rag_b200 import ops
    shapes = [(9000, 256, 300, 10), (9000, 768, 520, 24), (40000, 128, 300, 10)]
    for it in range(9):
        n, d, b, k = shapes[it % len(shapes)]
        g = torch.Generator(device="cuda").manual_seed(1000 + it)
        x = torch.nn.functional.normalize(torch.randn((n, d), device="cuda", generator=g), dim=1).to(torch.bfloat16)
        x[63::64] = x[31::64][: x[63::64].shape[0]]          # exact duplicates: ties at every rank
        rows = torch.randint(0, n, (b,), device="cuda", generator=g)
        q = x[rows].float() + 0.5 * torch.randn((b, d), device="cuda", generator=g) / d ** 0.5
        q = torch.nn.functional.normalize(q, dim=1).to(torch.bfloat16)
        ref = [t.clone() for t in ops.dense_topk(x, q, k, algo="exact")]
        for _ in range(8):
            out = [t.clone() for t in ops.dense_topk(x, q, k, algo="mma")]
            torch.cuda.synchronize()
            ok = out[3] == 0                      # an uncertified query says so and is re-run by the caller
            assert int((~ok).sum()) <= b // 100
            assert torch.equal(out[1][ok], ref[1][ok]) and torch.equal(out[2][ok], ref[2][ok])
            assert out[0][ok].cpu().numpy().tobytes() == ref[0][ok].cpu().numpy().tobytes()


def test_dense_mma_sampled_bound_large_matrix():
    """Enough tiles that the sample pass really strides (every 16th tile) and uses
    whole-tile maxima; oracle on a subset of the queries, scan kernel on all."""
    from classmate_rag_b200 import ops
    rng = np.random.default_rng(78)
    n, d, b, k = 300_000, 128, 64, 10
    bits = _corpus(rng, n, d, dup_every=1000)
    qbits = _queries(rng, bits, b)
    emb, q = _to_dev(bits), _to_dev(qbits)
    s, i, c, f = [t.clone() for t in ops.dense_topk(emb, q, k, algo="mma")]
    s2, i2, c2, f2 = [t.clone() for t in ops.dense_topk(emb, q, k, algo="scan")]
    torch.cuda.synchronize()
    assert torch.equal(i, i2) and torch.equal(c, c2) and torch.equal(f, f2)
    assert s.cpu().numpy().tobytes() == s2.cpu().numpy().tobytes()
    for bb in range(4):
        want_ids, want_sc = o.dense_topk(qbits[bb], bits, k)
        assert i[bb].cpu().numpy().tolist() == want_ids.tolist()
        assert s[bb].cpu().numpy().tobytes() == want_sc.tobytes()


def test_dense_mma_adversarial_order_flags_or_exact():
    """Rows sorted by similarity to the query (every later tile beats the sampled ones):
    the candidate buffer may overflow; the call must then say so, never return a wrong
    certified answer."""
    from classmate_rag_b200 import ops
    rng = np.random.default_rng(79)
    n, d, k = 40_000, 64, 10
    bits = _corpus(rng, n, d)
    qbits = _queries(rng, bits, 9)
    sc = o.exact_dots(qbits[0], bits)
    bits = bits[np.argsort(sc, kind="stable")]          # ascending similarity to query 0
    s, i, c, f = [t.cpu().numpy() for t in ops.dense_topk(_to_dev(bits), _to_dev(qbits), k, algo="mma")]
    for bb in range(qbits.shape[0]):
        if f[bb] == 0:
            want_ids, want_sc = o.dense_topk(qbits[bb], bits, k)
            assert i[bb].tolist() == want_ids.tolist() and s[bb].tobytes() == want_sc.tobytes()


def test_dense_mma_all_duplicates():
    rng = np.random.default_rng(3)
    bits = np.repeat(_corpus(rng, 1, 128), 5000, axis=0)
    from classmate_rag_b200 import ops
    s, i, c, f = ops.dense_topk(_to_dev(bits), _to_dev(np.repeat(bits[:1], 9, axis=0)), 10, algo="mma")
    torch.cuda.synchronize()
    # 5000 identical scores: far more candidates than slots -> flagged, as documented
    assert (f.cpu().numpy() == 1).all()


def test_dense_mma_unsupported_shapes_raise():
    from classmate_rag_b200 import ops
    emb = torch.zeros((100, 64), dtype=torch.bfloat16, device="cuda")
    q = torch.zeros((9, 64), dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError):
        ops.dense_topk(emb, q, 4, algo="mma")          # fewer than 256 rows
    emb = torch.zeros((1000, 32), dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError):
        ops.dense_topk(emb, torch.zeros((9, 32), dtype=torch.bfloat16, device="cuda"), 4, algo="mma")   # dim < 64
    s, i, c, f = ops.dense_topk(emb, torch.zeros((9, 32), dtype=torch.bfloat16, device="cuda"), 4)      # auto -> scan
    torch.cuda.synchronize()
    assert (c.cpu().numpy() == 4).all()


@pytest.mark.parametrize("n,d,b,keep", [(9000, 256, 12, 0.3), (20000, 768, 32, 0.9), (5000, 64, 130, 0.02), (70000, 64, 9, 0.5)])
def test_dense_mma_row_mask(n, d, b, keep):
    """`where` filter / tombstones on the tcgen05 path: the mask becomes one bit per row and is
    applied in the epilogue (sample pass included, so the bound only sees allowed rows)."""
    from classmate_rag_b200 import ops
    rng = np.random.default_rng(n + b)
    bits = _corpus(rng, n, d, dup_every=64)
    qbits = _queries(rng, bits, b)
    mask = (rng.random(n) < keep).astype(np.uint8)
    mask[-7:] = 1
    emb, q, m = _to_dev(bits), _to_dev(qbits), torch.from_numpy(mask).cuda()
    s, i, c, f = [t.clone() for t in ops.dense_topk(emb, q, 10, row_mask=m, row_offset=1000, algo="mma")]
    s2, i2, c2, f2 = ops.dense_topk(emb, q, 10, row_mask=m, row_offset=1000, algo="scan")
    torch.cuda.synchronize()
    ok = f.cpu().numpy() == 0          # a very selective mask may overflow a list: flagged, never wrong
    assert ok.sum() >= b // 2
    assert torch.equal(i[ok], i2[ok]) and s[ok].cpu().numpy().tobytes() == s2[ok].cpu().numpy().tobytes()
    for bb in range(min(b, 3)):
        want_ids, want_sc = o.dense_topk(qbits[bb], bits, 10, mask=mask, row_offset=1000)
        if ok[bb]:
            assert i[bb, :len(want_ids)].cpu().numpy().tolist() == want_ids.tolist()
            assert s[bb, :len(want_ids)].cpu().numpy().tobytes() == want_sc.tobytes()
    # the certified wrapper always ends exact
    s3, i3, c3, f3 = ops.dense_topk_certified(emb, q, 10, row_mask=m, row_offset=1000)
    torch.cuda.synchronize()
    assert int(f3.sum()) == 0 and torch.equal(i3, i2)
    # an unaligned mask view works too
    big = torch.zeros(n + 3, dtype=torch.uint8, device="cuda")
    big[3:] = m
    s4, i4, c4, f4 = ops.dense_topk(emb, q, 10, row_mask=big[3:], row_offset=1000, algo="mma")
    torch.cuda.synchronize()
    assert torch.equal(i4[ok], i2[ok])


# ---- exhaustive exact scan (CMR_DENSE_EXACT) and the certified wrapper ----

@pytest.mark.parametrize("n,d,k,b", [(20000, 768, 10, 3), (3000, 384, 24, 2), (2500, 8, 3, 2), (700, 64, 100, 1), (5, 64, 10, 2)])
def test_dense_exact_scan_matches_oracle(n, d, k, b):
    from classmate_rag_b200 import ops
    rng = np.random.default_rng(n * 3 + d)
    bits = _corpus(rng, n, d, dup_every=100 if n > 200 else 0)
    qbits = _queries(rng, bits, b)
    mask = (rng.random(n) < 0.5).astype(np.uint8) if n > 1000 else None
    m = None if mask is None else torch.from_numpy(mask).cuda()
    s, i, c, f = ops.dense_topk(_to_dev(bits), _to_dev(qbits), k, row_mask=m, row_offset=77, algo="exact")
    torch.cuda.synchronize()
    s, i, c, f = s.cpu().numpy(), i.cpu().numpy(), c.cpu().numpy(), f.cpu().numpy()
    for bb in range(b):
        want_ids, want_sc = o.dense_topk(qbits[bb], bits, k, mask=mask, row_offset=77)
        nn = len(want_ids)
        assert c[bb] == nn and f[bb] == 0
        assert i[bb, :nn].tolist() == want_ids.tolist() and s[bb, :nn].tobytes() == want_sc.tobytes()
        assert (i[bb, nn:] == -1).all()


def test_certified_wrapper_resolves_flagged_queries():
    """5000 identical rows: both fast paths must flag; the wrapper re-runs the query on the
    exhaustive scan and returns the certified answer (ties in ascending row order)."""
    from classmate_rag_b200 import ops
    rng = np.random.default_rng(3)
    bits = np.repeat(_corpus(rng, 1, 128), 5000, axis=0)
    other = _corpus(rng, 3000, 128)
    allb = np.concatenate([other[:1500], bits, other[1500:]])
    q = np.concatenate([bits[:1], _queries(rng, other, 2)])          # query 0 hits the duplicates
    for nq in (3, 12):                                               # scan path / tcgen05 path
        qq = np.concatenate([q] * (nq // 3))
        s, i, c, f = ops.dense_topk_certified(_to_dev(allb), _to_dev(qq), 10)
        torch.cuda.synchronize()
        assert int(f.sum()) == 0
        for bb in range(qq.shape[0]):
            want_ids, want_sc = o.dense_topk(qq[bb], allb, 10)
            assert i[bb].cpu().numpy().tolist() == want_ids.tolist()
            assert s[bb].cpu().numpy().tobytes() == want_sc.tobytes()
        assert i[0].cpu().numpy().tolist() == list(range(1500, 1510))


def test_dense_mma_many_groups_block_bound_kernel():
    """k = 100 (KP = 128) over 2.5M rows: 610 sampled tiles x 8 groups = 4880 group maxima,
    more than one warp stages -> the block-wide bound kernel.  Checked against the scan path
    (and the oracle on one query)."""
    from classmate_rag_b200 import ops
    n, d, k, b = 2_500_000, 64, 100, 9
    g = torch.Generator(device="cuda").manual_seed(11)
    emb = torch.nn.functional.normalize(torch.randn((n, d), generator=g, device="cuda"), dim=1).to(torch.bfloat16)
    q = torch.nn.functional.normalize(emb[:b].float() + 0.3 * torch.randn((b, d), generator=g, device="cuda") / d ** 0.5,
                                      dim=1).to(torch.bfloat16)
    s, i, c, f = [t.clone() for t in ops.dense_topk(emb, q, k, algo="mma")]
    s2, i2, c2, f2 = ops.dense_topk(emb, q, k, algo="scan")
    torch.cuda.synchronize()
    assert int(f.sum()) == 0 and torch.equal(i, i2) and s.cpu().numpy().tobytes() == s2.cpu().numpy().tobytes()
    bits = emb.view(torch.int16).cpu().numpy().view(np.uint16)
    want_ids, want_sc = o.dense_topk(q[0].view(torch.int16).cpu().numpy().view(np.uint16), bits, k)
    assert i[0].cpu().numpy().tolist() == want_ids.tolist() and s[0].cpu().numpy().tobytes() == want_sc.tobytes()


def test_dense_mma_clustered_rows_do_not_overflow_the_cta_lists():
    """200 near-copies of the query's source row sit next to each other (chunks of one document):
    they all land in one or two row tiles, i.e. in one or two CTAs' private lists."""
    from classmate_rag_b200 import ops
    rng = np.random.default_rng(21)
    n, d = 60_000, 128
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    base = x[31_000].copy()
    for i in range(200):
        v = base + 0.05 * rng.standard_normal(d).astype(np.float32) / np.sqrt(d) * (1 + i / 50)
        x[31_000 + i] = v / np.linalg.norm(v)
    bits = o.f32_to_bf16_bits(x)
    qbits = np.repeat(o.f32_to_bf16_bits(base[None]), 9, axis=0)
    s, i, c, f = ops.dense_topk(_to_dev(bits), _to_dev(qbits), 24, algo="mma")
    torch.cuda.synchronize()
    assert int(f.sum()) == 0
    want_ids, want_sc = o.dense_topk(qbits[0], bits, 24)
    assert i[0].cpu().numpy().tolist() == want_ids.tolist() and s[0].cpu().numpy().tobytes() == want_sc.tobytes()
