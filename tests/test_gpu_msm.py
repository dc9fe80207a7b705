"""GPU parity: Pippenger MSM (C ABI) against the oracle's mathematical definition
(sum_i s_i * P_i, oracle/bn254.py) - SURVEY.md 8 row a13 (best_multiexp).  Affine output is
canonical, so equality is bit-exact."""
import numpy as np
import pytest

import bn254 as O
from util import fr_to_np, g1_to_np, np_jac_to_g1, np_to_g1

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bases_4096(ctx):
    pts = O.g1_multiples(4096)
    return pts, ctx.bases_register(g1_to_np(pts))


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 257])
def test_msm_small_vs_naive(ctx, bases_4096, n):
    pts, bs = bases_4096
    sc = O.fr_stream(0x5EED + n, n)
    got = np_jac_to_g1(ctx.msm(bs, fr_to_np(sc)))
    assert got == O.msm_naive(sc, pts[:n])


def test_msm_closed_form_4096(ctx, bases_4096):
    # bases (i+1)G  =>  MSM = (sum s_i (i+1)) G
    pts, bs = bases_4096
    sc = O.fr_stream(99, 4096)
    got = np_jac_to_g1(ctx.msm(bs, fr_to_np(sc)))
    want = O.g1_mul(O.G1_GEN, sum(s * (i + 1) for i, s in enumerate(sc)) % O.R_MOD)
    assert got == want


def test_msm_edge_scalars(ctx, bases_4096):
    pts, bs = bases_4096
    n = 64
    cases = {
        "zeros": [0] * n,
        "ones": [1] * n,                       # every point in the same bucket
        "minus_one": [O.R_MOD - 1] * n,        # top window / signed-digit carry path
        "cancel": [5, O.R_MOD - 5] + [0] * (n - 2),
        "pow2": [1 << (4 * i) for i in range(n)][:n],
        "window_edges": [(1 << 15), (1 << 15) + 1, (1 << 16) - 1, (1 << 16), (1 << 31), (1 << 253)] + [0] * (n - 6),
    }
    for name, sc in cases.items():
        sc = [s % O.R_MOD for s in sc]
        got = np_jac_to_g1(ctx.msm(bs, fr_to_np(sc)))
        assert got == O.msm_naive(sc, pts[:n]), name


def test_msm_repeated_points_doubling_path(ctx):
    # identical bases force the P+P and P-P exceptional cases inside a bucket
    g = O.G1_GEN
    pts = [g, g, g, O.g1_neg(g), None, O.g1_mul(g, 65536), O.g1_mul(g, 65536)]
    bs = ctx.bases_register(g1_to_np(pts))
    for sc in ([7, 7, 7, 7, 3, 1, 1], [1, 1, 0, 2, 9, 0, 0], [65536, 0, 0, 0, 0, 1, 0], [3, 0, 0, 3, 0, 0, 0]):
        got = np_jac_to_g1(ctx.msm(bs, fr_to_np(sc)))
        assert got == O.msm_naive(sc, pts)
    bs.free()


def test_msm_batch_skewed_like_advice(ctx, bases_4096):
    """advice-column-like scalars: mostly 0/1, bytes, 64-bit limbs, a few full-size"""
    pts, bs = bases_4096
    n, m = 4096, 3
    rows = []
    for v in range(m):
        raw = O.fr_stream(200 + v, n)
        col = []
        for i, x in enumerate(raw):
            sel = x % 10
            col.append(0 if sel < 4 else 1 if sel < 6 else (x >> 20) % 256 if sel < 8 else (x >> 30) % (1 << 64) if sel < 9 else x)
        rows.append(col)
    arr = np.stack([fr_to_np(c) for c in rows])
    got = np_to_g1(ctx.msm_batch(bs, arr))
    for v in range(m):
        want = O.g1_mul(O.G1_GEN, sum(s * (i + 1) for i, s in enumerate(rows[v])) % O.R_MOD)
        assert got[v] == want


def test_msm_2p16_closed_form(ctx):
    n = 1 << 16
    pts = O.g1_multiples(n)
    bs = ctx.bases_register(g1_to_np(pts))
    sc = O.fr_stream(0x5EED, n)
    got = np_jac_to_g1(ctx.msm(bs, fr_to_np(sc)))
    want = O.g1_mul(O.G1_GEN, sum(s * (i + 1) for i, s in enumerate(sc)) % O.R_MOD)
    assert got == want
    bs.free()


@pytest.mark.parametrize("log_n", [17, 18])
def test_msm_binned_sort_path(ctx, log_n, monkeypatch):
    """the two-pass binned counting sort (msm.cu k_bin_scatter / k_bin_sort: 128 bins up to 2^17 scalars, 256 bins at
    2^18) forced through the generic entry point: uniform vectors against the closed form, and a vector of one repeated
    scalar, which overflows a bin and must come back through the general kernels with the same answer"""
    monkeypatch.setenv("B2R_MSM_BINSORT", "1")
    n = 1 << log_n
    pts = O.g1_multiples(n)
    bs = ctx.bases_register(g1_to_np(pts))
    tri = n * (n + 1) // 2
    vecs = [O.fr_stream(0xB1 + log_n, n), O.fr_stream(0xB2 + log_n, n), [0x1234567 * 0x10001] * n]
    got = np_to_g1(ctx.msm_batch(bs, np.stack([fr_to_np(v) for v in vecs])))
    for j, v in enumerate(vecs):
        k = (v[0] * tri if j == 2 else sum(s * (i + 1) for i, s in enumerate(v))) % O.R_MOD
        assert got[j] == O.g1_mul(O.G1_GEN, k), f"vector {j}"
    bs.free()


def test_msm_uniform_hint_same_result(ctx):
    """b2r_msm_g1_batch_dev_ex with B2R_MSM_UNIFORM: identical commitments for uniform vectors (binned sort) and for a
    skewed vector that overflows a bin (detected, general path)"""
    import torch
    n = 1 << 15       # > 2^14 so that the table has 16-bit windows (the binned sort's width)
    pts = O.g1_multiples(n)
    bs = ctx.bases_register(g1_to_np(pts))
    vecs = [O.fr_stream(0xC1, n), [(7 if i % 3 else 0) for i in range(n)]]
    arr = torch.from_numpy(np.stack([fr_to_np(v) for v in vecs]).view(np.int64)).cuda()
    out_a = torch.zeros(2 * 8, dtype=torch.int64, device="cuda")
    out_b = torch.zeros(2 * 8, dtype=torch.int64, device="cuda")
    ctx.msm_batch_dev(bs, arr.data_ptr(), 2, n, out_a.data_ptr())
    ctx.msm_batch_dev(bs, arr.data_ptr(), 2, n, out_b.data_ptr(), uniform=True)
    torch.cuda.synchronize()
    assert torch.equal(out_a, out_b)
    got = np_to_g1(out_b.cpu().numpy().view(np.uint64).reshape(2, 8))
    for j, v in enumerate(vecs):
        assert got[j] == O.g1_mul(O.G1_GEN, sum(s * (i + 1) for i, s in enumerate(v)) % O.R_MOD)
    bs.free()


def test_msm_2p20_config5(ctx):
    """BASELINE configs[4] size (standalone MSM 2^20): bases (i+1) G from the C oracle, (a) uniform seeded scalars and
    (b) advice-like skew (zeros, bits, bytes, 64-bit limbs, a few full-size values), each against the closed form
    (sum s_i (i+1)) G AND against the CPU best_multiexp restatement (oracle/poly.c), with and without the uniform hint"""
    import torch
    import cpu_oracle as CO
    import plonk as PL
    from util import random_fr_np
    n = 1 << 20
    bases = CO.g1_multiples(n)
    bs = ctx.bases_register(bases)
    uni = random_fr_np(n, 0x5EED)
    rng = np.random.default_rng(20)
    sel = rng.integers(0, 10, size=n)
    skew = uni.copy()
    skew[sel < 4] = 0
    one = fr_to_np([1])[0]
    skew[(sel >= 4) & (sel < 6)] = one
    small = PL.ints_to_np([int(x) for x in rng.integers(0, 256, size=n)])
    limb = PL.ints_to_np([int(x) for x in rng.integers(0, 1 << 63, size=n, dtype=np.uint64)])
    m = (sel >= 6) & (sel < 8)
    skew[m] = small[m]
    m = sel == 8
    skew[m] = limb[m]
    arr = np.stack([uni, skew])
    t = torch.from_numpy(arr.view(np.int64)).cuda()
    outs = []
    for hint in (False, True):
        out = torch.zeros(2 * 8, dtype=torch.int64, device="cuda")
        ctx.msm_batch_dev(bs, t.data_ptr(), 2, n, out.data_ptr(), uniform=hint)
        torch.cuda.synchronize()
        outs.append(out.cpu().numpy().view(np.uint64).reshape(2, 8))
    assert np.array_equal(outs[0], outs[1])
    got = np_to_g1(outs[0])
    for j in range(2):
        vals = PL.np_to_ints(arr[j])
        want = O.g1_mul(O.G1_GEN, sum(s * (i + 1) for i, s in enumerate(vals)) % O.R_MOD)
        assert got[j] == want, f"vector {j} vs closed form"
        assert np.array_equal(CO.best_multiexp(arr[j], bases), outs[0][j]), f"vector {j} vs CPU best_multiexp"
    bs.free()


@pytest.mark.parametrize("variant", ["3", "4"])
def test_accumulate_ring_variants_same_result(ctx, monkeypatch, variant):
    """the opt-in accumulation kernels that stage the table points through a shared-memory ring (msm.cu
    k_accum_entries_ring: 3 = per-lane cp.async.bulk + mbarrier, 4 = LDGSTS groups): identical commitments for a dense
    vector, a sparse one and a ragged length"""
    import cpu_oracle as CO
    from util import random_fr_np
    n = (1 << 15) + 77
    bases = CO.g1_multiples(n)
    bs = ctx.bases_register(bases)
    dense = random_fr_np(n, 41)
    sparse = dense.copy()
    sparse[np.random.default_rng(2).random(n) < 0.9] = 0
    arr = np.stack([dense, sparse])
    monkeypatch.setenv("B2R_MSM_VARIANT", "0")
    want = ctx.msm_batch(bs, arr)
    monkeypatch.setenv("B2R_MSM_VARIANT", variant)
    got = ctx.msm_batch(bs, arr)
    assert np.array_equal(got, want)
    assert np.array_equal(CO.best_multiexp(dense, bases), want[0])
    bs.free()
