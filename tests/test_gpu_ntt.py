"""GPU parity: NTT kernels (through the C ABI) against the oracle (oracle/bn254.py).

Mirrors how halo2's EvaluationDomain uses best_fft: forward FFT with
an arbitrary root, lagrange_to_coeff, coeff_to_extended, extended_to_coeff.  Bit-exact:
the outputs are compared as canonical integers / raw limbs.
"""
import numpy as np
import pytest

import bn254 as O
from util import fr_to_np, np_to_fr, random_fr_np

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 8, 9, 10, 11, 12])
def test_ntt_matches_oracle(ctx, log_n):
    n = 1 << log_n
    a = O.fr_stream(0x5EED + log_n, n)
    w = O.omega_for(log_n)
    got = np_to_fr(ctx.ntt(fr_to_np(a), fr_to_np([w])[0], log_n))
    want = O.ntt(a, w) if n > 64 else O.dft_naive(a, w)
    assert got == want


def test_ntt_arbitrary_root(ctx):
    # best_fft takes any primitive root, not only the domain's: use omega^5 (still primitive)
    log_n = 9
    a = O.fr_stream(77, 1 << log_n)
    w = pow(O.omega_for(log_n), 5, O.R_MOD)
    got = np_to_fr(ctx.ntt(fr_to_np(a), fr_to_np([w])[0], log_n))
    assert got == O.ntt(a, w)


@pytest.mark.parametrize("k", [1, 4, 10, 12])
def test_lagrange_to_coeff(ctx, k):
    ev = O.fr_stream(3 + k, 1 << k)
    got = np_to_fr(ctx.intt(fr_to_np(ev), k))
    assert got == O.intt(ev, O.omega_for(k))


@pytest.mark.parametrize("k,ext_k", [(3, 5), (8, 10), (10, 12), (10, 10), (9, 11)])
def test_coeff_to_extended_and_back(ctx, k, ext_k):
    co = O.fr_stream(11 + k, 1 << k)
    ext = ctx.coset_ntt(fr_to_np(co), k, ext_k)
    assert np_to_fr(ext) == O.coeff_to_extended(co, k, ext_k)
    back = np_to_fr(ctx.coset_intt(ext, ext_k))
    assert back == O.extended_to_coeff(O.coeff_to_extended(co, k, ext_k), ext_k)
    assert back[: 1 << k] == co and all(x == 0 for x in back[1 << k:])


def test_edge_values(ctx):
    # zeros, ones, r-1 everywhere
    k = 6
    for vals in ([0] * 64, [1] * 64, [O.R_MOD - 1] * 64, [1] + [0] * 63):
        got = np_to_fr(ctx.ntt(fr_to_np(vals), fr_to_np([O.omega_for(k)])[0], k))
        assert got == O.ntt(vals, O.omega_for(k))


@pytest.mark.parametrize("log_n", [17, 19, 22])
def test_full_size_roundtrip_and_linearity(ctx, log_n):
    """BASELINE sizes (2^17 columns, 2^19 extended domain, 2^22 microbench): properties that
    do not need an O(n log n) Python oracle - iFFT(FFT(a)) == a bit-for-bit, and
    FFT(a)[0] == sum(a), FFT(a + b) == FFT(a) + FFT(b) on sampled outputs."""
    n = 1 << log_n
    a = random_fr_np(n, 1000 + log_n)
    w = fr_to_np([O.omega_for(log_n)])[0]
    fa = ctx.ntt(a, w, log_n)
    back = ctx.intt(fa, log_n)
    assert np.array_equal(back, a)
    # out[0] = sum of inputs; out[n/2] = alternating sum
    ai = None
    idx = [0, n // 2]
    got = np_to_fr(fa[idx])
    step = max(1, n >> 16)  # exact sums over all n in Python would be slow: use a structured input instead
    del ai, step
    rinv = pow(O.MONT_R, -1, O.R_MOD)
    lim = a.astype(object)
    vals = (lim[:, 0] + (lim[:, 1] << 64) + (lim[:, 2] << 128) + (lim[:, 3] << 192))
    s_all = int(vals.sum()) * rinv % O.R_MOD
    s_alt = int(vals[0::2].sum() - vals[1::2].sum()) * rinv % O.R_MOD
    assert got == [s_all, s_alt]


def test_batch_dev_matches_single(ctx):
    import torch
    k, batch = 10, 5
    a = random_fr_np(batch << k, 5)
    t = torch.from_numpy(a.view(np.int64)).cuda()
    w = fr_to_np([O.omega_for(k)])[0]
    ctx.ntt_batch_dev(t.data_ptr(), batch, w, k)
    ctx.sync()
    got = t.cpu().numpy().view(np.uint64).reshape(batch, 1 << k, 4)
    for b in range(batch):
        assert np.array_equal(got[b], ctx.ntt(a.reshape(batch, 1 << k, 4)[b], w, k))


@pytest.mark.parametrize("log_n", [17, 19, 22])
def test_full_size_full_array_vs_cpu_oracle(ctx, log_n):
    """BASELINE sizes, every output element: best_fft against the CPU restatement (oracle/poly.c: bit-reversal +
    radix-2 DIT, a different algorithm from the product's Stockham passes), and at 2^17 -> 2^19 the three
    EvaluationDomain wrappers create_proof uses"""
    import cpu_oracle as CO
    n = 1 << log_n
    a = random_fr_np(n, 2000 + log_n)
    w = fr_to_np([O.omega_for(log_n)])[0]
    assert np.array_equal(ctx.ntt(a, w, log_n), CO.best_fft(a, w, log_n))
    if log_n == 17:
        co = ctx.intt(a, log_n)
        assert np.array_equal(co, CO.lagrange_to_coeff(a, log_n))
        ext = ctx.coset_ntt(co, log_n, log_n + 2)
        assert np.array_equal(ext, CO.coeff_to_extended(co, log_n, log_n + 2))
        assert np.array_equal(ctx.coset_intt(ext, log_n + 2), CO.extended_to_coeff(ext, log_n + 2))


@pytest.mark.parametrize("log_n,batch", [(13, 3), (17, 2), (19, 1), (22, 1)])
def test_tma_tile_loader_matches_plain_loads(ctx, monkeypatch, log_n, batch):
    """the opt-in pass kernel whose tiles are fetched by cp.async.bulk.tensor + mbarrier (ntt.cu k_ntt_pass_tma,
    B2R_NTT_TMA=1): same outputs, bit for bit, as the default kernel - forward, inverse, and the zero-padded coset
    extension (rows beyond the coefficients are zero-filled by the copy engine)"""
    import torch
    n = 1 << log_n
    a = random_fr_np(n * batch, 3000 + log_n)
    w = fr_to_np([O.omega_for(log_n)])[0]

    def run():
        t = torch.from_numpy(a.view(np.int64)).cuda()
        ctx.ntt_batch_dev(t.data_ptr(), batch, w, log_n)
        ctx.sync()
        fwd = t.cpu().numpy().copy()
        ctx.intt_batch_dev(t.data_ptr(), batch, log_n)
        ctx.sync()
        inv = t.cpu().numpy().copy()
        ext = None
        if log_n <= 19:
            e = torch.zeros(batch * 4 * n * 4, dtype=torch.int64, device="cuda")
            ctx.coset_ntt_batch_dev(t.data_ptr(), batch, log_n, log_n + 2, e.data_ptr())
            ctx.sync()
            ext = e.cpu().numpy().copy()
        return fwd, inv, ext

    monkeypatch.setenv("B2R_NTT_TMA", "0")
    base = run()
    monkeypatch.setenv("B2R_NTT_TMA", "1")
    tma = run()
    for x, y in zip(base, tma):
        assert (x is None and y is None) or np.array_equal(x, y)
    # the other opt-in variant: twiddles of the passes behind the first staged in shared memory (B2R_NTT_TWS=1)
    monkeypatch.setenv("B2R_NTT_TMA", "0")
    monkeypatch.setenv("B2R_NTT_TWS", "1")
    tws = run()
    monkeypatch.delenv("B2R_NTT_TWS")
    for x, y in zip(base, tws):
        assert (x is None and y is None) or np.array_equal(x, y)
    assert np.array_equal(base[1].view(np.uint64).reshape(-1, 4), a)     # the inverse undoes the forward transform
