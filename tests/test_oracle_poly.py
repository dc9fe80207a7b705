"""CPU: pins the oracle for hot path (b) (oracle/poly.c = restatement of halo2_proofs' best_fft /
best_multiexp / EvaluationDomain wrappers; oracle/bn254.py = the mathematical definitions).

halo2_proofs is a third-party crate that is not vendored under /root/reference and the reference
holds no golden vectors at this boundary (no test calls create_proof: SURVEY.md 8c), so these
results are pinned by algorithm-independent definitions that have canonical encodings:
the O(n^2) DFT, double-and-add MSM and the closed-form KZG identity commit(f) = f(s) * G.
"""
import numpy as np
import pytest

import bn254 as O
import cpu_oracle as CO
from util import fr_to_np, g1_to_np, np_to_fr, np_to_g1, random_fr_np


def test_constants_survey_8b():
    assert O.R_MOD == 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
    assert O.Q_MOD == 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
    assert O.ROOT_OF_UNITY == 0x03ddb9f5166d18b798865ea93dd31f743215cf6dd39329c8d34f1ed960c37c9c
    assert pow(O.ROOT_OF_UNITY, 1 << 28, O.R_MOD) == 1 and pow(O.ROOT_OF_UNITY, 1 << 27, O.R_MOD) != 1
    assert pow(O.ZETA, 3, O.R_MOD) == 1 and O.ZETA != 1
    assert O.DELTA == 0x09226b6e22c6f0ca64ec26aad4c86e715b5f898e5e963f25870e56bbe533e9a2
    assert O.to_mont(1, O.R_MOD) == 0x0e0a77c19a07df2f666ea36f7879462e36fc76959f60cd29ac96341c4ffffffb
    assert O.g1_is_on_curve(O.G1_GEN) and O.g1_mul(O.G1_GEN, O.R_MOD) is None


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 4, 6])
def test_best_fft_is_the_dft(log_n):
    a = O.fr_stream(log_n, 1 << log_n)
    w = O.omega_for(log_n)
    got = np_to_fr(CO.best_fft(fr_to_np(a), fr_to_np([w])[0], log_n, threads=3))
    assert got == O.dft_naive(a, w) == O.ntt(a, w)


@pytest.mark.parametrize("log_n,threads", [(10, 1), (10, 4), (12, 8), (11, 3)])
def test_best_fft_parallel_split(log_n, threads):
    """best_fft switches to the recursive thread-split variant when log_n > log_threads"""
    a = O.fr_stream(100 + log_n, 1 << log_n)
    w = pow(O.omega_for(log_n), 3, O.R_MOD)
    assert np_to_fr(CO.best_fft(fr_to_np(a), fr_to_np([w])[0], log_n, threads=threads)) == O.ntt(a, w)


@pytest.mark.parametrize("k,ext_k", [(3, 5), (8, 10), (9, 9)])
def test_domain_wrappers(k, ext_k):
    co = O.fr_stream(11 + k, 1 << k)
    ev = O.ntt(co, O.omega_for(k))
    assert np_to_fr(CO.lagrange_to_coeff(fr_to_np(ev), k)) == co
    ext = CO.coeff_to_extended(fr_to_np(co), k, ext_k)
    assert np_to_fr(ext) == O.coeff_to_extended(co, k, ext_k)
    # coset evaluation means: value j is f(ZETA * w_ext^j)
    wext = O.omega_for(ext_k)
    for j in (0, 1, (1 << ext_k) - 1):
        assert np_to_fr(ext[j:j + 1])[0] == O.eval_poly(co, O.ZETA * pow(wext, j, O.R_MOD) % O.R_MOD)
    back = np_to_fr(CO.extended_to_coeff(ext, ext_k))
    assert back[:1 << k] == co and not any(back[1 << k:])


@pytest.mark.parametrize("n", [1, 2, 5, 33, 200])
def test_best_multiexp_vs_double_and_add(n):
    pts = O.g1_multiples(n)
    sc = O.fr_stream(0xABC + n, n)
    got = np_to_g1(CO.best_multiexp(fr_to_np(sc), g1_to_np(pts), threads=4))[0]
    assert got == O.msm_naive(sc, pts)


def test_best_multiexp_edge_scalars():
    pts = O.g1_multiples(8)
    for sc in ([0] * 8, [1] * 8, [O.R_MOD - 1] * 8, [3, O.R_MOD - 3, 0, 0, 0, 0, 0, 0]):
        assert np_to_g1(CO.best_multiexp(fr_to_np(sc), g1_to_np(pts)))[0] == O.msm_naive(sc, pts)


def test_kzg_closed_form_2_12():
    """commit(f) over bases s^i G equals f(s) G, 2^12 terms, all host threads"""
    k = 12
    s = O.srs_secret(k)
    pw, acc = [], 1
    for _ in range(1 << k):
        pw.append(acc)
        acc = acc * s % O.R_MOD
    bases = CO.g1_scalar_muls(fr_to_np(pw))
    assert np_to_g1(bases[:3]) == [O.G1_GEN, O.g1_mul(O.G1_GEN, s), O.g1_mul(O.G1_GEN, s * s % O.R_MOD)]
    co = O.fr_stream(5, 1 << k)
    got = np_to_g1(CO.best_multiexp(fr_to_np(co), bases))[0]
    assert got == O.g1_mul(O.G1_GEN, O.eval_poly(co, s))


def test_g1_multiples_c_vs_python():
    assert np_to_g1(CO.g1_multiples(50)) == O.g1_multiples(50)


def test_full_size_fft_roundtrip_2_17():
    k = 17
    a = random_fr_np(1 << k, 1)
    w = fr_to_np([O.omega_for(k)])[0]
    fa = CO.best_fft(a, w, k)
    assert np.array_equal(CO.lagrange_to_coeff(fa, k), a)
