"""TEST INFRASTRUCTURE ONLY - CPU oracle for hot path (b): BN254 fields, G1, NTT, MSM, KZG SRS.

This file is a checker.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it.  The product (libb2rsa.so and the
b2rsa package) never does.

What it restates
----------------
The arithmetic lives in third-party crates that are NOT vendored under /root/reference
(no Cargo.lock):
  * halo2_proofs (privacy-scaling-explorations/halo2, tag believed v2022_10_22, selected
    transitively by halo2wrong rev 63bde545, reference Cargo.toml:13-14):
    arithmetic::best_fft, arithmetic::best_multiexp, poly::EvaluationDomain,
    poly::kzg::commitment::ParamsKZG.
  * halo2curves (bn256::{Fr,Fq,G1Affine,G1}).
The reference's only call sites are benches/bench.rs:235 (ParamsKZG::setup),
:236-237 (keygen_vk/pk) and :321-329 (create_proof).  No reference test pins values at
this boundary (parity unpinned), so this oracle is anchored on mathematics instead:
group/field elements have canonical encodings, so any correct algorithm gives the same
bytes.  Plain Python integers, no Montgomery tricks: deliberately a different algorithm
from the product's.

Memory format (halo2curves): Fr/Fq = 4 x u64 little-endian limbs, Montgomery form
(R = 2^256); G1Affine = {x, y} 64 B, identity = (0, 0).
"""
from __future__ import annotations

import hashlib
import struct

# --- constants (verified: primes, 2-adicity, generator orders; see tests/test_oracle_bn254.py)
R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001  # Fr modulus r
Q_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # Fq modulus q
MONT_R = 1 << 256
S_ADICITY = 28
MULT_GEN = 7
ROOT_OF_UNITY = pow(MULT_GEN, (R_MOD - 1) >> S_ADICITY, R_MOD)
ZETA = 0x30644E72E131A029048B6E193FD84104CC37A73FEC2BC5E9B8CA0B2D36636F23
DELTA = pow(MULT_GEN, 1 << S_ADICITY, R_MOD)
CURVE_B = 3
G1_GEN = (1, 2)


# --- encodings -------------------------------------------------------------------------
def to_mont(x: int, mod: int) -> int:
    return (x * MONT_R) % mod


def from_mont(x: int, mod: int) -> int:
    return (x * pow(MONT_R, -1, mod)) % mod


def fe_to_bytes(x: int, mod: int = R_MOD) -> bytes:
    """canonical int -> 32 bytes in halo2curves memory format (Montgomery, 4 LE u64)."""
    return to_mont(x % mod, mod).to_bytes(32, "little")


def fe_from_bytes(b: bytes, mod: int = R_MOD) -> int:
    return from_mont(int.from_bytes(b, "little"), mod)


def fr_array_to_bytes(xs) -> bytes:
    return b"".join(fe_to_bytes(x, R_MOD) for x in xs)


def fr_array_from_bytes(b: bytes):
    return [fe_from_bytes(b[i:i + 32], R_MOD) for i in range(0, len(b), 32)]


def g1_affine_to_bytes(p) -> bytes:
    if p is None:
        return b"\0" * 64
    return fe_to_bytes(p[0], Q_MOD) + fe_to_bytes(p[1], Q_MOD)


def g1_affine_from_bytes(b: bytes):
    x = fe_from_bytes(b[:32], Q_MOD)
    y = fe_from_bytes(b[32:64], Q_MOD)
    if x == 0 and y == 0:
        return None
    return (x, y)


def g1_array_to_bytes(ps) -> bytes:
    return b"".join(g1_affine_to_bytes(p) for p in ps)


def g1_jacobian_from_bytes(b: bytes):
    """96-byte {x,y,z} Montgomery Jacobian -> affine tuple or None."""
    x = fe_from_bytes(b[:32], Q_MOD)
    y = fe_from_bytes(b[32:64], Q_MOD)
    z = fe_from_bytes(b[64:96], Q_MOD)
    if z == 0:
        return None
    zi = pow(z, -1, Q_MOD)
    return (x * zi * zi % Q_MOD, y * zi * zi * zi % Q_MOD)


# --- G1 (y^2 = x^3 + 3 over Fq), affine, None = identity ------------------------------
def g1_is_on_curve(p) -> bool:
    if p is None:
        return True
    x, y = p
    return (y * y - x * x * x - CURVE_B) % Q_MOD == 0


def g1_neg(p):
    return None if p is None else (p[0], (-p[1]) % Q_MOD)


def g1_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if (y1 + y2) % Q_MOD == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, Q_MOD) % Q_MOD
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, Q_MOD) % Q_MOD
    x3 = (lam * lam - x1 - x2) % Q_MOD
    y3 = (lam * (x1 - x3) - y1) % Q_MOD
    return (x3, y3)


def _jac_double(P):
    X, Y, Z = P
    if Z == 0:
        return P
    A = X * X % Q_MOD
    B = Y * Y % Q_MOD
    C = B * B % Q_MOD
    D = 2 * ((X + B) * (X + B) - A - C) % Q_MOD
    E = 3 * A % Q_MOD
    F = E * E % Q_MOD
    X3 = (F - 2 * D) % Q_MOD
    Y3 = (E * (D - X3) - 8 * C) % Q_MOD
    Z3 = 2 * Y * Z % Q_MOD
    return (X3, Y3, Z3)


def _jac_add_affine(P, q):
    X1, Y1, Z1 = P
    if q is None:
        return P
    if Z1 == 0:
        return (q[0], q[1], 1)
    x2, y2 = q
    Z1Z1 = Z1 * Z1 % Q_MOD
    U2 = x2 * Z1Z1 % Q_MOD
    S2 = y2 * Z1 * Z1Z1 % Q_MOD
    if U2 == X1:
        if S2 == Y1:
            return _jac_double(P)
        return (0, 1, 0)
    H = (U2 - X1) % Q_MOD
    HH = H * H % Q_MOD
    HHH = H * HH % Q_MOD
    rr = (S2 - Y1) % Q_MOD
    V = X1 * HH % Q_MOD
    X3 = (rr * rr - HHH - 2 * V) % Q_MOD
    Y3 = (rr * (V - X3) - Y1 * HHH) % Q_MOD
    Z3 = Z1 * H % Q_MOD
    return (X3, Y3, Z3)


def _jac_to_affine(P):
    X, Y, Z = P
    if Z == 0:
        return None
    zi = pow(Z, -1, Q_MOD)
    return (X * zi * zi % Q_MOD, Y * zi * zi * zi % Q_MOD)


def g1_mul(p, k: int):
    """k * p by double-and-add (MSB first), Jacobian accumulator."""
    k %= R_MOD
    if p is None or k == 0:
        return None
    acc = (0, 1, 0)
    for bit in bin(k)[2:]:
        acc = _jac_double(acc)
        if bit == "1":
            acc = _jac_add_affine(acc, p)
    return _jac_to_affine(acc)


def msm_naive(scalars, points):
    """sum_i scalars[i] * points[i]; mathematical definition of best_multiexp's result."""
    acc = None
    for s, p in zip(scalars, points):
        acc = g1_add(acc, g1_mul(p, s))
    return acc


def g1_multiples(n: int, start=G1_GEN):
    """[1*G, 2*G, ..., n*G] as affine points (bench config 5 bases)."""
    out = []
    acc = (start[0], start[1], 1)
    pts = []
    for _ in range(n):
        pts.append(acc)
        acc = _jac_add_affine(acc, start)
    # batch normalise
    zs = [P[2] for P in pts]
    pref = [1] * (n + 1)
    for i, z in enumerate(zs):
        pref[i + 1] = pref[i] * z % Q_MOD
    inv = pow(pref[n], -1, Q_MOD)
    zinv = [0] * n
    for i in range(n - 1, -1, -1):
        zinv[i] = inv * pref[i] % Q_MOD
        inv = inv * zs[i] % Q_MOD
    for (X, Y, _), zi in zip(pts, zinv):
        out.append((X * zi * zi % Q_MOD, Y * zi * zi * zi % Q_MOD))
    return out


# --- NTT over Fr ------------------------------------------------------------------------
def omega_for(log_n: int) -> int:
    """EvaluationDomain's omega: ROOT_OF_UNITY^(2^(S - log_n))."""
    return pow(ROOT_OF_UNITY, 1 << (S_ADICITY - log_n), R_MOD)


def dft_naive(a, omega):
    """O(n^2) definition of best_fft's result: out[i] = sum_j a[j] * omega^(i*j)."""
    n = len(a)
    pw = [1] * n
    for i in range(1, n):
        pw[i] = pw[i - 1] * omega % R_MOD
    return [sum(a[j] * pw[(i * j) % n] for j in range(n)) % R_MOD for i in range(n)]


def ntt(a, omega):
    """recursive radix-2 (different structure from the product's kernels); natural in/out."""
    n = len(a)
    if n == 1:
        return list(a)
    w2 = omega * omega % R_MOD
    ev = ntt(a[0::2], w2)
    od = ntt(a[1::2], w2)
    out = [0] * n
    w = 1
    h = n // 2
    for i in range(h):
        t = w * od[i] % R_MOD
        out[i] = (ev[i] + t) % R_MOD
        out[i + h] = (ev[i] - t) % R_MOD
        w = w * omega % R_MOD
    return out


def intt(a, omega):
    n = len(a)
    ninv = pow(n, -1, R_MOD)
    return [x * ninv % R_MOD for x in ntt(a, pow(omega, -1, R_MOD))]


def coeff_to_extended(coeffs, k: int, ext_k: int):
    """EvaluationDomain::coeff_to_extended: evaluate on the coset ZETA * <omega_ext>.

    distribute_powers_zeta multiplies coefficient i by [1, ZETA, ZETA^2][i % 3]
    (g_coset = ZETA), zero-pads to 2^ext_k, then best_fft with extended_omega.
    """
    assert len(coeffs) == 1 << k
    zp = [1, ZETA, ZETA * ZETA % R_MOD]
    a = [c * zp[i % 3] % R_MOD for i, c in enumerate(coeffs)]
    a += [0] * ((1 << ext_k) - len(a))
    return ntt(a, omega_for(ext_k))


def extended_to_coeff(evals, ext_k: int):
    """EvaluationDomain::extended_to_coeff: inverse of coeff_to_extended on the full
    extended length (callers truncate): iFFT, then multiply coefficient i by
    [1, ZETA^-1 = ZETA^2, ZETA^-2 = ZETA][i % 3]."""
    a = intt(evals, omega_for(ext_k))
    zi = [1, ZETA * ZETA % R_MOD, ZETA]
    return [c * zi[i % 3] % R_MOD for i, c in enumerate(a)]


# --- KZG SRS with a seeded secret ----------------------------------
def srs_secret(k: int) -> int:
    h = hashlib.blake2b(b"b2r-srs" + struct.pack("<I", k), digest_size=64).digest()
    return int.from_bytes(h, "little") % R_MOD


def lagrange_at(s: int, k: int):
    """[L_i(s)] over the size-2^k domain: L_i(s) = omega^i (s^n - 1) / (n (s - omega^i))."""
    n = 1 << k
    w = omega_for(k)
    zn = (pow(s, n, R_MOD) - 1) % R_MOD
    ninv = pow(n, -1, R_MOD)
    out = []
    wi = 1
    for _ in range(n):
        out.append(wi * zn % R_MOD * ninv % R_MOD * pow((s - wi) % R_MOD, -1, R_MOD) % R_MOD)
        wi = wi * w % R_MOD
    return out


def eval_poly(coeffs, x: int) -> int:
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R_MOD
    return acc


# --- seeded Fr stream (shared convention with the product's test inputs) --------------
def fr_stream(seed: int, n: int):
    """n uniform-ish Fr values: blake2b(seed || counter) mod r, 64-byte digests."""
    out = []
    for i in range(n):
        h = hashlib.blake2b(struct.pack("<QQ", seed, i), digest_size=64).digest()
        out.append(int.from_bytes(h, "little") % R_MOD)
    return out
