"""Shared helpers for the tests: conversions between Python integers (the oracle's domain)
and the halo2curves memory format the C ABI uses (uint64[...,4] Montgomery limbs)."""
import numpy as np

import bn254 as O

MASK64 = (1 << 64) - 1


def int_to_limbs(x: int) -> list:
    return [(x >> (64 * i)) & MASK64 for i in range(4)]


def limbs_to_int(l) -> int:
    return int(l[0]) | (int(l[1]) << 64) | (int(l[2]) << 128) | (int(l[3]) << 192)


def fr_to_np(xs) -> np.ndarray:
    """canonical ints -> uint64[n,4] Montgomery"""
    return np.array([int_to_limbs(O.to_mont(x % O.R_MOD, O.R_MOD)) for x in xs], dtype=np.uint64).reshape(-1, 4)


def np_to_fr(a: np.ndarray) -> list:
    a = np.asarray(a).reshape(-1, 4)
    rinv = pow(O.MONT_R, -1, O.R_MOD)
    return [limbs_to_int(r) * rinv % O.R_MOD for r in a]


def g1_to_np(ps) -> np.ndarray:
    out = np.zeros((len(ps), 8), dtype=np.uint64)
    for i, p in enumerate(ps):
        if p is None:
            continue
        out[i, :4] = int_to_limbs(O.to_mont(p[0], O.Q_MOD))
        out[i, 4:] = int_to_limbs(O.to_mont(p[1], O.Q_MOD))
    return out


def np_to_g1(a: np.ndarray):
    a = np.asarray(a).reshape(-1, 8)
    qinv = pow(O.MONT_R, -1, O.Q_MOD)
    out = []
    for r in a:
        x = limbs_to_int(r[:4]) * qinv % O.Q_MOD
        y = limbs_to_int(r[4:]) * qinv % O.Q_MOD
        out.append(None if (x == 0 and y == 0) else (x, y))
    return out


def np_jac_to_g1(a: np.ndarray):
    """uint64[12] Jacobian (Montgomery) -> affine tuple / None"""
    a = np.asarray(a).reshape(12)
    qinv = pow(O.MONT_R, -1, O.Q_MOD)
    x, y, z = (limbs_to_int(a[4 * i:4 * i + 4]) * qinv % O.Q_MOD for i in range(3))
    if z == 0:
        return None
    zi = pow(z, -1, O.Q_MOD)
    return (x * zi * zi % O.Q_MOD, y * zi * zi * zi % O.Q_MOD)


def random_fr_np(n: int, seed: int) -> np.ndarray:
    """n uniform field elements directly as Montgomery limbs (the Montgomery map is a
    bijection of [0, r), so uniform limbs below r are uniform field elements)."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 62) - 1)
    top = np.uint64(O.R_MOD >> 192)
    bad = a[:, 3] >= top  # conservative rejection on the top limb only
    while bad.any():
        a[bad, 3] = rng.integers(0, 1 << 62, size=int(bad.sum()), dtype=np.uint64)
        bad = a[:, 3] >= top
    return a
