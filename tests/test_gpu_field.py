"""GPU: the device field arithmetic of csrc/field.cuh against Python integers, through the C ABI
(b2r_field_selftest).  bn256::{Fr, Fq} hold values in Montgomery form (x R mod p, R = 2^256); every operation below is
compared on the raw Montgomery words, so a single wrong carry shows.  Operands: random, the extremes p - 1 .. p - 4 in
every position (the bounds of the fused sums of products are tight there), 0, 1, and single-word values."""
import numpy as np
import pytest

import bn254 as O
from util import int_to_limbs, limbs_to_int

pytestmark = pytest.mark.gpu

R = O.MONT_R


def to_np(xs):
    return np.array([int_to_limbs(x) for x in xs], dtype=np.uint64).reshape(-1, 4)


def from_np(a):
    return [limbs_to_int(r) for r in np.asarray(a).reshape(-1, 4)]


def operands(mod, n, seed):
    rng = np.random.default_rng(seed)

    def col(kind):
        out = []
        for i in range(n):
            k = kind if kind is not None else i % 8
            if k == 5:
                out.append(mod - 1 - int(rng.integers(0, 4)))
            elif k == 6:
                out.append(int(rng.integers(0, 4)))
            elif k == 7:
                out.append(int(rng.integers(0, 1 << 32)) << (32 * int(rng.integers(0, 8))))
            else:
                out.append(int.from_bytes(rng.bytes(40), "little") % mod)
        return [x % mod for x in out]
    a, b, c, d = col(None), col(None), col(None), col(None)
    # a block where every operand is an extreme at the same time
    m = n // 8
    for v in (a, b, c, d):
        v[:m] = [mod - 1 - int(rng.integers(0, 4)) for _ in range(m)]
    return a, b, c, d


@pytest.mark.parametrize("field", ["fr", "fq"])
def test_device_field_ops_match_python_integers(ctx, field):
    mod = O.R_MOD if field == "fr" else O.Q_MOD
    rinv = pow(R, -1, mod)
    n = 4096
    a, b, c, d = operands(mod, n, 11 if field == "fr" else 12)
    A, B, Cc, D = to_np(a), to_np(b), to_np(c), to_np(d)
    exp = {
        "mul": [x * y * rinv % mod for x, y in zip(a, b)],
        "sqr": [x * x * rinv % mod for x in a],
        "mul_add_mul": [(x * y + z * w) * rinv % mod for x, y, z, w in zip(a, b, c, d)],
        "mul_sub_mul": [(x * y - z * w) * rinv % mod for x, y, z, w in zip(a, b, c, d)],
        "dot4": [(x * y + z * w + x * z + y * w) * rinv % mod for x, y, z, w in zip(a, b, c, d)],
        "add": [(x + y) % mod for x, y in zip(a, b)],
        "sub": [(x - y) % mod for x, y in zip(a, b)],
    }
    for op, want in exp.items():
        got = from_np(ctx.field_selftest(field, op, A, B, Cc, D))
        bad = [i for i in range(n) if got[i] != want[i]]
        assert not bad, (field, op, bad[:4], hex(a[bad[0]]), hex(b[bad[0]]))
    # inverse (binary Euclid on the device): inv(aR) = a^-1 R, i.e. a_m * inv_m = R^2 (mod p); inv(0) = 0
    k = 256
    got = from_np(ctx.field_selftest(field, "inv", A[:k]))
    for x, y in zip(a[:k], got):
        assert (y == 0) if x == 0 else (x * y % mod == R * R % mod)


def test_field_selftest_rejects_bad_arguments(ctx):
    import b2rsa
    a = to_np([1, 2])
    with pytest.raises(b2rsa.B2RError):
        ctx.field_selftest("fr", "mul", a)              # second operand missing
    with pytest.raises(b2rsa.B2RError):
        ctx.field_selftest("fq", "dot4", a, a)          # c, d missing
