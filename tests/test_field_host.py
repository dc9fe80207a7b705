"""CPU: the device field algorithm (csrc/field.cuh: interleaved two-accumulator Montgomery product,
add/sub/neg/inv) compiled for the HOST by g++ (row primitives fall back to their C emulation) and
checked against Python integers.  This is a test of the kernel source's carry logic in a container
without a GPU - not a CPU fallback: no product path calls the host build."""
import os
import subprocess
import tempfile

import bn254 as O

HERE = os.path.dirname(os.path.abspath(__file__))


import pytest


@pytest.mark.parametrize("sparse_p0", [0, 1])   # 1: the opt-in Fr reduction rows without a multiplier for the low word
def test_field_cuh_host_build_matches_python(sparse_p0):
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "fht")
        subprocess.check_call(["g++", "-O1", "-std=c++17", f"-DB2R_SPARSE_P0={sparse_p0}", "-o", exe, os.path.join(HERE, "host", "field_host_test.cpp")])
        out = subprocess.check_output([exe, "600"], text=True)
    R = O.MONT_R
    n = 0
    for line in out.splitlines():
        f = line.split()
        mod = O.R_MOD if f[0] == "fr" else O.Q_MOD
        a, b, mul, add, sub, neg, inv, sqr = (int(x, 16) for x in f[1:9])
        rinv = pow(R, -1, mod)
        assert mul == a * b * rinv % mod
        assert sqr == a * a * rinv % mod      # the dedicated squaring (36 + 64 wide products)
        assert add == (a + b) % mod and sub == (a - b) % mod and neg == (-a) % mod
        if inv:
            # Montgomery inverse: inv(aR) = a^-1 R  =>  a_m * inv_m = R^2 (mod p)
            assert a * inv % mod == R * R % mod
        a, b, c, d, mam, msm, dot4 = (int(x, 16) for x in f[9:16])
        assert mam == (a * b + c * d) * rinv % mod       # one reduction for two products
        assert msm == (a * b - c * d) * rinv % mod
        assert dot4 == (a * b + c * d + a * c + b * d) * rinv % mod
        n += 1
    assert n == 1200
