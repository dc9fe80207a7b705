"""CPU: pins the oracle for hot path (a) (oracle/rsa_witness.c) against the reference's own
known-answer tests.  Fixtures are the reference's literals, extracted by
tests/golden/make_kats.py:
  * RSA-2048 PKCS#1 v1.5 circuits            /root/reference/src/chip.rs:683-803
  * BigIntChip::mul unreduced-product KATs   /root/reference/src/big_integer/chip.rs:2797-3100
  * mul_mod identities                       /root/reference/src/big_integer/chip.rs:3109-3264
  * pow_mod_fixed_exp vs big_pow_mod         /root/reference/src/big_integer/chip.rs:2314-2355
Positive cases must satisfy every constraint of the MockProver-style checker, negative twins
must violate at least one - the reference asserts exactly that (prover.verify() is Ok / Err).
"""
import json
import os
import random

import numpy as np
import pytest

import cpu_oracle as CO
import rsa_fixtures as RF

HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "rsa_kats.json")))
MUL_KATS = json.load(open(os.path.join(HERE, "golden", "bigint_mul_kats.json")))
B = 1 << 64


@pytest.mark.parametrize("kat", KATS, ids=[k["name"] for k in KATS])
def test_rsa_signature_kats(kat):
    n, sig, h = int(kat["n"]), int(kat["sig"]), int(kat["hash"])
    v, adv, rows, bad, msg = CO.rsa_synthesize(2048, 17, n, sig, h)
    em_ok = pow(sig, 65537, n) == RF.emsa_pkcs1_v15(h.to_bytes(32, "big"), 2048)
    assert em_ok == (not kat["should_be_error"])          # the fixture itself, by plain integer arithmetic
    assert v == (0 if kat["should_be_error"] else 1)
    assert (bad > 0) == kat["should_be_error"], msg
    assert rows <= (1 << 17) - 6
    assert adv.shape == (5, 1 << 17, 4)


def test_rsa_row_budget_all_sizes():
    """the reference's k per key size (benches/bench.rs:353-373, src/chip.rs:337,666) must hold the
    restated layout: 1024 -> k=15, 2048 -> k=17, 4096 -> k=18"""
    for bits, k in ((1024, 15), (2048, 17), (4096, 18)):
        n, s, h = RF.instance(bits, 0)
        v, _, rows, bad, msg = CO.rsa_synthesize(bits, k, n, s, h)
        assert v == 1 and bad == 0, (bits, msg)
        assert (1 << (k - 1)) < rows <= (1 << k) - 6, (bits, rows)


def test_rsa_wrong_inputs_violate():
    n, s, h = RF.instance(2048, 3)
    for nn, ss, hh in ((n, s, h ^ 1), (n, s ^ 2, h), (n, (s + n) % (1 << 2048), h)):
        v, _, _, bad, _ = CO.rsa_synthesize(2048, 17, nn, ss, hh)
        assert v <= 0 and (v < 0 or bad > 0)


@pytest.mark.parametrize("kat", MUL_KATS, ids=[k["name"] for k in MUL_KATS])
def test_bigint_mul_kats(kat):
    a, b, ans = int(kat["a"]), int(kat["b"]), int(kat["ans"])
    limbs, bad = CO.bigint_op("mul_kat", 2048, 16, a, b, ans)
    assert bad == 0
    al = [(a >> (64 * i)) % B for i in range(32)]
    bl = [(b >> (64 * i)) % B for i in range(32)]
    conv = [sum(al[j] * bl[i - j] for j in range(32) if 0 <= i - j < 32) for i in range(63)]
    assert limbs == conv                                   # unreduced, no carries (chip.rs:400-412)
    assert sum(l << (64 * i) for i, l in enumerate(limbs)) == ans
    # negative twin (reference: test_bad_... variants): a wrong expected product must not verify
    _, bad = CO.bigint_op("mul_kat", 2048, 16, a, b, ans + 1)
    assert bad > 0


def _rand_n(rng, bits=2048):
    return rng.getrandbits(bits) | (1 << (bits - 1)) | 1


def test_bigint_mulmod_identities():
    rng = random.Random(0xB200)
    n = _rand_n(rng)
    b = rng.randrange(n)
    cases = [(0, b, 0), (n, 1, 0), (n - 1, n - 1, 1), (n - 1, n - 2, 2), (b, rng.randrange(n), None)]
    for a_, b_, want in cases:
        limbs, bad = CO.bigint_op("mul_mod", 2048, 16, a_, b_, n)
        assert bad == 0
        got = sum(l << (64 * i) for i, l in enumerate(limbs))
        assert got == (a_ * b_) % n
        if want is not None:
            assert got == want


def test_bigint_mulmod_zero_modulus_panics():
    limbs, _ = CO.bigint_op("mul_mod", 2048, 16, 5, 7, 0)
    assert limbs is None                                   # BigUint division by zero panics in the reference


@pytest.mark.parametrize("bits,k", [(1024, 15), (2048, 17)])
def test_bigint_pow_mod_fixed_exp(bits, k):
    rng = random.Random(bits)
    n = _rand_n(rng, bits)
    a = rng.randrange(n)
    for e in (65537, rng.getrandbits(7) | 1, 2):
        limbs, bad = CO.bigint_op("pow_mod_fixed_exp", bits, k, a, e, n)
        assert bad == 0
        assert sum(l << (64 * i) for i, l in enumerate(limbs)) == pow(a, e, n)


def test_bigint_add_sub_in_field():
    rng = random.Random(7)
    n = _rand_n(rng)
    a, b = rng.randrange(n), rng.randrange(n)
    limbs, bad = CO.bigint_op("add", 2048, 16, a, b)
    assert bad == 0 and len(limbs) == 33 and sum(l << (64 * i) for i, l in enumerate(limbs)) == a + b
    hi, lo = max(a, b), min(a, b)
    limbs, bad = CO.bigint_op("sub", 2048, 16, hi, lo)
    assert bad == 0 and limbs[-1] == 0 and sum(l << (64 * i) for i, l in enumerate(limbs[:-1])) == hi - lo
    limbs, bad = CO.bigint_op("sub", 2048, 16, lo, hi)      # chip.rs:310-373: returns b - a with the overflow bit set
    assert bad == 0 and limbs[-1] == 1 and sum(l << (64 * i) for i, l in enumerate(limbs[:-1])) == hi - lo
    _, bad = CO.bigint_op("assert_in_field", 2048, 16, lo, 0, hi)
    assert bad == 0
    _, bad = CO.bigint_op("assert_in_field", 2048, 16, hi, 0, lo)
    assert bad > 0
    _, bad = CO.bigint_op("assert_in_field", 2048, 16, hi, 0, hi)
    assert bad > 0


def test_fixture_generators_are_consistent():
    nl, sl, hl = RF.batch(2048, 3, start=5)
    assert nl.shape == (3, 32) and sl.shape == (3, 32) and hl.shape == (3, 4) and nl.dtype == np.uint64
    n, s, h = RF.instance(2048, 6)
    assert [int(x) for x in nl[1]] == [(n >> (64 * j)) % B for j in range(32)]
    assert pow(s, 65537, n) == RF.emsa_pkcs1_v15(h.to_bytes(32, "big"), 2048)


def _rand(bits, seed):
    import random
    return random.Random(seed).getrandbits(bits)


def test_refresh_matches_product():
    """chip.rs:1861-1899 (test_refresh): mul(a, b) refreshed to 64-bit limbs is the integer product, limb by limb;
    RefreshAux::new(32, 1, 1).increased_limbs_vec == [1, 0] is the reference's own KAT (mod.rs:504-509), checked in
    tests/test_host_circuit.py against the recorder."""
    for bits, k, seed in ((512, 14, 1), (1024, 15, 2)):
        a, b = _rand(bits, seed), _rand(bits, seed + 100)
        limbs, bad = CO.bigint_op("refresh", bits, k, a, b)
        assert bad == 0 and limbs is not None
        assert len(limbs) == 2 * (bits // 64)
        assert all(l < (1 << 64) for l in limbs)
        assert sum(l << (64 * i) for i, l in enumerate(limbs)) == a * b


def test_add_mod_sub_mod_match_integers():
    """chip.rs:1948-2110: add_mod / sub_mod for a, b < n, including the wrap-around cases.  BigIntChip::sub flags
    a - b as overflowed when a <= b (chip.rs:327-331: the limb test is a + max - b >= 2^(64 n2), i.e. a > b), so the
    reference returns the UNREDUCED representative n when the true result is 0: sub_mod(a, a, n) = n and
    add_mod(a, n - a, n) = n.  The restatement reproduces that."""
    bits, k = 1024, 15
    n = _rand(bits, 7) | (1 << (bits - 1)) | 1
    val = lambda limbs: sum(l << (64 * i) for i, l in enumerate(limbs))
    cases = [(_rand(bits, 8) % n, _rand(bits, 9) % n), (n - 1, n - 1), (0, 0), (5, n - 3), (n - 3, 5), (7, 7), (9, n - 9)]
    for a, b in cases:
        limbs, bad = CO.bigint_op("add_mod", bits, k, a, b, n)
        assert bad == 0 and val(limbs) == ((a + b) % n or (n if a + b else 0)), (a, b)
        limbs, bad = CO.bigint_op("sub_mod", bits, k, a, b, n)
        assert bad == 0 and val(limbs) == ((a - b) % n or n), (a, b)
    # chip.rs:2111-2148 (test_bad_sub_mod) style: an operand >= n breaks a constraint or is what the reference panics on
    limbs, bad = CO.bigint_op("sub_mod", bits, k, 1, n + 5, n)
    assert limbs is None or bad > 0


def test_pow_mod_variable_exponent():
    """chip.rs:2229-2271 (test_pow_mod): 5 exponent bits, against pow(a, e, n); an exponent wider than exp_limb_bits
    cannot be composed from its bits (to_bits) and fails"""
    bits, k = 512, 15
    n = _rand(bits, 21) | (1 << (bits - 1)) | 1
    a = _rand(bits, 22) % n
    for e in (0, 1, 2, 17, 31):
        limbs, bad = CO.bigint_op("pow_mod", bits, k, a, e, n, exp_limb_bits=5)
        assert bad == 0 and sum(l << (64 * i) for i, l in enumerate(limbs)) == pow(a, e, n), e
    _, bad = CO.bigint_op("pow_mod", bits, k, a, 33, n, exp_limb_bits=5)
    assert bad > 0


def test_predicates_match_integer_comparisons():
    """chip.rs:2395-2795 (test_is_zero ... test_bad_in_field): the seven predicates against Python integer comparison,
    on the reference's own shapes - a >> 128 < b, a <= a, a > a >> 128, a >= a, a in field n - and their negations"""
    bits, k = 512, 15
    a, b = _rand(bits, 31) | (1 << (bits - 1)), _rand(bits, 32) | (1 << (bits - 2))
    small = a >> 128
    truth = {"is_equal_fresh": lambda x, y: x == y, "is_less_than": lambda x, y: x < y, "is_less_than_or_equal": lambda x, y: x <= y,
             "is_greater_than": lambda x, y: x > y, "is_greater_than_or_equal": lambda x, y: x >= y, "is_in_field": lambda x, y: x < y}
    pairs = [(small, b), (b, small), (a, a), (a, a - 1), (a - 1, a), (0, 0), (0, 1), (a, a ^ (1 << 300)), ((1 << bits) - 1, (1 << bits) - 1)]
    for op, f in truth.items():
        for x, y in pairs:
            limbs, bad = CO.bigint_op(op, bits, k, x, y)
            assert bad == 0 and limbs == [int(f(x, y))], (op, x, y)
    for x in (0, 1, a, 1 << 448):
        limbs, bad = CO.bigint_op("is_zero", bits, k, x, 0)
        assert bad == 0 and limbs == [int(x == 0)]


def test_square_and_square_mod():
    """chip.rs:431-437 / 642-649: square = mul(a, a) (unreduced limbs), square_mod = mul_mod(a, a, n)"""
    bits, k = 512, 15
    nl = bits // 64
    n = _rand(bits, 41) | (1 << (bits - 1)) | 1
    a = _rand(bits, 42) % n
    limbs, bad = CO.bigint_op("square", bits, k, a, 0)
    al = [(a >> (64 * i)) & ((1 << 64) - 1) for i in range(nl)]
    assert bad == 0 and limbs == [sum(al[j] * al[i - j] for j in range(nl) if 0 <= i - j < nl) for i in range(2 * nl - 1)]
    limbs, bad = CO.bigint_op("square_mod", bits, k, a, 0, n)
    assert bad == 0 and sum(l << (64 * i) for i, l in enumerate(limbs)) == a * a % n


def test_rsa_variable_exponent_circuit():
    """src/chip.rs:372-400 shape (RSAPubE::Var): the pkcs1v15 circuit with e = 65537 given as a witness, 17 exponent bits"""
    bits, k = 512, 17
    n, sig, h = RF.instance(bits, 3)
    nl = bits // 64
    t = CO.RsaTable(bits, k)
    assert t.synthesize_var(RF.limbs64(n, nl), RF.limbs64(sig, nl), RF.limbs64(h, 4), 65537, 17) == 1
    assert t.check()[0] == 0
    t.free()
    t = CO.RsaTable(bits, k)                                  # wrong exponent: the power is not the encoded message
    assert t.synthesize_var(RF.limbs64(n, nl), RF.limbs64(sig, nl), RF.limbs64(h, 4), 65539, 17) == 0
    assert t.check()[0] > 0
    t.free()
