"""GPU parity: RSA witness synthesis (recorded program replayed on the GPU, through the C ABI)
against the CPU oracle's row-by-row restatement of the reference (oracle/rsa_witness.c).

Mirrors the reference's own tests for this path: the RSA-2048 PKCS#1 v1.5 known-answer circuits
(src/chip.rs:683-816, fixtures in tests/golden/rsa_kats.json) and the bench circuit
(benches/bench.rs:132-225) on seeded synthetic keys.  Bit-exact: all 5 x 2^k advice cells.
"""
import json
import os

import numpy as np
import pytest

import bn254 as O
import cpu_oracle as CO
import plonk as PL
import rsa_fixtures as RF

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "rsa_kats.json")))


@pytest.fixture(scope="module")
def prog2048(ctx):
    p = ctx.rsa_program(2048, 17)
    yield p
    p.free()


def test_program_shape(prog2048):
    info = prog2048.info()
    # same layout as the oracle: rows used must agree exactly
    n, s, h = RF.instance(2048, 0)
    _, _, rows, bad, _ = CO.rsa_synthesize(2048, 17, n, s, h)
    assert bad == 0
    assert info["rows_used"] == rows
    assert info["rows_used"] <= (1 << 17) - 6


@pytest.mark.parametrize("kat", KATS, ids=[k["name"] for k in KATS])
def test_reference_kats(prog2048, kat):
    n, sig, h = int(kat["n"]), int(kat["sig"]), int(kat["hash"])
    adv, valid = prog2048.witness_batch(RF.limbs64(n, 32)[None], RF.limbs64(sig, 32)[None], RF.limbs64(h, 4)[None])
    v, want, rows, bad, msg = CO.rsa_synthesize(2048, 17, n, sig, h)
    assert int(valid[0]) == (0 if kat["should_be_error"] else 1)
    assert v == int(valid[0])
    assert (bad != 0) == kat["should_be_error"]
    assert np.array_equal(adv[0], want)


def test_batch_synthetic_2048(prog2048):
    batch = 6
    nl, sl, hl = RF.batch(2048, batch)
    adv, valid = prog2048.witness_batch(nl, sl, hl)
    assert valid.tolist() == [1] * batch
    for i in range(batch):
        t = CO.RsaTable(2048, 17)
        assert t.synthesize(nl[i], sl[i], hl[i]) == 1
        assert t.check()[0] == 0
        assert np.array_equal(adv[i], t.advice()), f"instance {i}"
        t.free()


def test_invalid_inputs_2048(prog2048):
    """wrong hash, sig >= n (assert_in_field fails), n = 0 (division by zero: the reference panics)"""
    n, s, h = RF.instance(2048, 1)
    cases = [(n, s, h ^ 1), (n, s + n, h), (n, 0, h), (n, 1, h)]
    nl = np.stack([RF.limbs64(c[0], 32) for c in cases])
    sl = np.stack([RF.limbs64(c[1] % (1 << 2048), 32) for c in cases])
    hl = np.stack([RF.limbs64(c[2], 4) for c in cases])
    adv, valid = prog2048.witness_batch(nl, sl, hl)
    for i, c in enumerate(cases):
        v, want, rows, bad, msg = CO.rsa_synthesize(2048, 17, c[0], c[1] % (1 << 2048), c[2])
        if v < 0:
            assert int(valid[i]) == 0xFF
        else:
            assert int(valid[i]) == v == 0
            assert bad > 0
            assert np.array_equal(adv[i], want), f"case {i}"
    # zero modulus -> reference panics (BigUint division by zero) -> 0xFF status
    adv, valid = prog2048.witness_batch(np.zeros((1, 32), dtype=np.uint64), RF.limbs64(s, 32)[None], RF.limbs64(h, 4)[None])
    assert int(valid[0]) == 0xFF


@pytest.mark.parametrize("bits,k", [(1024, 15), (4096, 18)])
def test_other_key_sizes(ctx, bits, k):
    prog = ctx.rsa_program(bits, k)
    nl, sl, hl = RF.batch(bits, 2)
    adv, valid = prog.witness_batch(nl, sl, hl)
    assert valid.tolist() == [1, 1]
    for i in range(2):
        t = CO.RsaTable(bits, k)
        assert t.synthesize(nl[i], sl[i], hl[i]) == 1
        assert prog.info()["rows_used"] == t.rows()
        assert np.array_equal(adv[i], t.advice())
        t.free()
    prog.free()


@pytest.mark.parametrize("bits,k", [(768, 15), (3072, 18)])
def test_limb_counts_between_the_baseline_sizes(ctx, bits, k):
    """12 and 48 limbs (the BASELINE configs are 16 / 32 / 64): random odd moduli and random signatures below them - not
    valid signatures, which does not matter for the arithmetic: sig^65537 mod n, the 19 carry chains and the PKCS#1
    comparison must produce the oracle's table cell for cell, with is_valid = 0"""
    import random
    r = random.Random(bits)
    nl = bits // 64
    ns, ss, hs = [], [], []
    for _ in range(2):
        n = r.getrandbits(bits) | (1 << (bits - 1)) | 1
        ns.append(CO.int_to_limbs64(n, nl)); ss.append(CO.int_to_limbs64(r.getrandbits(bits) % n, nl)); hs.append(CO.int_to_limbs64(r.getrandbits(256), 4))
    prog = ctx.rsa_program(bits, k)
    adv, valid = prog.witness_batch(np.stack(ns), np.stack(ss), np.stack(hs))
    assert valid.tolist() == [0, 0]
    for i in range(2):
        t = CO.RsaTable(bits, k)
        assert t.synthesize(ns[i], ss[i], hs[i]) == 0
        assert prog.info()["rows_used"] == t.rows()
        assert np.array_equal(adv[i], t.advice())
        t.free()
    prog.free()


def test_layout_errors(ctx):
    import b2rsa
    with pytest.raises(b2rsa.B2RError) as e:
        ctx.rsa_program(2048, 16)  # does not fit 2^16 rows
    assert e.value.code == b2rsa.ERR_LAYOUT
    with pytest.raises(b2rsa.B2RError):
        ctx.rsa_program(2000, 17)  # bits_len % 64 != 0


def test_blinding_rows(prog2048):
    seed = 0x1234
    nl, sl, hl = RF.batch(2048, 2)
    adv, _ = prog2048.witness_batch(nl, sl, hl, blind_seed=seed)
    adv0, _ = prog2048.witness_batch(nl, sl, hl, blind_seed=0)
    n = 1 << 17
    assert np.array_equal(adv[:, :, : n - 6], adv0[:, :, : n - 6])
    assert not adv0[:, :, n - 6:].any()
    for p in range(2):
        for col in (0, 4):
            for row in (n - 6, n - 1):
                x = PL.blind_fe(seed, p, col, row) * O.MONT_R % O.R_MOD   # Montgomery limbs in memory
                got = sum(int(adv[p, col, row, j]) << (64 * j) for j in range(4))
                assert got == x
