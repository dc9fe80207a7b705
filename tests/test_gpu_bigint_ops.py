"""GPU parity for the BigIntInstructions methods the pkcs1v15 bench circuit does not call - refresh, add_mod, sub_mod,
variable-exponent pow_mod (reference src/big_integer/chip.rs:168-233, 452-529, 664-696) - and for the pkcs1v15 circuit
with RSAPubE::Var (src/chip.rs:58-70, 99-114).  Each program is what the reference's own unit-test circuit builds
(chip.rs:1861-2271); it is recorded by the host mirror, replayed on the GPU through the C ABI and compared, all
5 x 2^k advice cells bit for bit, with the oracle's row-by-row restatement on the same inputs."""
import random

import numpy as np
import pytest

import bn254 as O
import cpu_oracle as CO
import plonk as PL
import rsa_fixtures as RF
from util import fr_to_np, np_to_fr, np_to_g1

pytestmark = pytest.mark.gpu


def _inputs(bits, seed, batch):
    r = random.Random(seed)
    n = r.getrandbits(bits) | (1 << (bits - 1)) | 1
    rows = []
    for i in range(batch):
        a, b = r.getrandbits(bits) % n, r.getrandbits(bits) % n
        if i == 1:
            b = a                      # sub_mod(a, a, n): the reference returns the unreduced representative n
        if i == 2:
            a, b = n - 1, n - 1
        rows.append((a, b, n, (5 * i + 17) % 32))
    return rows


@pytest.mark.parametrize("op", ["refresh", "add_mod", "sub_mod", "pow_mod"])
def test_bigint_op_advice_matches_oracle(ctx, op):
    bits, k, ebits, batch = 512, 15, 5, 4
    nl = bits // 64
    prog = ctx.bigint_program(op, bits, k, ebits)
    assert prog.aux_words == nl + 1
    cases = _inputs(bits, 1000 + len(op), batch)
    a_l = np.stack([CO.int_to_limbs64(a, nl) for a, _, _, _ in cases])
    b_l = np.stack([CO.int_to_limbs64(b, nl) for _, b, _, _ in cases])
    aux = np.stack([np.concatenate([CO.int_to_limbs64(n, nl), np.array([e], dtype=np.uint64)]) for _, _, n, e in cases])
    adv, _ = prog.witness_batch(a_l, b_l, aux)
    for i, (a, b, n, e) in enumerate(cases):
        limbs, bad, want = CO.bigint_op(op, bits, k, a, e if op == "pow_mod" else b, n, exp_limb_bits=ebits, with_advice=True)
        assert limbs is not None and bad == 0
        assert np.array_equal(adv[i], want), f"{op}: instance {i}"
        val = sum(l << (64 * j) for j, l in enumerate(limbs))
        if op == "refresh":
            assert val == a * b
        elif op == "add_mod":
            assert val % n == (a + b) % n
        elif op == "sub_mod":
            assert val % n == (a - b) % n
        else:
            assert val == pow(a, e, n)
    info = prog.info()
    assert info["rows_used"] <= (1 << k) - 6
    prog.free()


PREDICATES = {"is_zero": lambda a, b: a == 0, "is_equal_fresh": lambda a, b: a == b, "is_less_than": lambda a, b: a < b,
              "is_less_than_or_equal": lambda a, b: a <= b, "is_greater_than": lambda a, b: a > b,
              "is_greater_than_or_equal": lambda a, b: a >= b, "is_in_field": lambda a, b: a < b}


@pytest.mark.parametrize("op", list(PREDICATES) + ["square", "square_mod"])
def test_predicates_and_squares_match_oracle(ctx, op):
    """the comparison family (chip.rs:754-1006) and square / square_mod (:431-437, :642-649): advice bit-exact against
    the oracle, and the result cell (reported through is_valid for the predicates) against integer comparison"""
    bits, k, batch = 512, 15, 6
    nl = bits // 64
    prog = ctx.bigint_program(op, bits, k)
    r = random.Random(77)
    n = r.getrandbits(bits) | (1 << (bits - 1)) | 1
    a0 = r.getrandbits(bits) % n
    pairs = [(a0, r.getrandbits(bits) % n), (a0, a0), (a0 >> 128, a0), (a0, a0 >> 128), (0, 0), (0, 1)]
    a_l = np.stack([CO.int_to_limbs64(a, nl) for a, _ in pairs])
    b_l = np.stack([CO.int_to_limbs64(b, nl) for _, b in pairs])
    aux = np.stack([np.concatenate([CO.int_to_limbs64(n, nl), np.zeros(1, dtype=np.uint64)]) for _ in pairs])
    adv, valid = prog.witness_batch(a_l, b_l, aux)
    for i, (a, b) in enumerate(pairs):
        limbs, bad, want = CO.bigint_op(op, bits, k, a, b, n, with_advice=True)
        assert limbs is not None and bad == 0
        assert np.array_equal(adv[i], want), f"{op}: instance {i}"
        if op in PREDICATES:
            assert limbs == [int(PREDICATES[op](a, b))] and int(valid[i]) == limbs[0]
        elif op == "square_mod":
            assert sum(l << (64 * j) for j, l in enumerate(limbs)) == a * a % n
    prog.free()


def test_rsa_var_witness_and_proof(ctx):
    """RSAPubE::Var end to end at RSA-512 / k = 17 (17 exponent bits -> 34 mul_mod): advice equal to the oracle, a wrong
    exponent gives is_valid = 0, and a full proof over the Var circuit's own keygen verifies with the oracle verifier"""
    bits, k, ebits = 512, 17, 17
    nl = bits // 64
    prog = ctx.rsa_program_var(bits, k, ebits)
    assert prog.aux_words == 5
    n_l, s_l, h_l = RF.batch(bits, 2)
    aux = np.concatenate([h_l, np.array([[65537], [65539]], dtype=np.uint64)], axis=1)
    adv, valid = prog.witness_batch(n_l, s_l, aux)
    assert valid.tolist() == [1, 0]
    for i, e in enumerate((65537, 65539)):
        t = CO.RsaTable(bits, k)
        assert t.synthesize_var(n_l[i], s_l[i], h_l[i], e, ebits) == (1 if i == 0 else 0)
        assert np.array_equal(adv[i], t.advice()), f"instance {i}"
        t.free()
    # the same constraint system (MainGate + RangeChip), so keygen / create_proof apply unchanged
    g, gl = ctx.srs_setup(k, fr_to_np([O.srs_secret(k)])[0])
    pk = ctx.rsa_keygen(prog, g, gl)
    aux_ok = np.concatenate([h_l, np.full((2, 1), 65537, dtype=np.uint64)], axis=1)
    proofs, status = pk.prove_batch(n_l, s_l, aux_ok, seed=11)
    assert status.tolist() == [1, 1]
    f, s_, t_ = pk.export_vk()
    vk = PL.vk_from_commitments(k, np_to_g1(f), np_to_g1(s_), np_to_fr(t_.reshape(1, 4))[0])
    assert PL.verify_proof(vk, O.srs_secret(k), bytes(proofs[0]))
    assert PL.verify_proof(vk, O.srs_secret(k), bytes(proofs[1]))
    pk.free(); g.free(); gl.free(); prog.free()


def test_sha_tail_witness_matches_oracle(ctx):
    """RSASignatureVerifier::verify_pkcs1v15_signature from the digest bytes on (reference src/lib.rs:183-248) on the GPU:
    every advice cell equal to the oracle's table for valid signatures and for a corrupted digest (is_valid = 0)"""
    import rsa_fixtures as RF
    bits, k = 2048, 17
    nl = bits // 64
    prog = ctx.rsa_program_sha_tail(bits, k)
    nls, sls, hls = RF.batch(bits, 3, start=20)
    hls[2, 3] ^= np.uint64(1 << 7)
    adv, valid = prog.witness_batch(nls, sls, hls)
    assert valid.tolist() == [1, 1, 0]
    for i in range(3):
        t = CO.RsaTable(bits, k)
        assert t.synthesize_digest(nls[i], sls[i], hls[i]) == int(valid[i])
        assert t.check()[0] == 0
        assert np.array_equal(t.advice(), adv[i]), f"instance {i}"
        t.free()
    assert prog.info()["rows_used"] < (1 << k) - 6
    prog.free()
