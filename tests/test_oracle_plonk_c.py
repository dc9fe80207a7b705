"""CPU: the C restatement of create_proof (oracle/plonk_prover.c) pinned to the Python one (oracle/plonk.py), byte for
byte, at a size the Python loops finish in seconds (RSA-512, k = 14); the oracle verifier accepts its proofs, and a
witness that violates a range lookup is refused the way halo2's permute_expression_pair refuses it.  The C prover is
what the GPU parity tests compare the product's proof bytes with at k = 17 / 18 (tests/test_gpu_prover.py)."""
import numpy as np
import pytest

import cpu_oracle as CO
import plonk as PL
import rsa_fixtures as RF

BITS, K = 512, 14


@pytest.fixture(scope="module")
def small():
    srs = PL.Srs(K)
    opk = PL.keygen(PL.circuit_layout(BITS, K), srs)
    return srs, opk


def _advice(i):
    v, adv, rows, bad, msg = CO.rsa_synthesize(BITS, K, *RF.instance(BITS, i))
    assert v == 1 and bad == 0
    return adv


def test_c_prover_matches_python_prover(small):
    srs, opk = small
    adv = _advice(1)
    want = PL.create_proof(opk, srs, [PL.np_to_ints(adv[c]) for c in range(5)], 0xB200, proof_index=2, nonce=5)
    got, chal = PL.create_proof_c(opk["arrays"], srs, adv, 0xB200, proof_index=2, nonce=5)
    assert len(got) == PL.proof_length() == 2848
    if got != want:
        first = next(j for j in range(0, len(want), 32) if got[j:j + 32] != want[j:j + 32]) // 32
        raise AssertionError(f"proof element {first} of {len(want) // 32} differs")
    assert PL.verify_proof(opk, srs.s, got)
    # single-threaded and multi-threaded runs agree (chunked polynomial evaluation, threaded quotient)
    again, _ = PL.create_proof_c(opk["arrays"], srs, adv, 0xB200, proof_index=2, nonce=5, threads=1)
    assert again == got


def test_c_prover_randomness_and_key_forms(small):
    srs, opk = small
    adv = _advice(3)
    base, _ = PL.create_proof_c(opk["arrays"], srs, adv, 7, proof_index=0, nonce=0)
    seen = {base}
    for kw in ({"seed": 7, "proof_index": 1, "nonce": 0}, {"seed": 7, "proof_index": 0, "nonce": 1}, {"seed": 8, "proof_index": 0, "nonce": 0},
               {"seed": bytes(range(32)), "proof_index": 0, "nonce": 0}):
        p, _ = PL.create_proof_c(opk["arrays"], srs, adv, kw["seed"], proof_index=kw["proof_index"], nonce=kw["nonce"])
        assert p not in seen and PL.verify_proof(opk, srs.s, p)
        seen.add(p)
    # a caller-supplied transcript_repr (what a Rust host passes for the real vk) changes every challenge
    p, _ = PL.create_proof_c(opk["arrays"], srs, adv, 7, transcript_repr=12345)
    assert p != base and not PL.verify_proof(opk, srs.s, p)
    assert PL.verify_proof(dict(opk, transcript_repr=12345), srs.s, p)


def test_c_prover_rejects_out_of_table_lookup_input(small):
    srs, opk = small
    adv = _advice(0).copy()
    rng = PL.circuit_layout(BITS, K)["range"]
    row = int(np.nonzero(rng[0])[0][0])           # a row whose column a..d cells are range-checked sublimbs
    adv[0, row] = PL.ints_to_np([1 << 20])[0]     # far outside every table tag
    with pytest.raises(ValueError):
        PL.create_proof_c(opk["arrays"], srs, adv, 1)
    with pytest.raises(ValueError):
        PL.create_proof(opk, srs, [PL.np_to_ints(adv[c]) for c in range(5)], 1)
