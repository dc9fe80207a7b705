"""GPU parity: keygen and the full prover (b2r_rsa_keygen / b2r_rsa_prove_batch, through the C ABI) against the
oracle's independent restatement of halo2's keygen_vk / keygen_pk / create_proof / verify_proof (oracle/plonk.py).

  * RSA-512 at k = 14 (a size the Python oracle proves in seconds): verifying key and PROOF BYTES are identical
    for the same seeded randomness - every commitment, challenge, evaluation and opening witness of the proof.
  * RSA-2048 at k = 17 (BASELINE configs[0]/[1] size): the proofs are accepted by the oracle verifier, a tampered
    proof and a proof for a wrong signature are rejected.
"""
import numpy as np
import pytest

import bn254 as O
import cpu_oracle as CO
import plonk as PL
import rsa_fixtures as RF
from util import fr_to_np, np_to_fr, np_to_g1

pytestmark = pytest.mark.gpu


def _setup(ctx, bits, k):
    prog = ctx.rsa_program(bits, k)
    g, gl = ctx.srs_setup(k, fr_to_np([O.srs_secret(k)])[0])
    pk = ctx.rsa_keygen(prog, g, gl)
    return prog, g, gl, pk


def _vk(pk):
    f, s, t = pk.export_vk()
    return PL.vk_from_commitments(pk.k, np_to_g1(f), np_to_g1(s), np_to_fr(t.reshape(1, 4))[0])


@pytest.fixture(scope="module")
def small(ctx):
    bits, k = 512, 14
    prog, g, gl, pk = _setup(ctx, bits, k)
    srs = PL.Srs(k)
    opk = PL.keygen(PL.circuit_layout(bits, k), srs)
    yield bits, k, pk, srs, opk
    pk.free(); g.free(); gl.free(); prog.free()


def test_keygen_matches_oracle(small):
    bits, k, pk, srs, opk = small
    vk = _vk(pk)
    assert pk.proof_bytes == PL.proof_length() == 2848
    assert vk["fixed_commitments"] == opk["fixed_commitments"]
    assert vk["sigma_commitments"] == opk["sigma_commitments"]
    assert vk["transcript_repr"] == opk["transcript_repr"]


def test_proof_bytes_match_oracle(small):
    bits, k, pk, srs, opk = small
    seed, batch = 0xB200, 2
    nl, sl, hl = RF.batch(bits, batch, start=1)
    proofs, status = pk.prove_batch(nl, sl, hl, seed)
    assert status.tolist() == [1] * batch
    for i in range(batch):
        v, adv, rows, bad, msg = CO.rsa_synthesize(bits, k, *RF.instance(bits, 1 + i))
        assert v == 1 and bad == 0
        want = PL.create_proof(opk, srs, [PL.np_to_ints(adv[c]) for c in range(5)], seed, proof_index=i)
        got = bytes(proofs[i])
        if got != want:   # name the first differing proof element
            first = next(j for j in range(0, len(want), 32) if got[j:j + 32] != want[j:j + 32]) // 32
            raise AssertionError(f"instance {i}: proof element {first} of {len(want) // 32} differs")
        assert PL.verify_proof(opk, srs.s, got)


def test_full_size_proofs_verify(ctx):
    bits, k = 2048, 17
    prog, g, gl, pk = _setup(ctx, bits, k)
    vk = _vk(pk)
    secret = O.srs_secret(k)
    nl, sl, hl = RF.batch(bits, 3)
    hl_bad = hl.copy()
    hl_bad[2, 0] ^= np.uint64(1)                      # third instance: wrong message hash
    proofs, status = pk.prove_batch(nl, sl, hl_bad, seed=7)
    assert status.tolist() == [1, 1, 0]
    assert PL.verify_proof(vk, secret, bytes(proofs[0]))
    assert PL.verify_proof(vk, secret, bytes(proofs[1]))
    assert not PL.verify_proof(vk, secret, bytes(proofs[2]))          # is_valid = 0 violates assert_one
    assert bytes(proofs[0]) != bytes(proofs[1])
    bad = bytearray(proofs[0])
    bad[32 * 40 + 3] ^= 1                                              # one evaluation
    assert not PL.verify_proof(vk, secret, bytes(bad))
    # same seed, same inputs -> same bytes; other seed -> other blinding -> other proof, still valid
    again, _ = pk.prove_batch(nl[:1], sl[:1], hl[:1], seed=7)
    assert bytes(again[0]) == bytes(proofs[0])
    other, _ = pk.prove_batch(nl[:1], sl[:1], hl[:1], seed=8)
    assert bytes(other[0]) != bytes(proofs[0]) and PL.verify_proof(vk, secret, bytes(other[0]))
    pk.free(); g.free(); gl.free(); prog.free()


@pytest.mark.parametrize("bits,k", [(1024, 15), (4096, 18)])
def test_other_key_sizes_verify(ctx, bits, k):
    """BASELINE configs[2] (RSA-4096, k=18) and the reference's own enabled bench size (RSA-1024, k=15)"""
    prog, g, gl, pk = _setup(ctx, bits, k)
    vk = _vk(pk)
    nl, sl, hl = RF.batch(bits, 2)
    proofs, status = pk.prove_batch(nl, sl, hl, seed=99)
    assert status.tolist() == [1, 1]
    for i in range(2):
        assert PL.verify_proof(vk, O.srs_secret(k), bytes(proofs[i]))
    pk.free(); g.free(); gl.free(); prog.free()


def test_prover_argument_errors(small):
    import b2rsa
    bits, k, pk, srs, opk = small
    nl, sl, hl = RF.batch(bits, 1)
    with pytest.raises(b2rsa.B2RError):
        pk.prove_batch(nl, sl, hl, seed=0)
