"""GPU: the SHA-256 front end of RSASignatureVerifier::verify_pkcs1v15_signature (reference src/lib.rs:183-248) at the
value level - message bytes hashed on the device (csrc/sha256.cu), the digest limbs fed to the recorded digest-tail
circuit, proofs created from the messages in one ABI call - against the oracle's independent SHA-256 restatement
(oracle/sha256.c, itself pinned to the FIPS vectors and hashlib on the CPU) and its row-by-row circuit synthesis."""
import hashlib
import random

import numpy as np
import pytest

import bn254 as O
import cpu_oracle as CO
import plonk as PL
import rsa_fixtures as RF
from util import fr_to_np, np_to_fr, np_to_g1

pytestmark = pytest.mark.gpu


def test_device_sha256_matches_oracle_and_hashlib(ctx):
    rnd = random.Random(2048)
    lens = list(range(0, 200)) + [247, 248, 255, 256, 257, 1000, 4097, 65536]
    msgs = [bytes(rnd.getrandbits(8) for _ in range(n)) for n in lens]
    msgs += [b"abc", b"", b"abcdbcdecdefdefgefghfghighijhijkijkljklmklmnlmnomnopnopq"]
    limbs, dig = ctx.sha256_batch(msgs)
    for i, m in enumerate(msgs):
        want = CO.sha256(m)
        assert want == hashlib.sha256(m).digest()
        assert bytes(dig[i]) == want, f"message {i} ({len(m)} bytes)"
        d_le, l = CO.sha256_hashed_limbs(m)
        assert np.array_equal(limbs[i], l), f"limbs of message {i}"
    assert bytes(dig[-3]).hex() == "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"   # FIPS 180-4 "abc"


def test_device_sha256_edge_cases(ctx):
    import b2rsa
    # a batch of empty messages (null-sized message buffer) and a large batch of ragged ones
    limbs, dig = ctx.sha256_batch([b"", b"", b""])
    assert all(bytes(d) == hashlib.sha256(b"").digest() for d in dig)
    msgs = [bytes([i & 0xff]) * (i % 131) for i in range(3000)]
    limbs, dig = ctx.sha256_batch(msgs)
    for i in (0, 1, 55, 56, 63, 64, 65, 119, 120, 130, 131, 2999):
        assert bytes(dig[i]) == hashlib.sha256(msgs[i]).digest()
    # decreasing offsets are refused
    offs = np.array([0, 4, 2], dtype=np.uint64)
    buf = np.zeros(8, dtype=np.uint8)
    out = np.zeros((2, 4), dtype=np.uint64)
    rc = ctx.lib.b2r_sha256_batch(ctx.h, buf.ctypes.data, offs.ctypes.data, 2, out.ctypes.data, None)
    assert rc == b2rsa.ERR_INVALID


def _signed(bits, count, tamper=(), wrong_key=()):
    """count (n, sig, msg) triples; for i in `tamper` the message that is hashed differs from the one that was signed,
    for i in `wrong_key` the public key is another key's modulus (one above the signature, so that the reference's
    assert_in_field passes and the verdict is a clean is_valid = 0) - the negative twins of src/lib.rs:541 and :626"""
    nl = bits // 64
    ks = RF.keys(bits)
    ns, ss, msgs = [], [], []
    for i in range(count):
        n, d = ks[i % len(ks)]
        msg = b"RSASignatureVerifier message %d " % i + bytes(range(i * 7 % 50))
        sig = pow(RF.emsa_pkcs1_v15(hashlib.sha256(msg).digest(), bits), d, n)
        if i in wrong_key:
            n = next(m for m, _ in ks if m != n and m > sig)
        ns.append(RF.limbs64(n, nl)); ss.append(RF.limbs64(sig, nl))
        msgs.append(msg + b"?" if i in tamper else msg)
    return np.stack(ns), np.stack(ss), msgs


def test_verifier_from_message_bytes_witness_and_proofs(ctx):
    """message -> device SHA-256 -> digest-tail circuit: every advice cell equals the oracle's table built from the
    oracle's own hash; create_proof from the messages in one call: status = is_valid per instance, the digests returned
    are the `hashed_bytes` of src/lib.rs:246-247, and every proof is accepted by the oracle verifier (the verifier
    circuit returns the bit instead of asserting it, so the tampered instance still has a satisfying witness)"""
    bits, k = 1024, 15
    nls, sls, msgs = _signed(bits, 4, tamper=(2,), wrong_key=(3,))
    prog = ctx.rsa_program_sha_tail(bits, k)
    limbs, dig = ctx.sha256_batch(msgs)
    adv, valid = prog.witness_batch(nls, sls, limbs)
    assert valid.tolist() == [1, 1, 0, 0]
    for i in range(4):
        _, l = CO.sha256_hashed_limbs(msgs[i])
        t = CO.RsaTable(bits, k)
        assert t.synthesize_digest(nls[i], sls[i], l) == int(valid[i])
        assert t.check()[0] == 0
        assert np.array_equal(t.advice(), adv[i]), f"instance {i}"
        t.free()
    g, gl = ctx.srs_setup(k, fr_to_np([O.srs_secret(k)])[0])
    pk = ctx.rsa_keygen(prog, g, gl)
    proofs, status, digests = pk.prove_msgs_batch(nls, sls, msgs, seed=0x5A, nonce=3)
    assert status.tolist() == [1, 1, 0, 0]
    for i in range(4):
        assert bytes(digests[i]) == hashlib.sha256(msgs[i]).digest()
    f, s_, t_ = pk.export_vk()
    vk = PL.vk_from_commitments(k, np_to_g1(f), np_to_g1(s_), np_to_fr(t_.reshape(1, 4))[0])
    for i in range(4):
        assert PL.verify_proof(vk, O.srs_secret(k), bytes(proofs[i])), f"proof {i}"
    # the same proofs as from pre-hashed limbs with the same key / nonce: the front end changes nothing downstream
    proofs2, status2 = pk.prove_batch(nls, sls, limbs, seed=0x5A, nonce=3)
    assert status2.tolist() == [1, 1, 0, 0] and np.array_equal(proofs, proofs2)
    pk.free(); g.free(); gl.free(); prog.free()


def test_bench_circuit_from_message_bytes(ctx):
    """the pkcs1v15 bench circuit (benches/bench.rs:132-225, sha2 disabled) asserts is_valid == 1: proving from the
    messages reports the tampered instance as status 0 and its proof is rejected"""
    bits, k = 512, 14
    nls, sls, msgs = _signed(bits, 2, tamper=(1,))
    prog = ctx.rsa_program(bits, k)
    g, gl = ctx.srs_setup(k, fr_to_np([O.srs_secret(k)])[0])
    pk = ctx.rsa_keygen(prog, g, gl)
    proofs, status, digests = pk.prove_msgs_batch(nls, sls, msgs, seed=bytes(range(32)), nonce=1)
    assert status.tolist() == [1, 0]
    f, s_, t_ = pk.export_vk()
    vk = PL.vk_from_commitments(k, np_to_g1(f), np_to_g1(s_), np_to_fr(t_.reshape(1, 4))[0])
    assert PL.verify_proof(vk, O.srs_secret(k), bytes(proofs[0]))
    assert not PL.verify_proof(vk, O.srs_secret(k), bytes(proofs[1]))
    # a key whose program takes more than the digest (RSAPubE::Var: 5 words) is refused
    pv = ctx.rsa_program_var(bits, 15, 5)
    g2, gl2 = ctx.srs_setup(15, fr_to_np([O.srs_secret(15)])[0])
    pkv = ctx.rsa_keygen(pv, g2, gl2)
    import b2rsa
    with pytest.raises(b2rsa.B2RError):
        pkv.prove_msgs_batch(nls, sls, msgs, seed=5)
    pkv.free(); g2.free(); gl2.free(); pv.free()
    pk.free(); g.free(); gl.free(); prog.free()
