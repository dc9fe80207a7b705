"""CPU: the oracle's SHA-256 restatement (oracle/sha256.c) pinned against the FIPS 180-4 example vectors and Python's
hashlib - the digest that the reference's RSASignatureVerifier (src/lib.rs:204-211) feeds to the RSA check and that its
bench / tests compute with the `sha2` crate (benches/bench.rs:255-268)."""
import hashlib
import random

import numpy as np

import cpu_oracle as CO
import rsa_fixtures as RF

# FIPS 180-4 / NIST CSRC example vectors
KATS = [
    (b"abc", "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"),
    (b"", "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855"),
    (b"abcdbcdecdefdefgefghfghighijhijkijkljklmklmnlmnomnopnopq", "248d6a61d20638b8e5c026930c3e6039a33ce45964ff2167f6ecedd419db06c1"),
    (b"a" * 1000000, "cdc76e5c9914fb9281a1c7e284d73e67f1809a48a497200e046d39ccc7112cd0"),
]


def test_fips_vectors():
    for msg, want in KATS:
        assert CO.sha256(msg).hex() == want


def test_padding_boundaries_against_hashlib():
    rnd = random.Random(7)
    for n in list(range(0, 200)) + [247, 248, 255, 256, 1000, 4097]:
        m = bytes(rnd.getrandbits(8) for _ in range(n))
        assert CO.sha256(m) == hashlib.sha256(m).digest(), n


def test_hashed_limbs_are_what_the_circuit_consumes():
    """src/lib.rs:210-236: reversed digest bytes -> four limbs = the big-endian digest integer in 64-bit limbs, which is
    how tests/rsa_fixtures.py (and benches/bench.rs:189-191 for the sha-disabled circuit) present the hash"""
    for i in range(5):
        msg = b"message %d" % i
        d_le, limbs = CO.sha256_hashed_limbs(msg)
        h = hashlib.sha256(msg).digest()
        assert bytes(d_le) == h[::-1]
        assert np.array_equal(limbs, RF.limbs64(int.from_bytes(h, "big"), 4))


def test_message_level_verifier_on_the_oracle_table():
    """message -> digest limbs -> orc_rsa_synthesize_digest: valid for the signed message, 0 for any other message,
    constraints satisfied either way (the verifier returns the bit, src/lib.rs:245)"""
    bits, k = 1024, 15
    nl = bits // 64
    n, d = RF.keys(bits)[0]
    msg = b"halo2-rsa message-level check"
    sig = pow(RF.emsa_pkcs1_v15(hashlib.sha256(msg).digest(), bits), d, n)
    for m, want in ((msg, 1), (msg + b"!", 0)):
        _, limbs = CO.sha256_hashed_limbs(m)
        t = CO.RsaTable(bits, k)
        assert t.synthesize_digest(RF.limbs64(n, nl), RF.limbs64(sig, nl), limbs) == want
        assert t.check()[0] == 0
        t.free()


def test_device_sha256_source_host_build_matches_hashlib(tmp_path):
    """the DEVICE implementation (csrc/sha256.cuh, what k_sha256_msgs runs) compiled for the host by g++ - like the host
    build of field.cuh - against hashlib for every padding boundary, and its limbs against the oracle's"""
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / "sha_host")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(here, "host", "sha256_host_test.cpp")])
    rnd = random.Random(11)
    msgs = [bytes(rnd.getrandbits(8) for _ in range(n)) for n in list(range(0, 200)) + [247, 248, 255, 256, 1000, 4097]]
    msgs += [m for m, _ in KATS[:3]]
    inp = "\n".join(m.hex() if m else "-" for m in msgs) + "\n"
    out = subprocess.run([exe], input=inp, capture_output=True, text=True, check=True).stdout.splitlines()
    assert len(out) == len(msgs)
    for m, line in zip(msgs, out):
        f = line.split()
        assert f[0] == hashlib.sha256(m).hexdigest(), len(m)
        _, limbs = CO.sha256_hashed_limbs(m)
        assert [int(x, 16) for x in f[1:5]] == [int(v) for v in limbs]
