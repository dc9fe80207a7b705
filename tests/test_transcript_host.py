"""CPU: the prover's host-side transcript (csrc/transcript.hpp: Blake2b-512 'Halo2-Transcript',
Challenge255, compressed points) compiled by g++ and compared with the oracle's hashlib transcript."""
import os
import subprocess

import bn254 as O
import plonk as P

HERE = os.path.dirname(os.path.abspath(__file__))


def test_transcript_matches_oracle(tmp_path):
    exe = str(tmp_path / "tht")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(HERE, "host", "transcript_host_test.cpp")])
    lines = subprocess.check_output([exe], text=True).split()
    t = P.Transcript()
    t.common_scalar(123456789)
    t.write_point((1, 2))
    c1 = t.squeeze()
    for i in range(40):
        t.write_scalar(c1 * (i + 7) % O.R_MOD)
    c2 = t.squeeze()
    c3 = t.squeeze()
    t.write_point(None)
    t.write_point((1, O.Q_MOD - 2))
    c4 = t.squeeze()
    assert [int(x, 16) for x in lines[:4]] == [c1, c2, c3, c4]
    assert bytes.fromhex(lines[4]) == bytes(t.out)
    assert P.decompress_point(bytes(t.out[-32:])) == (1, O.Q_MOD - 2)
    assert P.decompress_point(bytes(t.out[:32])) == (1, 2)


def test_blake2b_known_answer():
    """the hashlib construction the oracle uses is plain Blake2b-512 with a 16-byte personalisation"""
    import hashlib
    h = hashlib.blake2b(digest_size=64, person=b"Halo2-Transcript")
    h.update(b"\x00")
    assert len(h.digest()) == 64
    assert hashlib.blake2b(b"abc").hexdigest().startswith("ba80a53f981c4d0d6a2797b69f12f6e9")   # RFC 7693 appendix A


def test_blinding_stream_matches_oracle(tmp_path):
    """csrc/devutil.cuh blind_value (ChaCha20 block, from_bytes_wide reduction) against oracle/plonk.py blind_fe"""
    exe = str(tmp_path / "bht")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(HERE, "host", "blind_host_test.cpp")])
    cases = [(0xB200, 0, 0, 0, 131066), (0xB200, 0, 3, 40, 0), (7, 5, 63, 33, 131071), (0xFFFFFFFFFFFFFFFF, 1 << 40, 1000, 24, 16383),
             ("key", 9, 2, 8, 77), ("key", (1 << 64) - 1, 0xFFFFFFFF, 44, (1 << 24) - 1)]
    args = [str(x) for c in cases for x in c]
    lines = subprocess.check_output([exe] + args, text=True).split()
    for c, line in zip(cases, lines):
        seed = bytes(range(32)) if c[0] == "key" else c[0]
        assert int(line, 16) == P.blind_fe(seed, c[2], c[3], c[4], nonce=c[1]), c
    # RFC 7539 section 2.3.2 block-function vector through the oracle's vectorised ChaCha20
    import struct
    b = P._chacha20_blocks(struct.unpack("<8I", bytes(range(32))), [1], 0x09000000, 0x4a000000)
    assert [int(x) for x in b[0][:4]] == [0xe4e7f110, 0x15593bd1, 0x1fdd0f50, 0xc47120a3]
