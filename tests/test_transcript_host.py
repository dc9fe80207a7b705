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
