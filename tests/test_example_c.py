"""The C ABI from plain C: examples/rsa_example.c (the reference's examples/rsa_example.rs flow: message bytes +
pkcs1v15 signature + public key -> proof) compiles as C11 against include/b2rsa.h, links against libb2rsa.so alone,
refuses to run without a device (CPU), and on a B200 produces proofs that the oracle verifier accepts (GPU)."""
import hashlib
import os
import struct
import subprocess

import numpy as np
import pytest

import rsa_fixtures as RF

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "halo2-rsa_b200", "lib")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    if not os.path.exists(os.path.join(LIBDIR, "libb2rsa.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "halo2-rsa_b200"), "-j8", "-s"])
    out = str(tmp_path_factory.mktemp("example") / "rsa_example")
    subprocess.check_call(["gcc", "-O2", "-std=c11", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "rsa_example.c"), "-L" + LIBDIR, "-lb2rsa", "-Wl,-rpath," + LIBDIR, "-o", out])
    return out


def test_example_builds_and_has_no_cpu_fallback(exe, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the loud-failure path is exercised on CPU boxes")
    p = subprocess.run([exe, "1024", "15", os.devnull, str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert p.returncode == 1 and "no CPU fallback" in p.stderr


@pytest.mark.gpu
def test_example_proves_from_message_bytes(exe, tmp_path):
    import bn254 as O
    import plonk as PL
    from util import np_to_fr, np_to_g1
    bits, k = 1024, 15
    ks = RF.keys(bits)
    lines, msgs = [], []
    for i in range(3):
        n, d = ks[i % len(ks)]
        msg = b"examples/rsa_example.c instance %d" % i
        sig = pow(RF.emsa_pkcs1_v15(hashlib.sha256(msg).digest(), bits), d, n)
        sent = msg if i != 1 else msg + b"x"          # instance 1: the message is not the one that was signed
        msgs.append(sent)
        lines.append("%x %x %s" % (n, sig, sent.hex()))
    inp, out = tmp_path / "in.txt", tmp_path / "out.bin"
    inp.write_text("\n".join(lines) + "\n")
    p = subprocess.run([exe, str(bits), str(k), str(inp), str(out)], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "valid=2" in p.stdout
    raw = out.read_bytes()
    batch, pb, nf, ns = struct.unpack_from("<4I", raw, 0)
    assert batch == 3 and pb == 2848
    o = 16
    status = list(raw[o:o + batch]); o += batch
    digests = [raw[o + 32 * i:o + 32 * (i + 1)] for i in range(batch)]; o += 32 * batch
    proofs = [raw[o + pb * i:o + pb * (i + 1)] for i in range(batch)]; o += pb * batch
    fixed = np.frombuffer(raw[o:o + nf * 64], dtype=np.uint64).reshape(nf, 8); o += nf * 64
    sigma = np.frombuffer(raw[o:o + ns * 64], dtype=np.uint64).reshape(ns, 8); o += ns * 64
    repr_ = np.frombuffer(raw[o:o + 32], dtype=np.uint64).reshape(1, 4)
    assert status == [1, 0, 1]
    for i in range(batch):
        assert digests[i] == hashlib.sha256(msgs[i]).digest()
    # the example's SRS secret is 0xB200 (its Montgomery limbs are spelled out in the C source)
    secret = 0xB200
    vk = PL.vk_from_commitments(k, np_to_g1(fixed), np_to_g1(sigma), np_to_fr(repr_)[0])
    for i in range(batch):
        assert PL.verify_proof(vk, secret, proofs[i]), f"proof {i}"
