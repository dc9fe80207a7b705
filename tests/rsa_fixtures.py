"""Synthetic RSA inputs shared by tests and bench.py: seeded keys (tests/golden/rsa_keys.json),
deterministic PKCS#1 v1.5 signatures of H_i = SHA-256(le64(i))."""
import hashlib
import json
import os
import struct

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
E = 65537
DIGEST_INFO_SHA256 = bytes.fromhex("3031300d060960864801650304020105000420")
_keys = None


def keys(bits: int):
    global _keys
    if _keys is None:
        _keys = json.load(open(os.path.join(_HERE, "golden", "rsa_keys.json")))
    return [(int(k["n"]), int(k["d"])) for k in _keys[str(bits)]]


def emsa_pkcs1_v15(h: bytes, bits: int) -> int:
    k = bits // 8
    t = DIGEST_INFO_SHA256 + h
    em = b"\x00\x01" + b"\xff" * (k - len(t) - 3) + b"\x00" + t
    return int.from_bytes(em, "big")


def instance(bits: int, i: int):
    """-> (n, sig, hashed) as integers; hashed = int(H_i) big-endian, as the reference feeds it
    (bench.rs:189-191 reverses the digest bytes and reads them little-endian)."""
    ks = keys(bits)
    n, d = ks[i % len(ks)]
    h = hashlib.sha256(struct.pack("<Q", i)).digest()
    sig = pow(emsa_pkcs1_v15(h, bits), d, n)
    return n, sig, int.from_bytes(h, "big")


def limbs64(x: int, n: int) -> np.ndarray:
    return np.array([(x >> (64 * j)) & ((1 << 64) - 1) for j in range(n)], dtype=np.uint64)


def batch(bits: int, count: int, start: int = 0):
    """-> (n_limbs[count, nl], sig_limbs[count, nl], hash_limbs[count, 4]) uint64"""
    nl = bits // 64
    ns, ss, hs = [], [], []
    for i in range(start, start + count):
        n, s, h = instance(bits, i)
        ns.append(limbs64(n, nl)); ss.append(limbs64(s, nl)); hs.append(limbs64(h, 4))
    return np.stack(ns), np.stack(ss), np.stack(hs)
