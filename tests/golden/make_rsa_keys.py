"""Generates the seeded synthetic RSA keys used by tests and bench.py:
key j of size `bits` = two primes from a seeded search (random.Random(0xB2000000 + j + bits)),
e = 65537.  Output: tests/golden/rsa_keys.json {bits: [{n, d}]}.  Deterministic; committed so
that the GPU box (no sympy dependency at run time, no /root/reference) uses the same keys.
    python tests/golden/make_rsa_keys.py
"""
import json
import os
import random

import sympy

HERE = os.path.dirname(os.path.abspath(__file__))
E = 65537


def gen_prime(rng, bits):
    while True:
        c = rng.getrandbits(bits) | (1 << (bits - 1)) | (1 << (bits - 2)) | 1
        if c % E == 1:
            continue
        if sympy.isprime(c):
            return c


def gen_key(bits, j):
    rng = random.Random(0xB2000000 + j + bits)
    while True:
        p, q = gen_prime(rng, bits // 2), gen_prime(rng, bits // 2)
        if p == q:
            continue
        n = p * q
        if n.bit_length() != bits:
            continue
        d = pow(E, -1, (p - 1) * (q - 1))
        return {"n": str(n), "d": str(d)}


if __name__ == "__main__":
    out = {}
    for bits, count in ((512, 2), (1024, 4), (2048, 8), (4096, 4)):
        out[str(bits)] = [gen_key(bits, j) for j in range(count)]
        print(bits, "done")
    json.dump(out, open(os.path.join(HERE, "rsa_keys.json"), "w"), indent=1)
