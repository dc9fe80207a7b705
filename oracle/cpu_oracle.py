"""TEST INFRASTRUCTURE ONLY - ctypes loader for the C oracle (oracle/*.c).

Used by tests/ as the full-size checker and by bench.py as the timed CPU baseline
("port": this repo's restatement of the reference's CPU algorithms; the Rust reference
cannot be built in this image).  Never imported by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    if force or not os.path.exists(LIB) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(LIB)
            for f in os.listdir(_HERE) if f.endswith((".c", ".h"))):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
    return _lib


def _p(a):
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


def ncores():
    return os.cpu_count() or 1


def sha256(msg: bytes) -> bytes:
    """oracle/sha256.c: FIPS 180-4 (the digest the reference takes from the `sha2` crate, benches/bench.rs:255-268)"""
    m = np.frombuffer(bytes(msg) or b"\0", dtype=np.uint8).copy()
    out = np.zeros(32, dtype=np.uint8)
    lib().orc_sha256(_p(m), C.c_uint64(len(msg)), _p(out))
    return out.tobytes()


def sha256_hashed_limbs(msg: bytes):
    """-> (digest bytes least significant first = the byte cells of src/lib.rs:210-211, the four 64-bit limbs of :222-236)"""
    m = np.frombuffer(bytes(msg) or b"\0", dtype=np.uint8).copy()
    d = np.zeros(32, dtype=np.uint8)
    l = np.zeros(4, dtype=np.uint64)
    lib().orc_sha256_hashed_limbs(_p(m), C.c_uint64(len(msg)), _p(d), _p(l))
    return d, l


def best_fft(a: np.ndarray, omega: np.ndarray, log_n: int, threads: int = 0) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64).copy()
    omega = np.ascontiguousarray(omega, dtype=np.uint64)
    lib().orc_best_fft(_p(a), _p(omega), C.c_uint(log_n), C.c_int(threads or ncores()))
    return a


def lagrange_to_coeff(a, k, threads=0):
    a = np.ascontiguousarray(a, dtype=np.uint64).copy()
    lib().orc_lagrange_to_coeff(_p(a), C.c_uint(k), C.c_int(threads or ncores()))
    return a


def coeff_to_extended(co, k, ext_k, threads=0):
    co = np.ascontiguousarray(co, dtype=np.uint64)
    out = np.empty((1 << ext_k, 4), dtype=np.uint64)
    lib().orc_coeff_to_extended(_p(co), C.c_uint(k), C.c_uint(ext_k), _p(out), C.c_int(threads or ncores()))
    return out


def extended_to_coeff(a, ext_k, threads=0):
    a = np.ascontiguousarray(a, dtype=np.uint64).copy()
    lib().orc_extended_to_coeff(_p(a), C.c_uint(ext_k), C.c_int(threads or ncores()))
    return a


def best_multiexp(scalars: np.ndarray, bases: np.ndarray, threads: int = 0) -> np.ndarray:
    """-> uint64[8] affine (Montgomery), identity = zeros"""
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    bases = np.ascontiguousarray(bases, dtype=np.uint64)
    n = scalars.shape[0]
    assert bases.shape[0] >= n
    out = np.zeros(8, dtype=np.uint64)
    lib().orc_best_multiexp(_p(scalars), _p(bases), C.c_size_t(n), C.c_int(threads or ncores()), _p(out))
    return out


def g1_multiples(n: int, threads: int = 0) -> np.ndarray:
    out = np.zeros((n, 8), dtype=np.uint64)
    lib().orc_g1_multiples(_p(out), C.c_size_t(n), C.c_int(threads or ncores()))
    return out


def g1_scalar_muls(scalars: np.ndarray, threads: int = 0) -> np.ndarray:
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    out = np.zeros((scalars.shape[0], 8), dtype=np.uint64)
    lib().orc_g1_scalar_muls(_p(scalars), _p(out), C.c_size_t(scalars.shape[0]), C.c_int(threads or ncores()))
    return out


def fr_sub_arrays(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """elementwise a - b over Fr on Montgomery arrays"""
    a = np.ascontiguousarray(a, dtype=np.uint64).copy()
    b = np.ascontiguousarray(b, dtype=np.uint64)
    assert a.shape == b.shape
    lib().orc_fr_sub_array(_p(a), _p(b), C.c_size_t(a.shape[0]))
    return a


# ---- hot path (a): witness synthesis oracle (oracle/rsa_witness.c) ---------------------------
class RsaTable:
    """One synthesized circuit table: the oracle's Circuit::synthesize for the bench circuit
    (reference benches/bench.rs:132-225) + a MockProver-style checker."""

    def __init__(self, bits_len: int, k: int):
        L = lib()
        L.orc_table_new.restype = C.c_void_p
        L.orc_table_rows.restype = C.c_uint64
        L.orc_check.restype = C.c_long
        self.k, self.bits_len, self.nl = k, bits_len, bits_len // 64
        self.h = C.c_void_p(L.orc_table_new(C.c_uint(k), C.c_int(64), C.c_int(self.nl)))

    def synthesize(self, n_limbs, sig_limbs, hash_limbs, e: int = 65537) -> int:
        """-> 1/0 = value of is_valid, -1 = the reference would have panicked"""
        e_le = np.frombuffer(e.to_bytes((e.bit_length() + 7) // 8, "little"), dtype=np.uint8).copy()
        n_limbs = np.ascontiguousarray(n_limbs, dtype=np.uint64)
        sig_limbs = np.ascontiguousarray(sig_limbs, dtype=np.uint64)
        hash_limbs = np.ascontiguousarray(hash_limbs, dtype=np.uint64)
        assert n_limbs.shape == (self.nl,) and sig_limbs.shape == (self.nl,) and hash_limbs.shape == (4,)
        return int(lib().orc_rsa_synthesize(self.h, C.c_int(self.bits_len), _p(e_le), C.c_int(e_le.size),
                                            _p(n_limbs), _p(sig_limbs), _p(hash_limbs)))

    def synthesize_digest(self, n_limbs, sig_limbs, hash_limbs, e: int = 65537) -> int:
        """RSASignatureVerifier::verify_pkcs1v15_signature from the digest bytes on (reference src/lib.rs:183-248);
        hash_limbs: the digest as four little-endian 64-bit limbs (their bytes are the 32 byte cells)"""
        e_le = np.frombuffer(e.to_bytes((e.bit_length() + 7) // 8, "little"), dtype=np.uint8).copy()
        n_limbs = np.ascontiguousarray(n_limbs, dtype=np.uint64)
        sig_limbs = np.ascontiguousarray(sig_limbs, dtype=np.uint64)
        digest_le = np.frombuffer(np.ascontiguousarray(hash_limbs, dtype="<u8").tobytes(), dtype=np.uint8).copy()
        assert digest_le.size == 32
        return int(lib().orc_rsa_synthesize_digest(self.h, C.c_int(self.bits_len), _p(e_le), C.c_int(e_le.size),
                                                   _p(n_limbs), _p(sig_limbs), _p(digest_le)))

    def synthesize_var(self, n_limbs, sig_limbs, hash_limbs, e: int, exp_limb_bits: int) -> int:
        """the same circuit with RSAPubE::Var: e is an assigned one-limb integer, exp_limb_bits of its bits are used"""
        n_limbs = np.ascontiguousarray(n_limbs, dtype=np.uint64)
        sig_limbs = np.ascontiguousarray(sig_limbs, dtype=np.uint64)
        hash_limbs = np.ascontiguousarray(hash_limbs, dtype=np.uint64)
        return int(lib().orc_rsa_synthesize_var(self.h, C.c_int(self.bits_len), C.c_int(exp_limb_bits), C.c_uint64(e),
                                                _p(n_limbs), _p(sig_limbs), _p(hash_limbs)))

    def advice(self) -> np.ndarray:
        out = np.empty((5, 1 << self.k, 4), dtype=np.uint64)
        lib().orc_table_advice(self.h, _p(out))
        return out

    def rows(self) -> int:
        return int(lib().orc_table_rows(self.h))

    def check(self):
        """-> (number of violated constraints, description of the first few)"""
        buf = C.create_string_buffer(4096)
        bad = lib().orc_check(self.h, buf, C.c_size_t(4096))
        return int(bad), buf.value.decode()

    def free(self):
        if self.h:
            lib().orc_table_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def int_to_limbs64(x: int, n: int) -> np.ndarray:
    return np.array([(x >> (64 * i)) & ((1 << 64) - 1) for i in range(n)], dtype=np.uint64)


def rsa_synthesize(bits_len: int, k: int, n: int, sig: int, hashed: int, e: int = 65537):
    """convenience: integers in -> (is_valid, advice uint64[5,2^k,4], rows, violations, msg)"""
    t = RsaTable(bits_len, k)
    nl = bits_len // 64
    v = t.synthesize(int_to_limbs64(n, nl), int_to_limbs64(sig, nl), int_to_limbs64(hashed, 4), e)
    bad, msg = t.check()
    out = (v, t.advice(), t.rows(), bad, msg)
    t.free()
    return out


# ---- single BigIntChip operations (the reference's impl_bigint_test_circuit! bodies) ------------
BIGINT_OPS = {"mul_kat": 0, "mul_mod": 1, "pow_mod_fixed_exp": 2, "add": 3, "sub": 4, "assert_in_field": 5,
              "refresh": 6, "add_mod": 7, "sub_mod": 8, "pow_mod": 9, "is_zero": 10, "is_equal_fresh": 11, "is_less_than": 12,
              "is_less_than_or_equal": 13, "is_greater_than": 14, "is_greater_than_or_equal": 15, "is_in_field": 16,
              "square": 17, "square_mod": 18}


def bigint_op(op: str, bits_len: int, k: int, a: int, b: int = 0, n: int = 0, exp_limb_bits: int = 5, with_advice: bool = False):
    """runs one BigIntChip operation in a fresh table, the way the reference's unit-test circuits do
    (src/big_integer/chip.rs:1470-3264).  -> (result limbs as ints or None if the reference would have
    panicked, number of violated constraints reported by the MockProver-style checker)"""
    nl = bits_len // 64
    t = RsaTable(bits_len, k)
    nw = 2 * nl if op == "mul_kat" else nl
    aw, bw = int_to_limbs64(a, nl), int_to_limbs64(b, nl)
    nwords = int_to_limbs64(n, nw)
    if op == "pow_mod":           # b = the exponent (one 64-bit word), the length argument carries exp_limb_bits
        nw = exp_limb_bits
    out = np.zeros((2 * nl + 2, 4), dtype=np.uint64)
    L = lib()
    L.orc_bigint_op.restype = C.c_int
    r = L.orc_bigint_op(t.h, C.c_int(BIGINT_OPS[op]), C.c_int(bits_len), _p(aw), _p(bw), _p(nwords), C.c_int(nw), _p(out))
    bad, _ = t.check()
    adv = t.advice() if with_advice else None
    t.free()
    if r < 0:
        return (None, bad, adv) if with_advice else (None, bad)
    limbs = [int(out[i, 0]) | (int(out[i, 1]) << 64) | (int(out[i, 2]) << 128) | (int(out[i, 3]) << 192) for i in range(r)]
    return (limbs, bad, adv) if with_advice else (limbs, bad)
