"""b2rsa - host-side binding of libb2rsa.so (hand-written sm_100a kernels behind a C ABI).

This package is plumbing: it loads the in-tree shared library with ctypes and exposes the
entry points of include/b2rsa.h on numpy arrays (host buffers) and raw device pointers
(e.g. ``torch.Tensor.data_ptr()``).  There is no CPU fallback: if the library is missing
or no sm_100 device is present, construction raises.

Element format everywhere: halo2curves bn256 memory format - ``uint64[..., 4]``
little-endian limbs in Montgomery form; G1Affine = ``uint64[..., 8]`` (x, y).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2R_LIB") or os.path.join(_HERE, "..", "lib", "libb2rsa.so")   # B2R_LIB: A/B builds during kernel work

OK = 0
ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_NOMEM, ERR_LAYOUT, ERR_SYNTH = -1, -2, -3, -4, -5, -6


class B2RError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libb2rsa error {code}: {msg}")
        self.code = code


def load_library(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} not built: run `make -C halo2-rsa_b200` (or __graft_entry__.build()); "
            "b2rsa has no CPU fallback")
    lib = C.CDLL(os.path.abspath(path))
    vp, u32, u64, i32, sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32, C.c_size_t
    sigs = {
        "b2r_ctx_create": [i32, C.POINTER(vp)],
        "b2r_ctx_destroy": [vp],
        "b2r_ctx_set_stream": [vp, vp],
        "b2r_ctx_sync": [vp],
        "b2r_dev_alloc": [vp, sz, C.POINTER(vp)],
        "b2r_dev_free": [vp, vp],
        "b2r_h2d": [vp, vp, vp, sz],
        "b2r_d2h": [vp, vp, vp, sz],
        "b2r_ntt_fr": [vp, vp, vp, u32],
        "b2r_ntt_fr_dev": [vp, vp, vp, u32],
        "b2r_ntt_fr_batch_dev": [vp, vp, sz, vp, u32],
        "b2r_intt_fr": [vp, vp, u32],
        "b2r_intt_fr_batch_dev": [vp, vp, sz, u32],
        "b2r_coset_ntt_fr": [vp, vp, u32, u32, vp],
        "b2r_coset_ntt_fr_batch_dev": [vp, vp, sz, u32, u32, vp],
        "b2r_coset_intt_fr": [vp, vp, u32],
        "b2r_coset_intt_fr_batch_dev": [vp, vp, sz, u32],
        "b2r_bases_register": [vp, vp, sz, C.POINTER(vp)],
        "b2r_bases_free": [vp, vp],
        "b2r_msm_g1": [vp, vp, vp, sz, vp],
        "b2r_msm_g1_batch": [vp, vp, vp, sz, sz, vp],
        "b2r_msm_g1_batch_dev": [vp, vp, vp, sz, sz, vp],
        "b2r_msm_g1_batch_dev_ex": [vp, vp, vp, sz, sz, u32, vp],
        "b2r_profile_enable": [vp, i32],
        "b2r_profile_read": [vp, C.c_char_p, C.POINTER(C.c_double), C.POINTER(u64), C.POINTER(C.c_double)],
        "b2r_profile_dump": [vp, C.c_char_p, sz, i32],
        "b2r_bases_download": [vp, vp, vp, sz],
        "b2r_srs_setup": [vp, u32, vp, C.POINTER(vp), C.POINTER(vp)],
        "b2r_rsa_commit_batch": [vp, vp, vp, vp, vp, vp, sz, u64, u32, u32, vp, vp, vp, vp],
        "b2r_rsa_commit_batch_dev": [vp, vp, vp, vp, vp, vp, sz, u64, u32, u32, vp, vp, vp, vp],
        "b2r_rsa_program_build": [vp, u32, vp, sz, u32, C.POINTER(vp)],
        "b2r_rsa_program_build_var": [vp, u32, u32, u32, C.POINTER(vp)],
        "b2r_rsa_program_build_sha_tail": [vp, u32, vp, sz, u32, C.POINTER(vp)],
        "b2r_bigint_program_build": [vp, u32, u32, u32, u32, C.POINTER(vp)],
        "b2r_prog_aux_words": [vp],
        "b2r_prog_free": [vp, vp],
        "b2r_prog_info": [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)],
        "b2r_rsa_witness_batch": [vp, vp, vp, vp, vp, sz, u64, vp, vp],
        "b2r_rsa_witness_batch_dev": [vp, vp, vp, vp, vp, sz, u64, vp, vp],
        "b2r_rsa_keygen": [vp, vp, vp, vp, C.POINTER(vp)],
        "b2r_pk_free": [vp, vp],
        "b2r_pk_info": [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(u64)],
        "b2r_pk_export_vk": [vp, vp, vp, vp],
        "b2r_rsa_prove_batch": [vp, vp, vp, vp, vp, sz, u64, vp, vp],
        "b2r_rsa_prove_batch_dev": [vp, vp, vp, vp, vp, sz, u64, vp, vp],
        "b2r_rsa_prove_batch_ex": [vp, vp, vp, vp, vp, sz, vp, u64, u32, vp, vp],
        "b2r_pk_set_transcript_repr": [vp, vp],
        "b2r_field_selftest": [vp, u32, u32, vp, vp, vp, vp, vp, sz],
        "b2r_last_commitments": [vp, vp, sz, u32, C.POINTER(sz), C.POINTER(u32)],
        "b2r_sha256_batch": [vp, vp, vp, sz, vp, vp],
        "b2r_sha256_batch_dev": [vp, vp, vp, sz, vp, vp],
        "b2r_rsa_prove_msgs_batch": [vp, vp, vp, vp, vp, vp, sz, vp, u64, u32, vp, vp, vp],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)  # AttributeError if the ABI symbol is missing: loud by design
        fn.argtypes = args
        fn.restype = i32
    lib.b2r_prog_num_limbs.argtypes = [vp]
    lib.b2r_prog_num_limbs.restype = i32
    lib.b2r_last_error.argtypes = [vp]
    lib.b2r_last_error.restype = C.c_char_p
    lib.b2r_version.argtypes = []
    lib.b2r_version.restype = C.c_char_p
    lib.b2r_launch_count.argtypes = [vp]
    lib.b2r_launch_count.restype = u64
    return lib


def _host_ptr(a: np.ndarray) -> C.c_void_p:
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


def _fr_array(a, shape_tail=4) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    assert a.shape[-1] == shape_tail, a.shape
    return a


class Context:
    """One device, one stream (mirrors b2r_ctx)."""

    def __init__(self, device: int = 0, lib: C.CDLL | None = None):
        self.lib = lib or load_library()
        h = C.c_void_p()
        rc = self.lib.b2r_ctx_create(device, C.byref(h))
        if rc != OK:
            raise B2RError(rc, self.lib.b2r_last_error(None).decode())
        self.h = h
        self.device = device

    # -- plumbing
    def _ck(self, rc: int):
        if rc != OK:
            raise B2RError(rc, self.lib.b2r_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.b2r_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int | None):
        self._ck(self.lib.b2r_ctx_set_stream(self.h, C.c_void_p(cuda_stream or 0)))

    def sync(self):
        self._ck(self.lib.b2r_ctx_sync(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.b2r_launch_count(self.h))

    def profile_enable(self, on: bool):
        self._ck(self.lib.b2r_profile_enable(self.h, 1 if on else 0))

    def profile_read(self, name: str):
        """-> (total device ms, launches, work units) recorded under one kernel name since the last clear"""
        ms, cnt, units = C.c_double(), C.c_uint64(), C.c_double()
        self._ck(self.lib.b2r_profile_read(self.h, name.encode(), C.byref(ms), C.byref(cnt), C.byref(units)))
        return ms.value, cnt.value, units.value

    def profile_dump(self, clear: bool = True) -> dict:
        """-> {kernel name: (total device ms, launches)} since the last clear"""
        buf = C.create_string_buffer(1 << 16)
        self._ck(self.lib.b2r_profile_dump(self.h, buf, 1 << 16, 1 if clear else 0))
        out = {}
        for item in buf.value.decode().split(";"):
            if item:
                name, ms, cnt = item.split(":")
                out[name] = (float(ms), int(cnt))
        return out

    def dev_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._ck(self.lib.b2r_dev_alloc(self.h, nbytes, C.byref(p)))
        return p.value

    def dev_free(self, dptr: int):
        self._ck(self.lib.b2r_dev_free(self.h, C.c_void_p(dptr)))

    def h2d(self, dptr: int, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        self._ck(self.lib.b2r_h2d(self.h, C.c_void_p(dptr), _host_ptr(arr), arr.nbytes))

    def d2h(self, arr: np.ndarray, dptr: int):
        self._ck(self.lib.b2r_d2h(self.h, _host_ptr(arr), C.c_void_p(dptr), arr.nbytes))

    # -- NTT (best_fft and the EvaluationDomain wrappers)
    def ntt(self, a: np.ndarray, omega: np.ndarray, log_n: int) -> np.ndarray:
        """best_fft(a, omega, log_n): returns the transformed copy (natural order)."""
        a = _fr_array(a).copy()
        assert a.shape == (1 << log_n, 4)
        omega = _fr_array(omega.reshape(4))
        self._ck(self.lib.b2r_ntt_fr(self.h, _host_ptr(a), _host_ptr(omega), log_n))
        return a

    def intt(self, a: np.ndarray, k: int) -> np.ndarray:
        a = _fr_array(a).copy()
        assert a.shape == (1 << k, 4)
        self._ck(self.lib.b2r_intt_fr(self.h, _host_ptr(a), k))
        return a

    def coset_ntt(self, coeffs: np.ndarray, k: int, ext_k: int) -> np.ndarray:
        coeffs = _fr_array(coeffs)
        assert coeffs.shape == (1 << k, 4)
        out = np.empty((1 << ext_k, 4), dtype=np.uint64)
        self._ck(self.lib.b2r_coset_ntt_fr(self.h, _host_ptr(coeffs), k, ext_k, _host_ptr(out)))
        return out

    def coset_intt(self, a: np.ndarray, ext_k: int) -> np.ndarray:
        a = _fr_array(a).copy()
        assert a.shape == (1 << ext_k, 4)
        self._ck(self.lib.b2r_coset_intt_fr(self.h, _host_ptr(a), ext_k))
        return a

    def ntt_batch_dev(self, dptr: int, batch: int, omega: np.ndarray, log_n: int):
        omega = _fr_array(omega.reshape(4))
        self._ck(self.lib.b2r_ntt_fr_batch_dev(self.h, C.c_void_p(dptr), batch, _host_ptr(omega), log_n))

    def intt_batch_dev(self, dptr: int, batch: int, k: int):
        self._ck(self.lib.b2r_intt_fr_batch_dev(self.h, C.c_void_p(dptr), batch, k))

    def coset_ntt_batch_dev(self, src: int, batch: int, k: int, ext_k: int, dst: int):
        self._ck(self.lib.b2r_coset_ntt_fr_batch_dev(self.h, C.c_void_p(src), batch, k, ext_k, C.c_void_p(dst)))

    def coset_intt_batch_dev(self, dptr: int, batch: int, ext_k: int):
        self._ck(self.lib.b2r_coset_intt_fr_batch_dev(self.h, C.c_void_p(dptr), batch, ext_k))

    # -- MSM (best_multiexp through ParamsKZG::commit / commit_lagrange)
    def bases_register(self, bases: np.ndarray) -> "Bases":
        bases = _fr_array(bases, 8)
        h = C.c_void_p()
        self._ck(self.lib.b2r_bases_register(self.h, _host_ptr(bases), bases.shape[0], C.byref(h)))
        return Bases(self, h, bases.shape[0])

    def msm(self, bases: "Bases", scalars: np.ndarray) -> np.ndarray:
        """best_multiexp(scalars, bases[:len(scalars)]) -> uint64[12] Jacobian (z=1 or 0)."""
        scalars = _fr_array(scalars)
        out = np.zeros(12, dtype=np.uint64)
        self._ck(self.lib.b2r_msm_g1(self.h, bases.h, _host_ptr(scalars), scalars.shape[0], _host_ptr(out)))
        return out

    def msm_batch(self, bases: "Bases", scalars: np.ndarray) -> np.ndarray:
        """scalars uint64[m, n, 4] -> uint64[m, 8] affine commitments."""
        scalars = _fr_array(scalars)
        m, n = scalars.shape[0], scalars.shape[1]
        out = np.zeros((m, 8), dtype=np.uint64)
        self._ck(self.lib.b2r_msm_g1_batch(self.h, bases.h, _host_ptr(scalars), m, n, _host_ptr(out)))
        return out

    MSM_UNIFORM = 1

    def msm_batch_dev(self, bases: "Bases", scalars_dptr: int, m: int, n: int, out_dptr: int, uniform: bool = False):
        """device-resident scalars / outputs; uniform=True passes B2R_MSM_UNIFORM (binned counting sort, same result)"""
        if uniform:
            self._ck(self.lib.b2r_msm_g1_batch_dev_ex(self.h, bases.h, C.c_void_p(scalars_dptr), m, n, self.MSM_UNIFORM, C.c_void_p(out_dptr)))
        else:
            self._ck(self.lib.b2r_msm_g1_batch_dev(self.h, bases.h, C.c_void_p(scalars_dptr), m, n, C.c_void_p(out_dptr)))

    FIELD_OPS = {"mul": 0, "sqr": 1, "mul_add_mul": 2, "mul_sub_mul": 3, "dot4": 4, "add": 5, "sub": 6, "inv": 7}

    def field_selftest(self, field: str, op: str, a, b=None, c=None, d=None) -> np.ndarray:
        """diagnostic: the device field arithmetic (csrc/field.cuh) on uint64[n,4] Montgomery operands"""
        a = _fr_array(a)
        ops = [None if x is None else _fr_array(x) for x in (b, c, d)]
        out = np.empty_like(a)
        ptr = lambda x: C.c_void_p(None) if x is None else _host_ptr(x)
        self._ck(self.lib.b2r_field_selftest(self.h, {"fr": 0, "fq": 1}[field], self.FIELD_OPS[op], _host_ptr(a), ptr(ops[0]), ptr(ops[1]),
                                             ptr(ops[2]), _host_ptr(out), a.shape[0]))
        return out

    def srs_setup(self, k: int, secret: np.ndarray):
        """ParamsKZG::setup(k) for a chosen secret (uint64[4] Montgomery) -> (g, g_lagrange)"""
        secret = _fr_array(np.asarray(secret).reshape(4))
        g, gl = C.c_void_p(), C.c_void_p()
        self._ck(self.lib.b2r_srs_setup(self.h, k, _host_ptr(secret), C.byref(g), C.byref(gl)))
        return Bases(self, g, 1 << k), Bases(self, gl, 1 << k)

    def rsa_commit_batch(self, prog: "RsaProgram", g_lagrange: "Bases", n_limbs, sig_limbs, hash_limbs, ext_k: int,
                         advice_dptr: int, ext_dptr: int, blind_seed: int = 0):
        """fused hot path, HOST inputs -> (commitments uint64[batch,5,8], is_valid uint8[batch])"""
        n_limbs = np.ascontiguousarray(n_limbs, dtype=np.uint64)
        sig_limbs = np.ascontiguousarray(sig_limbs, dtype=np.uint64)
        hash_limbs = np.ascontiguousarray(hash_limbs, dtype=np.uint64)
        batch = n_limbs.shape[0]
        cm = np.zeros((batch, 5, 8), dtype=np.uint64)
        valid = np.zeros(batch, dtype=np.uint8)
        self._ck(self.lib.b2r_rsa_commit_batch(self.h, prog.h, g_lagrange.h, _host_ptr(n_limbs), _host_ptr(sig_limbs),
                                               _host_ptr(hash_limbs), batch, blind_seed, prog.k, ext_k,
                                               C.c_void_p(advice_dptr), C.c_void_p(ext_dptr or 0), _host_ptr(cm), _host_ptr(valid)))
        return cm, valid

    def rsa_commit_batch_raw(self, prog, g_lagrange, n_ptr: int, s_ptr: int, h_ptr: int, batch: int, ext_k: int,
                             advice_dptr: int, ext_dptr: int, cm_ptr: int, valid_ptr: int, blind_seed: int = 0, host: bool = False):
        """same, raw pointers (pinned host buffers when host=True, device pointers otherwise); asynchronous for host=False"""
        fn = self.lib.b2r_rsa_commit_batch if host else self.lib.b2r_rsa_commit_batch_dev
        self._ck(fn(self.h, prog.h, g_lagrange.h, C.c_void_p(n_ptr), C.c_void_p(s_ptr), C.c_void_p(h_ptr), batch, blind_seed,
                    prog.k, ext_k, C.c_void_p(advice_dptr), C.c_void_p(ext_dptr or 0), C.c_void_p(cm_ptr), C.c_void_p(valid_ptr)))

    # -- keygen + full prover (keygen_vk / keygen_pk / create_proof of the reference's bench)
    def rsa_keygen(self, prog: "RsaProgram", g: "Bases", g_lagrange: "Bases") -> "ProvingKey":
        h = C.c_void_p()
        self._ck(self.lib.b2r_rsa_keygen(self.h, prog.h, g.h, g_lagrange.h, C.byref(h)))
        return ProvingKey(self, h, prog)

    # -- device-resident commitment block of the last prove call (the multi-GPU all-gather payload)
    def last_commitments_info(self):
        b, per = C.c_size_t(), C.c_uint32()
        self._ck(self.lib.b2r_last_commitments(self.h, None, 0, 0, C.byref(b), C.byref(per)))
        return b.value, per.value

    def last_commitments(self) -> np.ndarray:
        """-> uint64[batch, 31, 8] (host copy)"""
        b, per = self.last_commitments_info()
        out = np.zeros((b, per, 8), dtype=np.uint64)
        self._ck(self.lib.b2r_last_commitments(self.h, _host_ptr(out), b * per, 0, None, None))
        return out

    def last_commitments_dev(self, dst_ptr: int, capacity_points: int):
        """device-to-device copy into a caller-owned buffer (e.g. a torch tensor's data_ptr) on the context's stream"""
        self._ck(self.lib.b2r_last_commitments(self.h, C.c_void_p(dst_ptr), capacity_points, 1, None, None))

    # -- SHA-256 front end of RSASignatureVerifier (reference src/lib.rs:204-211), value level
    @staticmethod
    def _pack_msgs(msgs):
        offs = np.zeros(len(msgs) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(m) for m in msgs], dtype=np.uint64)
        buf = np.frombuffer(b"".join(bytes(m) for m in msgs) or b"\0", dtype=np.uint8).copy()
        return buf, offs

    def sha256_batch(self, msgs):
        """-> (hash_limbs uint64[batch, 4], digests uint8[batch, 32]) computed on the device"""
        buf, offs = self._pack_msgs(msgs)
        limbs = np.zeros((len(msgs), 4), dtype=np.uint64)
        dig = np.zeros((len(msgs), 32), dtype=np.uint8)
        self._ck(self.lib.b2r_sha256_batch(self.h, _host_ptr(buf), _host_ptr(offs), len(msgs), _host_ptr(limbs), _host_ptr(dig)))
        return limbs, dig

    # -- RSA witness (Circuit::synthesize of the pkcs1v15 circuit)
    def rsa_program(self, bits_len: int, k: int, e: int = 65537) -> "RsaProgram":
        e_le = np.frombuffer(e.to_bytes((e.bit_length() + 7) // 8, "little"), dtype=np.uint8).copy()
        h = C.c_void_p()
        self._ck(self.lib.b2r_rsa_program_build(self.h, bits_len, _host_ptr(e_le), e_le.size, k, C.byref(h)))
        return RsaProgram(self, h, bits_len, k)

    def rsa_program_var(self, bits_len: int, k: int, exp_limb_bits: int) -> "RsaProgram":
        """pkcs1v15 circuit with RSAPubE::Var: third input array = hash limbs then the exponent word"""
        h = C.c_void_p()
        self._ck(self.lib.b2r_rsa_program_build_var(self.h, bits_len, exp_limb_bits, k, C.byref(h)))
        return RsaProgram(self, h, bits_len, k)

    def rsa_program_sha_tail(self, bits_len: int, k: int, e: int = 65537) -> "RsaProgram":
        """RSASignatureVerifier's digest-byte composition + verification (reference src/lib.rs:183-248)"""
        e_le = np.frombuffer(e.to_bytes((e.bit_length() + 7) // 8, "little"), dtype=np.uint8).copy()
        h = C.c_void_p()
        self._ck(self.lib.b2r_rsa_program_build_sha_tail(self.h, bits_len, _host_ptr(e_le), e_le.size, k, C.byref(h)))
        return RsaProgram(self, h, bits_len, k)

    BIGINT_OPS = {"refresh": 6, "add_mod": 7, "sub_mod": 8, "pow_mod": 9, "is_zero": 10, "is_equal_fresh": 11, "is_less_than": 12,
                  "is_less_than_or_equal": 13, "is_greater_than": 14, "is_greater_than_or_equal": 15, "is_in_field": 16,
                  "square": 17, "square_mod": 18}

    def bigint_program(self, op: str, bits_len: int, k: int, exp_limb_bits: int = 5) -> "RsaProgram":
        """one BigIntInstructions method as the reference's unit-test circuits drive it; witness inputs (a, b, n | e)"""
        h = C.c_void_p()
        self._ck(self.lib.b2r_bigint_program_build(self.h, self.BIGINT_OPS[op], bits_len, exp_limb_bits, k, C.byref(h)))
        return RsaProgram(self, h, bits_len, k)


class Bases:
    def __init__(self, ctx: Context, h, n: int):
        self.ctx, self.h, self.n = ctx, h, n

    def download(self, n: int | None = None) -> np.ndarray:
        n = self.n if n is None else n
        out = np.zeros((n, 8), dtype=np.uint64)
        self.ctx._ck(self.ctx.lib.b2r_bases_download(self.ctx.h, self.h, _host_ptr(out), n))
        return out

    def free(self):
        if self.h:
            self.ctx._ck(self.ctx.lib.b2r_bases_free(self.ctx.h, self.h))
            self.h = None


class ProvingKey:
    """keygen_pk output, resident on the GPU (mirrors b2r_pk)."""

    def __init__(self, ctx: Context, h, prog: "RsaProgram"):
        self.ctx, self.h, self.prog = ctx, h, prog
        k, ek, nf, ns, pb = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint64()
        ctx._ck(ctx.lib.b2r_pk_info(h, C.byref(k), C.byref(ek), C.byref(nf), C.byref(ns), C.byref(pb)))
        self.k, self.ext_k, self.num_fixed, self.num_sigma, self.proof_bytes = k.value, ek.value, nf.value, ns.value, pb.value

    def export_vk(self):
        """-> (fixed commitments uint64[num_fixed, 8], sigma commitments uint64[num_sigma, 8], transcript_repr uint64[4])"""
        f = np.zeros((self.num_fixed, 8), dtype=np.uint64)
        s = np.zeros((self.num_sigma, 8), dtype=np.uint64)
        t = np.zeros(4, dtype=np.uint64)
        self.ctx._ck(self.ctx.lib.b2r_pk_export_vk(self.h, _host_ptr(f), _host_ptr(s), _host_ptr(t)))
        return f, s, t

    def set_transcript_repr(self, repr_limbs):
        """install the Rust verifying key's transcript_repr (uint64[4], reduced Montgomery limbs)"""
        t = np.ascontiguousarray(repr_limbs, dtype=np.uint64).reshape(4)
        self.ctx._ck(self.ctx.lib.b2r_pk_set_transcript_repr(self.h, _host_ptr(t)))

    def prove_batch(self, n_limbs, sig_limbs, hash_limbs, seed, nonce: int | None = None):
        """create_proof for a batch -> (proofs uint8[batch, proof_bytes], status uint8[batch]).
        seed: int (64-bit test seed) or 32 bytes (the ChaCha20 key).  nonce=None with an int seed calls the plain
        entry point (the context supplies a fresh nonce per call); an explicit nonce makes the call reproducible."""
        n_limbs = np.ascontiguousarray(n_limbs, dtype=np.uint64)
        sig_limbs = np.ascontiguousarray(sig_limbs, dtype=np.uint64)
        hash_limbs = np.ascontiguousarray(hash_limbs, dtype=np.uint64)
        batch = n_limbs.shape[0]
        proofs = np.zeros((batch, self.proof_bytes), dtype=np.uint8)
        status = np.zeros(batch, dtype=np.uint8)
        self.prove_batch_raw(n_limbs.ctypes.data, sig_limbs.ctypes.data, hash_limbs.ctypes.data, batch, seed, proofs.ctypes.data,
                             status.ctypes.data, nonce=nonce)
        return proofs, status

    PROVE_INPUTS_ON_DEVICE, PROVE_SEED64 = 1, 2

    def prove_msgs_batch(self, n_limbs, sig_limbs, msgs, seed, nonce: int = 0):
        """RSASignatureVerifier::verify_pkcs1v15_signature from the message bytes on: SHA-256 on the device, then
        create_proof -> (proofs, status, digests uint8[batch, 32]).  seed: int (test seed) or the 32-byte ChaCha20 key"""
        n_limbs = np.ascontiguousarray(n_limbs, dtype=np.uint64)
        sig_limbs = np.ascontiguousarray(sig_limbs, dtype=np.uint64)
        batch = n_limbs.shape[0]
        assert len(msgs) == batch
        buf, offs = Context._pack_msgs(msgs)
        flags = 0
        if isinstance(seed, (bytes, bytearray)):
            if len(seed) != 32:
                raise ValueError("seed bytes must be the 32-byte ChaCha20 key")
            key = bytes(seed)
        else:
            key = int(seed).to_bytes(8, "little") + bytes(24)
            flags |= self.PROVE_SEED64
        kb = (C.c_uint8 * 32).from_buffer_copy(key)
        proofs = np.zeros((batch, self.proof_bytes), dtype=np.uint8)
        status = np.zeros(batch, dtype=np.uint8)
        dig = np.zeros((batch, 32), dtype=np.uint8)
        self.ctx._ck(self.ctx.lib.b2r_rsa_prove_msgs_batch(self.ctx.h, self.h, _host_ptr(n_limbs), _host_ptr(sig_limbs), _host_ptr(buf), _host_ptr(offs),
                                                          batch, C.cast(kb, C.c_void_p), nonce, flags, _host_ptr(proofs), _host_ptr(status), _host_ptr(dig)))
        return proofs, status, dig

    def prove_batch_raw(self, n_ptr: int, s_ptr: int, h_ptr: int, batch: int, seed, proofs_ptr: int, status_ptr: int,
                        inputs_on_device: bool = False, nonce: int | None = None):
        if nonce is not None or isinstance(seed, (bytes, bytearray)):
            flags = self.PROVE_INPUTS_ON_DEVICE if inputs_on_device else 0
            if isinstance(seed, (bytes, bytearray)):
                if len(seed) != 32:
                    raise ValueError("seed bytes must be the 32-byte ChaCha20 key")
                key = bytes(seed)
            else:
                key = int(seed).to_bytes(8, "little") + bytes(24)
                flags |= self.PROVE_SEED64
            kb = (C.c_uint8 * 32).from_buffer_copy(key)
            self.ctx._ck(self.ctx.lib.b2r_rsa_prove_batch_ex(self.ctx.h, self.h, C.c_void_p(n_ptr), C.c_void_p(s_ptr), C.c_void_p(h_ptr), batch,
                                                            C.cast(kb, C.c_void_p), nonce or 0, flags, C.c_void_p(proofs_ptr), C.c_void_p(status_ptr)))
            return
        fn = self.ctx.lib.b2r_rsa_prove_batch_dev if inputs_on_device else self.ctx.lib.b2r_rsa_prove_batch
        self.ctx._ck(fn(self.ctx.h, self.h, C.c_void_p(n_ptr), C.c_void_p(s_ptr), C.c_void_p(h_ptr), batch,
               seed, C.c_void_p(proofs_ptr), C.c_void_p(status_ptr)))

    def free(self):
        if self.h:
            self.ctx._ck(self.ctx.lib.b2r_pk_free(self.ctx.h, self.h))
            self.h = None


class RsaProgram:
    """Recorded layout of the reference's pkcs1v15 circuit (benches/bench.rs:132-225)."""

    def __init__(self, ctx: Context, h, bits_len: int, k: int):
        self.ctx, self.h, self.bits_len, self.k = ctx, h, bits_len, k
        self.num_limbs = bits_len // 64
        self.aux_words = int(ctx.lib.b2r_prog_aux_words(h))   # words per instance of the third input array

    def info(self):
        r, v, l = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.ctx._ck(self.ctx.lib.b2r_prog_info(self.h, C.byref(r), C.byref(v), C.byref(l)))
        return {"rows_used": r.value, "num_values": v.value, "num_levels": l.value}

    def witness_batch(self, n_limbs, sig_limbs, hash_limbs, blind_seed: int = 0):
        """-> (advice uint64[batch, 5, 2^k, 4], is_valid uint8[batch])"""
        n_limbs = np.ascontiguousarray(n_limbs, dtype=np.uint64)
        sig_limbs = np.ascontiguousarray(sig_limbs, dtype=np.uint64)
        hash_limbs = np.ascontiguousarray(hash_limbs, dtype=np.uint64)
        batch = n_limbs.shape[0]
        assert n_limbs.shape == (batch, self.num_limbs) and sig_limbs.shape == (batch, self.num_limbs)
        assert hash_limbs.shape == (batch, self.aux_words)
        advice = np.empty((batch, 5, 1 << self.k, 4), dtype=np.uint64)
        valid = np.zeros(batch, dtype=np.uint8)
        self.ctx._ck(self.ctx.lib.b2r_rsa_witness_batch(
            self.ctx.h, self.h, _host_ptr(n_limbs), _host_ptr(sig_limbs), _host_ptr(hash_limbs), batch,
            blind_seed, _host_ptr(advice), _host_ptr(valid)))
        return advice, valid

    def witness_batch_dev(self, n_dptr: int, sig_dptr: int, hash_dptr: int, batch: int, blind_seed: int,
                          advice_dptr: int, valid_dptr: int):
        self.ctx._ck(self.ctx.lib.b2r_rsa_witness_batch_dev(
            self.ctx.h, self.h, C.c_void_p(n_dptr), C.c_void_p(sig_dptr), C.c_void_p(hash_dptr), batch,
            blind_seed, C.c_void_p(advice_dptr), C.c_void_p(valid_dptr)))

    def free(self):
        if self.h:
            self.ctx._ck(self.ctx.lib.b2r_prog_free(self.ctx.h, self.h))
            self.h = None
