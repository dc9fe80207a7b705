"""CPU: the drop-in boundary.  libb2rsa.so must load without a GPU, export every symbol that
include/b2rsa.h declares, and fail LOUDLY (B2R_ERR_NO_DEVICE) instead of falling back to a CPU
path when no sm_100 device is present.  No compute entry point is called here."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b2rsa.h")
LIB = os.path.join(ROOT, "halo2-rsa_b200", "lib", "libb2rsa.so")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2r_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "halo2-rsa_b200"), "-j8", "-s"])
    return C.CDLL(LIB)


def test_header_declares_the_hot_path_entry_points():
    syms = declared_symbols()
    for need in ("b2r_ntt_fr", "b2r_intt_fr", "b2r_coset_ntt_fr", "b2r_coset_intt_fr", "b2r_msm_g1", "b2r_msm_g1_batch",
                 "b2r_bases_register", "b2r_srs_setup", "b2r_rsa_program_build", "b2r_rsa_witness_batch",
                 "b2r_rsa_commit_batch", "b2r_last_error", "b2r_ctx_create"):
        assert need in syms
    assert len(syms) >= 35


def test_library_exports_every_declared_symbol(lib):
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_header_is_plain_c():
    """the header must compile as C (extern "C" boundary: plain pointers and sizes, no C++/torch types)"""
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", HEADER])


def test_python_binding_signatures_cover_the_header(lib):
    import b2rsa
    l = b2rsa.load_library()
    for s in declared_symbols():
        assert getattr(l, s).restype is not None or s == "b2r_version"


def test_version_and_no_cpu_fallback(lib):
    import torch
    lib.b2r_version.restype = C.c_char_p
    assert lib.b2r_version().decode().startswith("b2rsa")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the loud-failure path is exercised on CPU boxes")
    h = C.c_void_p()
    rc = lib.b2r_ctx_create(0, C.byref(h))
    assert rc == -3 and not h.value          # B2R_ERR_NO_DEVICE: there is no CPU implementation behind the ABI
    import b2rsa
    with pytest.raises(b2rsa.B2RError):
        b2rsa.Context(0)
