"""CPU: host logic of the product - the circuit recorder (csrc/circuit.hpp), i.e. the host-side
mirror of RSAChip / BigIntChip / maingate that produces the GPU witness program - compiled by g++
and compared with the oracle's independently written row-by-row synthesis: same rows, same fixed
columns, same copy constraints, same range-lookup tags, for every BASELINE key size."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import cpu_oracle as CO
import rsa_fixtures as RF

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("host") / "circuit_host_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", out, os.path.join(HERE, "host", "circuit_host_test.cpp")])
    return out


def record(exe, bits, k, e=65537):
    p = subprocess.run([exe, str(bits), str(k), str(e)], capture_output=True, text=True)
    return p.returncode, dict(re.findall(r"(\w+)=(\S+)", p.stdout)), p.stdout


def oracle_digest(bits, k, e=65537):
    t = CO.RsaTable(bits, k)
    n, s, h = RF.instance(bits, 0)
    nl = bits // 64
    assert t.synthesize(RF.limbs64(n, nl), RF.limbs64(s, nl), RF.limbs64(h, 4), e) >= 0
    out = np.zeros(5, dtype=np.uint64)
    CO.lib().orc_table_layout_digest(t.h, C.c_void_p(out.ctypes.data))
    t.free()
    return [int(x) for x in out]


@pytest.mark.parametrize("bits,k", [(1024, 15), (2048, 17), (4096, 18)])
def test_recorded_layout_equals_oracle_layout(exe, bits, k):
    rc, d, raw = record(exe, bits, k)
    assert rc == 0, raw
    rows, hf, ncop, hc, hr = oracle_digest(bits, k)
    assert int(d["rows"]) == rows
    assert int(d["fixed"]) == hf
    assert int(d["ncopies"]) == ncop and int(d["copies"]) == hc
    assert int(d["range"]) == hr
    nl = bits // 64
    assert int(d["bigops"]) == 19 + 2          # 17 squarings + 2 multiplications, 2 sub_unchecked (assert_in_field)
    assert int(d["nodes"]) > 19 * 2 * nl * nl  # at least the mul_add chains


def test_layout_is_data_independent():
    """the call sequence does not depend on the witness -> same digest for other inputs"""
    t = CO.RsaTable(2048, 17)
    digs = []
    for i in (0, 5):
        t = CO.RsaTable(2048, 17)
        n, s, h = RF.instance(2048, i)
        t.synthesize(RF.limbs64(n, 32), RF.limbs64(s, 32), RF.limbs64(h, 4))
        out = np.zeros(5, dtype=np.uint64)
        CO.lib().orc_table_layout_digest(t.h, C.c_void_p(out.ctypes.data))
        digs.append(out.tolist())
        t.free()
    assert digs[0] == digs[1]


def test_other_exponents_and_errors(exe):
    rc, d, _ = record(exe, 1024, 15, e=3)
    assert rc == 0 and int(d["bigops"]) == 2 + 2 + 2      # 2 squarings + 2 multiplications + in-field subs
    assert int(d["rows"]) == oracle_digest(1024, 15, e=3)[0]
    rc, d, raw = record(exe, 2048, 16)                   # does not fit 2^16 rows -> B2R_ERR_LAYOUT
    assert rc == 2 and "error=-5" in raw
    rc, d, raw = record(exe, 2000, 17)                   # bits_len % limb_width != 0 (chip.rs:1175 assert)
    assert rc == 2 and "error=-1" in raw


def _digest_of(table):
    out = np.zeros(5, dtype=np.uint64)
    CO.lib().orc_table_layout_digest(table.h, C.c_void_p(out.ctypes.data))
    return [int(x) for x in out]


def _assert_same_layout(d, dig):
    rows, hf, ncop, hc, hr = dig
    assert int(d["rows"]) == rows
    assert int(d["fixed"]) == hf
    assert int(d["ncopies"]) == ncop and int(d["copies"]) == hc
    assert int(d["range"]) == hr


@pytest.mark.parametrize("op,opid", [("refresh", 6), ("add_mod", 7), ("sub_mod", 8), ("pow_mod", 9)] + [(o, CO.BIGINT_OPS[o]) for o in (
    "is_zero", "is_equal_fresh", "is_less_than", "is_less_than_or_equal", "is_greater_than", "is_greater_than_or_equal", "is_in_field",
    "square", "square_mod")])
def test_bigint_ops_layout_equals_oracle_layout(exe, op, opid):
    """the BigIntInstructions methods the pkcs1v15 circuit does not call (refresh, add_mod, sub_mod, variable-exponent
    pow_mod: src/big_integer/chip.rs:168-233, 452-529, 664-696) recorded by the host mirror vs the oracle's row-by-row
    restatement of the reference's unit-test circuits"""
    bits, k, ebits = 512, 15, 5
    p = subprocess.run([exe, str(bits), str(k), "op", str(opid), str(ebits)], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout
    d = dict(re.findall(r"(\w+)=(\S+)", p.stdout))
    import random
    r = random.Random(opid)
    n = r.getrandbits(bits) | (1 << (bits - 1)) | 1
    a, b = r.getrandbits(bits) % n, r.getrandbits(bits) % n
    nl = bits // 64
    t = CO.RsaTable(bits, k)
    L = CO.lib()
    L.orc_bigint_op.restype = C.c_int
    out = np.zeros((2 * nl + 2, 4), dtype=np.uint64)
    aw = CO.int_to_limbs64(a, nl)
    bw = CO.int_to_limbs64(17 if op == "pow_mod" else b, nl)
    nw = CO.int_to_limbs64(n, nl)
    rc = L.orc_bigint_op(t.h, C.c_int(opid), C.c_int(bits), CO._p(aw), CO._p(bw), CO._p(nw), C.c_int(ebits if op == "pow_mod" else nl), CO._p(out))
    assert rc > 0 and t.check()[0] == 0
    _assert_same_layout(d, _digest_of(t))
    t.free()


def test_rsa_var_layout_equals_oracle_layout(exe):
    """pkcs1v15 circuit with RSAPubE::Var (src/chip.rs:58-70, 99-114): exponent assigned as a witness, 17 bits walked"""
    bits, k = 1024, 18
    p = subprocess.run([exe, str(bits), str(k), "var", "17"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout
    d = dict(re.findall(r"(\w+)=(\S+)", p.stdout))
    n, s, h = RF.instance(bits, 0)
    nl = bits // 64
    t = CO.RsaTable(bits, k)
    assert t.synthesize_var(RF.limbs64(n, nl), RF.limbs64(s, nl), RF.limbs64(h, 4), 65537, 17) == 1
    _assert_same_layout(d, _digest_of(t))
    t.free()


def test_refresh_aux_kat(exe):
    """the reference's own KAT: RefreshAux::new(32, 1, 1).increased_limbs_vec == [1, 0] (src/big_integer/mod.rs:504-509)"""
    p = subprocess.run([exe, "aux", "32", "1", "1"], capture_output=True, text=True)
    assert p.stdout.strip() == "aux=1,0"
    p = subprocess.run([exe, "aux", "64", "32", "32"], capture_output=True, text=True)
    inc = [int(x) for x in p.stdout.strip()[4:].split(",")]
    assert len(inc) == 64 and inc[0] == 1 and max(inc) == 2   # 2048-bit operands: every product limb spills 1-2 limbs


@pytest.mark.parametrize("bits,k", [(1024, 15), (2048, 17)])
def test_sha_tail_layout_equals_oracle_layout(exe, bits, k):
    """RSASignatureVerifier::verify_pkcs1v15_signature from the digest bytes on (src/lib.rs:183-248: byte cells ->
    assign_constant / mul_add composition into four limbs -> verify in the same region), recorder vs oracle; the oracle
    table satisfies every gate / lookup / copy constraint and reports the same verdict as the bench circuit"""
    p = subprocess.run([exe, str(bits), str(k), "digest", "65537"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout
    d = dict(re.findall(r"(\w+)=(\S+)", p.stdout))
    nl = bits // 64
    n, s, h = RF.instance(bits, 2)
    t = CO.RsaTable(bits, k)
    assert t.synthesize_digest(RF.limbs64(n, nl), RF.limbs64(s, nl), RF.limbs64(h, 4)) == 1
    assert t.check()[0] == 0
    _assert_same_layout(d, _digest_of(t))
    rows_tail = t.rows()
    t.free()
    # against the bench circuit: 4 range-checked limb assignments (2 rows each: 8 sublimbs, 4 per row) and the assert_one row are replaced by
    # 32 byte cells + 4 x (1 + 2 x 8) composition rows
    t2 = CO.RsaTable(bits, k)
    assert t2.synthesize(RF.limbs64(n, nl), RF.limbs64(s, nl), RF.limbs64(h, 4)) == 1
    assert rows_tail - t2.rows() == (32 + 4 * 17) - (4 * 2 + 1)
    t2.free()
    # a wrong digest byte -> is_valid = 0, constraints still satisfied (the verifier returns the bit, it does not assert it)
    hb = RF.limbs64(h, 4).copy()
    hb[1] ^= np.uint64(1 << 40)
    t3 = CO.RsaTable(bits, k)
    assert t3.synthesize_digest(RF.limbs64(n, nl), RF.limbs64(s, nl), hb) == 0
    assert t3.check()[0] == 0
    t3.free()
