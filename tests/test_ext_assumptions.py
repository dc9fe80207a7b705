"""CPU: every third-party fact this repo restates from memory (oracle/EXT_ASSUMPTIONS.md) is held by ONE named constant
/ table on each side; this module reads them out of the oracle (Python constants, C enums) and the product (CUDA / C++
sources) and checks that the two sides and the document agree.  It cannot tell whether the recalled facts are RIGHT
(that needs the Rust crates: shim/tests/vk_matches.rs) - it makes sure a correction is a one-place edit per side and
that nobody changes one side only."""
import os
import re

import bn254 as O
import plonk as PL

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _src(*parts):
    return open(os.path.join(ROOT, *parts)).read()


PROVER = _src("halo2-rsa_b200", "csrc", "prover.cu")
CIRCUIT = _src("halo2-rsa_b200", "csrc", "circuit.hpp")
TRANSCRIPT = _src("halo2-rsa_b200", "csrc", "transcript.hpp")
NTT = _src("halo2-rsa_b200", "csrc", "ntt.cu")
CPROVER = _src("oracle", "plonk_prover.c")
DOC = _src("oracle", "EXT_ASSUMPTIONS.md")


def _enum(text, first, last):
    """names of a C enum run `first, ..., last` in declaration order"""
    m = re.search(re.escape(first) + r"\s*=\s*0(.*?)" + re.escape(last), text, re.S)
    assert m, first
    return [first] + [n for n in re.findall(r"[A-Za-z_0-9]+", m.group(1))] + [last]


def _int_consts(text, names):
    out = {}
    for n in names:
        m = re.search(r"\b" + n + r"\s*=\s*(\d+)", text)
        assert m, n
        out[n] = int(m.group(1))
    return out


def test_document_lists_every_assumption_with_both_sides():
    rows = [l for l in DOC.splitlines() if re.match(r"\| [A-D]\d+ \|", l)]
    ids = [l.split("|")[1].strip() for l in rows]
    assert ids == [f"A{i}" for i in range(1, 7)] + [f"B{i}" for i in range(1, 7)] + [f"C{i}" for i in range(1, 10)] + [f"D{i}" for i in range(1, 10)]
    for l in rows:                                   # id | assumption | upstream item | oracle | product | confidence
        cells = [c.strip() for c in l.split("|")[1:-1]]
        assert len(cells) == 6 and all(cells[:5]), l
    # every source file the table points at exists
    for path in set(re.findall(r"`((?:oracle|csrc|include|shim)/[A-Za-z0-9_./]+)`", DOC)) | set(re.findall(r"`([a-z_0-9]+\.(?:py|c|h|cu|cuh|hpp))`", DOC)):
        cands = [os.path.join(ROOT, path), os.path.join(ROOT, "oracle", path), os.path.join(ROOT, "halo2-rsa_b200", "csrc", path),
                 os.path.join(ROOT, "halo2-rsa_b200", path)]
        assert any(os.path.exists(c) for c in cands), path


def test_C1_C5_fixed_column_order():
    want = ["SA", "SB", "SC", "SD", "SE", "MUL_AB", "MUL_CD", "SE_NEXT", "CONST", "TAG_COMP", "TAG_OVER", "T_TAG", "T_VALUE", "S_COMP", "S_OVER"]
    assert [n[3:] for n in _enum(PROVER, "FX_SA", "FX_S_OVER")] == want
    assert [n[2:] for n in _enum(CPROVER, "F_SA", "F_S_OVER")] == want
    py = [PL.F_SA, PL.F_SB, PL.F_SC, PL.F_SD, PL.F_SE, PL.F_MUL_AB, PL.F_MUL_CD, PL.F_SE_NEXT, PL.F_CONST, PL.F_TAG_COMP, PL.F_TAG_OVER,
          PL.F_T_TAG, PL.F_T_VALUE, PL.F_S_COMP, PL.F_S_OVER]
    assert py == list(range(15)) == list(range(PL.NUM_FIXED))
    # the recorder's nine MainGate columns, same order
    m = re.search(r"enum FixedCol\s*\{\s*F_SA\s*=\s*0(.*?)\}", CIRCUIT, re.S)
    assert m and [n for n in re.findall(r"F_[A-Z_]+", "F_SA" + m.group(1))][:9] == ["F_" + n for n in want[:9]]


def test_C2_C6_C7_shape_constants():
    c = _int_consts(PROVER, ["NADV", "NFIXED", "NPERM", "NLOOK", "NSETS", "CHUNK", "BF", "QD"])
    o = _int_consts(CPROVER, ["NADV", "NFIX", "NPERM", "NLOOK", "CHUNK", "NSETS", "BF", "QD"])
    assert (c["NADV"], c["NFIXED"], c["NPERM"], c["NLOOK"], c["NSETS"], c["CHUNK"], c["BF"], c["QD"]) == \
           (o["NADV"], o["NFIX"], o["NPERM"], o["NLOOK"], o["NSETS"], o["CHUNK"], o["BF"], o["QD"]) == \
           (PL.NUM_ADVICE, PL.NUM_FIXED, len(PL.PERM_COLUMNS), len(PL.LOOKUPS), (len(PL.PERM_COLUMNS) + PL.CHUNK - 1) // PL.CHUNK, PL.CHUNK, PL.BF, 4)
    assert PL.PERM_COLUMNS == [("a", i) for i in range(5)] + [("i", 0)]                       # C2
    assert int(re.search(r"BLINDING_ROWS\s*=\s*(\d+)", CIRCUIT).group(1)) == PL.BF + 1        # C7: usable rows = n - 6
    # C6: lookups = (advice column, tag column, selector column)
    assert PL.LOOKUPS == [(0, PL.F_TAG_COMP, PL.F_S_COMP), (1, PL.F_TAG_COMP, PL.F_S_COMP), (2, PL.F_TAG_COMP, PL.F_S_COMP),
                          (3, PL.F_TAG_COMP, PL.F_S_COMP), (0, PL.F_TAG_OVER, PL.F_S_OVER)]
    assert re.search(r"lookup_acol\(int l\) \{ return l < 4 \? l : 0; \}", PROVER)
    assert re.search(r"lookup_ftag\(int l\) \{ return l < 4 \? FX_TAG_COMP : FX_TAG_OVER; \}", PROVER)
    assert re.search(r"lookup_fsel\(int l\) \{ return l < 4 \? FX_S_COMP : FX_S_OVER; \}", PROVER)
    assert re.search(r"LK_ACOL\[NLOOK\] = \{0, 1, 2, 3, 0\}", CPROVER)


def test_D3_D4_proof_shape():
    pb = re.search(r"\*proof_bytes = 32 \* \((.*?)\);", PROVER).group(1)
    c = _int_consts(PROVER, ["NADV", "NLOOK", "NSETS", "QD", "NPOINTS", "NEVAL"])
    assert 32 * eval(pb, {}, c) == PL.proof_length() == 2848
    assert c["NEVAL"] == len(PL.ADVICE_QUERIES) + len(PL.FIXED_QUERIES) + 1 + len(PL.PERM_COLUMNS) + (3 * c["NSETS"] - 1) + 5 * c["NLOOK"] == 58
    assert PL.ADVICE_QUERIES == [(0, 0), (1, 0), (2, 0), (3, 0), (4, 0), (4, 1)]               # D4: e is also queried at the next row


def test_A3_D1_transcript_constants():
    assert '"Halo2-Transcript"' in TRANSCRIPT and 'b"Halo2-Transcript"' in _src("oracle", "plonk.py") and '"Halo2-Transcript"' in CPROVER
    # prefix bytes 0 squeeze / 1 point / 2 scalar on the three sides
    assert re.search(r"b\[0\] = 2;", TRANSCRIPT) and re.search(r"b\[0\] = 1;", TRANSCRIPT) and re.search(r"uint8_t p = 0;", TRANSCRIPT)
    assert re.search(r"b\[0\] = 2;", CPROVER) and re.search(r"b\[0\] = 1;", CPROVER)
    # A3: sign of y in bit 7 of the last byte
    assert "c[31] |= (uint8_t)((y.l[0] & 1u) << 7)" in TRANSCRIPT
    assert "c[31] |= (uint8_t)((y[0] & 1) << 7)" in CPROVER
    assert PL.compress_point((1, 3))[31] == 0x80 and PL.compress_point((1, 2))[31] == 0 and PL.compress_point(None) == bytes(32)


def test_A6_field_constants():
    # the literals in the product's sources are the numbers the oracle derives / checks mathematically
    def words_le(hexwords):
        return sum(int(w, 16) << (32 * i) for i, w in enumerate(hexwords))
    root = re.search(r"fe_t root_of_unity\(\).*?w\[8\] = \{(.*?)\};", NTT, re.S).group(1)
    assert words_le(re.findall(r"0x([0-9a-f]+)u", root)) == O.ROOT_OF_UNITY
    zeta = re.search(r"fe_t fr_zeta\(\).*?w\[8\] = \{(.*?)\};", NTT, re.S).group(1)
    assert words_le(re.findall(r"0x([0-9a-f]+)u", zeta)) == O.ZETA
    delta = re.search(r"fe_t fr_delta\(\).*?w\[8\] = \{(.*?)\};", PROVER, re.S).group(1)
    assert words_le(re.findall(r"0x([0-9a-f]+)u", delta)) == O.DELTA
    r = O.R_MOD
    assert pow(O.ROOT_OF_UNITY, 1 << 28, r) == 1 and pow(O.ROOT_OF_UNITY, 1 << 27, r) != 1
    assert pow(O.ZETA, 3, r) == 1 and O.ZETA != 1
    assert O.DELTA == pow(7, 1 << 28, r)


def test_D2_stand_in_is_overridable():
    """the vk transcript representative cannot be reproduced here: the ABI must let the Rust side install the real one"""
    hdr = _src("include", "b2rsa.h")
    assert "b2r_pk_set_transcript_repr" in hdr and "b2r_pk_set_transcript_repr" in PROVER
    assert "b2r_pk_set_transcript_repr" in _src("shim", "src", "lib.rs") and "hash_into" in _src("shim", "src", "lib.rs")
