"""TEST INFRASTRUCTURE ONLY - oracle keygen / prover / verifier for the RSA circuit.

Restates, in plain Python integers (FFT and MSM through the C oracle), what the reference's bench does
around the two hot paths (benches/bench.rs:228-345): ParamsKZG::setup, keygen_vk / keygen_pk,
create_proof::<KZGCommitmentScheme<Bn256>, ProverGWC, Challenge255, _, Blake2bWrite, _> and
verify_proof with VerifierGWC.  The algorithm lives in the third-party crates halo2_proofs
(privacy-scaling-explorations/halo2, 2022-10 era) and maingate (halo2wrong rev 63bde545), neither
vendored under /root/reference, so this is a restatement of their published protocol:

  constraint system  MainGate (5 advice a..e, 9 fixed, one degree-3 gate, instance column) +
                     RangeChip (4 composition lookups on a..d + 1 overflow lookup on a into the tagged
                     table (t_tag, t_value)); permutation over [a, b, c, d, e, instance]; degree 5,
                     5 blinding factors, extended domain 2^(k+2), 4 quotient pieces
  proof order        advice commitments | theta | per lookup A', S' | beta, gamma | permutation Z's |
                     lookup Z's | random poly | y | h pieces | x | evaluations | v | GWC witnesses
  transcript         Blake2b-512 personalised "Halo2-Transcript", prefix bytes 0/1/2, Challenge255

"parity unpinned": the reference holds no proof bytes (its only assertion at this level is
verify_proof(..).is_ok(), benches/bench.rs:336-343, on OsRng randomness), Rust's Debug-string based
vk.transcript_repr cannot be reproduced without the crate, and the compressed-point flag bit is
recalled.  What IS checked: the product's proofs are byte-identical to this independent
restatement for the same seeded randomness, and this verifier accepts them.  The verifier replaces
the pairing check by the equivalent identity under the known SRS secret s (the SRS is seeded):
(s - z) * W == C - e * G  for every opening.
"""
import ctypes as C
import hashlib
import struct

import numpy as np

import bn254 as O
import cpu_oracle as CO

R = O.R_MOD
Q = O.Q_MOD
MASK64 = (1 << 64) - 1

NUM_ADVICE = 5
BF = 5  # blinding factors: max(3, max queries per advice column = 2) + 2
# fixed column indices
F_SA, F_SB, F_SC, F_SD, F_SE, F_MUL_AB, F_MUL_CD, F_SE_NEXT, F_CONST = range(9)
F_TAG_COMP, F_TAG_OVER, F_T_TAG, F_T_VALUE, F_S_COMP, F_S_OVER = range(9, 15)
NUM_FIXED = 15
ADVICE_QUERIES = [(0, 0), (1, 0), (2, 0), (3, 0), (4, 0), (4, 1)]
FIXED_QUERIES = [(i, 0) for i in range(NUM_FIXED)]
PERM_COLUMNS = [("a", 0), ("a", 1), ("a", 2), ("a", 3), ("a", 4), ("i", 0)]
CHUNK = 3  # cs.degree() - 2
# lookups: (advice column, tag fixed column, selector fixed column)
LOOKUPS = [(0, F_TAG_COMP, F_S_COMP), (1, F_TAG_COMP, F_S_COMP), (2, F_TAG_COMP, F_S_COMP), (3, F_TAG_COMP, F_S_COMP),
           (0, F_TAG_OVER, F_S_OVER)]
# blinding streams (shared convention with the product, csrc/prover.cu)
ST_ADVICE, ST_LOOKUP_A, ST_LOOKUP_S, ST_LOOKUP_Z, ST_PERM_Z, ST_RANDOM_POLY = 0, 8, 16, 24, 32, 40


# ---- conversions between Python ints and the C oracle's arrays --------------------------------
def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def ints_to_np(xs):
    """canonical ints -> uint64[n,4] Montgomery"""
    a = np.frombuffer(b"".join(x.to_bytes(32, "little") for x in xs), dtype=np.uint64).reshape(-1, 4).copy()
    CO.lib().orc_fr_to_mont_array(_ptr(a), C.c_size_t(a.shape[0]))
    return a


def np_to_ints(a):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4).copy()
    CO.lib().orc_fr_from_mont_array(_ptr(a), C.c_size_t(a.shape[0]))
    b = a.tobytes()
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def g1_np_to_point(a):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(2, 4).copy()
    CO.lib().orc_fq_from_mont_array(_ptr(a), C.c_size_t(2))
    b = a.tobytes()
    x, y = int.from_bytes(b[:32], "little"), int.from_bytes(b[32:], "little")
    return None if x == 0 and y == 0 else (x, y)


# ---- keyed randomness (the stand-in for OsRng on both sides) -----------------------------------
# Blinding stream v2 (product: csrc/devutil.cuh): one ChaCha20 block (RFC 7539 block function) per blinded cell,
# key = 32 seed bytes, word 12 = row | stream << 24, word 13 = proof index, words 14-15 = call nonce; the 64 output
# bytes are a little-endian integer reduced mod r (halo2curves Fr::from_bytes_wide, i.e. what Fr::random does).
_SEED64_PAD = b"b2rsa-blind-seed64-v2"


def blind_key(seed):
    """seed: 32 bytes (the ChaCha20 key) or an int (64-bit test seed, expanded with a fixed pad) -> 8 key words"""
    if isinstance(seed, int):
        kb = (seed & MASK64).to_bytes(8, "little") + _SEED64_PAD + bytes(24 - len(_SEED64_PAD))
    else:
        kb = bytes(seed)
    assert len(kb) == 32
    return struct.unpack("<8I", kb)


def _chacha20_blocks(key_words, w12, w13, nonce):
    """vectorised over w12 (uint32 array) -> uint32[len, 16] output blocks"""
    w12 = np.asarray(w12, dtype=np.uint32)
    m = w12.shape[0]
    init = np.empty((16, m), dtype=np.uint32)
    const = (0x61707865, 0x3320646E, 0x79622D32, 0x6B206574)
    for i in range(4):
        init[i] = const[i]
    for i in range(8):
        init[4 + i] = key_words[i]
    init[12] = w12
    init[13] = w13 & 0xFFFFFFFF
    init[14] = nonce & 0xFFFFFFFF
    init[15] = (nonce >> 32) & 0xFFFFFFFF
    x = init.copy()

    def rotl(v, n):
        return (v << np.uint32(n)) | (v >> np.uint32(32 - n))

    def qr(a, b, c, d):
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16)
        x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12)
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8)
        x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7)

    with np.errstate(over="ignore"):
        for _ in range(10):
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
        x += init
    return np.ascontiguousarray(x.T)


def blind_fes(seed, proof, stream, rows, nonce=0):
    """canonical blinding values of (proof, stream) at the given rows"""
    rows = np.asarray(list(rows), dtype=np.uint32)
    if rows.size == 0:
        return []
    blocks = _chacha20_blocks(blind_key(seed), rows | np.uint32(stream << 24), proof, nonce).astype("<u4").tobytes()
    return [int.from_bytes(blocks[64 * i:64 * i + 64], "little") % R for i in range(rows.size)]


def blind_fe(seed, proof, stream, row, nonce=0):
    return blind_fes(seed, proof, stream, [row], nonce)[0]


# ---- domain ---------------------------------------------------------------------------------------
class Domain:
    """EvaluationDomain::new(j = cs.degree() = 5, k)"""

    def __init__(self, k, degree=5):
        self.k, self.n = k, 1 << k
        self.qd = degree - 1
        ext_k = k
        while (1 << ext_k) < self.n * self.qd:
            ext_k += 1
        self.ext_k, self.ext_n = ext_k, 1 << ext_k
        self.step = self.ext_n // self.n
        self.omega = O.omega_for(k)
        self.omega_inv = pow(self.omega, -1, R)
        self.ext_omega = O.omega_for(ext_k)
        zn, wn = pow(O.ZETA, self.n, R), pow(self.ext_omega, self.n, R)
        self.t_inv = [pow((zn * pow(wn, i, R) - 1) % R, -1, R) for i in range(self.step)]

    def lagrange_to_coeff(self, vals):
        return np_to_ints(CO.lagrange_to_coeff(ints_to_np(vals), self.k))

    def coeff_to_lagrange(self, coeffs):
        return np_to_ints(CO.best_fft(ints_to_np(coeffs), ints_to_np([self.omega])[0], self.k))

    def coeff_to_extended(self, coeffs):
        return np_to_ints(CO.coeff_to_extended(ints_to_np(coeffs), self.k, self.ext_k))

    def extended_to_coeff(self, ext):
        return np_to_ints(CO.extended_to_coeff(ints_to_np(ext), self.ext_k))[: self.n * self.qd]

    def rotate_omega(self, x, rot):
        return x * pow(self.omega, rot, R) % R if rot >= 0 else x * pow(self.omega_inv, -rot, R) % R

    def l_i(self, x, rot):
        """Lagrange basis polynomial of row `rot mod n` at x"""
        w = pow(self.omega, rot % self.n, R)
        return (pow(x, self.n, R) - 1) * pow(self.n, -1, R) % R * w % R * pow((x - w) % R, -1, R) % R


def eval_poly(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R
    return acc


# ---- SRS -------------------------------------------------------------------------------------------
class Srs:
    """ParamsKZG::setup(k, rng) for a seeded secret: g[i] = s^i G, g_lagrange[i] = L_i(s) G"""

    def __init__(self, k, secret=None):
        self.k, self.n = k, 1 << k
        self.s = O.srs_secret(k) if secret is None else secret
        pw, acc = [], 1
        for _ in range(self.n):
            pw.append(acc)
            acc = acc * self.s % R
        self.g = CO.g1_scalar_muls(ints_to_np(pw))
        self.g_lagrange = CO.g1_scalar_muls(ints_to_np(O.lagrange_at(self.s, k)))

    def commit(self, coeffs):
        return g1_np_to_point(CO.best_multiexp(ints_to_np(coeffs), self.g))

    def commit_lagrange(self, vals):
        return g1_np_to_point(CO.best_multiexp(ints_to_np(vals), self.g_lagrange))


# ---- transcript (Blake2bWrite / Blake2bRead with Challenge255) ---------------------------------
def compress_point(P):
    if P is None:
        return bytes(32)
    b = bytearray(P[0].to_bytes(32, "little"))
    b[31] |= (P[1] & 1) << 7
    return bytes(b)


def decompress_point(b):
    if b == bytes(32):
        return None
    sign = b[31] >> 7
    x = int.from_bytes(b, "little") & ((1 << 255) - 1)
    y = pow((x * x * x + 3) % Q, (Q + 1) // 4, Q)
    if y * y % Q != (x * x * x + 3) % Q:
        raise ValueError("point not on curve")
    if (y & 1) != sign:
        y = Q - y
    return (x, y)


class Transcript:
    def __init__(self, proof=None):
        self.h = hashlib.blake2b(digest_size=64, person=b"Halo2-Transcript")
        self.out = bytearray()
        self.inp, self.pos = proof, 0

    def common_point(self, P):
        x, y = (0, 0) if P is None else P
        self.h.update(b"\x01" + x.to_bytes(32, "little") + y.to_bytes(32, "little"))

    def common_scalar(self, s):
        self.h.update(b"\x02" + s.to_bytes(32, "little"))

    def write_point(self, P):
        self.common_point(P)
        self.out += compress_point(P)

    def write_scalar(self, s):
        self.common_scalar(s)
        self.out += s.to_bytes(32, "little")

    def read_point(self):
        P = decompress_point(bytes(self.inp[self.pos:self.pos + 32]))
        self.pos += 32
        self.common_point(P)
        return P

    def read_scalar(self):
        s = int.from_bytes(self.inp[self.pos:self.pos + 32], "little")
        self.pos += 32
        if s >= R:
            raise ValueError("non-canonical scalar")
        self.common_scalar(s)
        return s

    def squeeze(self):
        self.h.update(b"\x00")
        return int.from_bytes(self.h.copy().digest(), "little") % R


# ---- keygen ------------------------------------------------------------------------------------------
def circuit_layout(bits_len, k, e=65537):
    """runs the witness oracle once (any valid instance: the layout is data independent) and returns
    what keygen needs: fixed columns, range selectors / tags, copy constraints, tag -> bits"""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
    import rsa_fixtures as RF
    t = CO.RsaTable(bits_len, k)
    nl = bits_len // 64
    nn, ss, hh = RF.instance(bits_len, 0)
    assert t.synthesize(RF.limbs64(nn, nl), RF.limbs64(ss, nl), RF.limbs64(hh, 4), e) == 1
    n = 1 << k
    L = CO.lib()
    fixed = np.zeros((9, n, 4), dtype=np.uint64)
    L.orc_table_fixed(t.h, _ptr(fixed))
    rng = np.zeros((4, n), dtype=np.uint8)
    L.orc_table_range(t.h, _ptr(rng))
    L.orc_table_copies.restype = C.c_size_t
    nc = L.orc_table_copies(t.h, None, C.c_size_t(0))
    copies = np.zeros((nc, 4), dtype=np.uint32)
    L.orc_table_copies(t.h, _ptr(copies), C.c_size_t(nc))
    tb = np.zeros(16, dtype=np.int32)
    L.orc_table_tag_bits(t.h, _ptr(tb))
    rows = t.rows()
    t.free()
    return {"k": k, "fixed9": [np_to_ints(fixed[i]) for i in range(9)], "range": rng, "copies": copies.tolist(),
            "tag_bits": {i: int(b) for i, b in enumerate(tb) if b}, "rows": rows}


def build_permutation(n, copies):
    """halo2 permutation::keygen::Assembly::copy (cycle merge, smaller into larger, splice)"""
    ncol = len(PERM_COLUMNS)
    mapping = [[(c, r) for r in range(n)] for c in range(ncol)]
    aux = [[(c, r) for r in range(n)] for c in range(ncol)]
    sizes = [[1] * n for _ in range(ncol)]
    for c1, r1, c2, r2 in copies:
        lc, rc = aux[c1][r1], aux[c2][r2]
        if lc == rc:
            continue
        if sizes[lc[0]][lc[1]] < sizes[rc[0]][rc[1]]:
            lc, rc = rc, lc
        sizes[lc[0]][lc[1]] += sizes[rc[0]][rc[1]]
        i = rc
        while True:
            aux[i[0]][i[1]] = lc
            i = mapping[i[0]][i[1]]
            if i == rc:
                break
        mapping[c1][r1], mapping[c2][r2] = mapping[c2][r2], mapping[c1][r1]
    return mapping


def keygen_arrays(layout, srs, threads=0):
    """keygen_vk + keygen_pk with every polynomial kept as a uint64[., 4] Montgomery array (the form the C prover
    oracle/plonk_prover.c consumes; Python integer lists of the 2^(k+2) cosets would not fit comfortably at k = 18)"""
    k = layout["k"]
    dom = Domain(k)
    n, u = dom.n, dom.n - (BF + 1)
    rng = layout["range"]
    fixed = [list(col) for col in layout["fixed9"]]
    fixed.append([int(x) for x in rng[1]])   # tag_composition
    fixed.append([int(x) for x in rng[3]])   # tag_overflow
    t_tag, t_val = [0], [0]
    for tag in sorted(layout["tag_bits"]):
        bits = layout["tag_bits"][tag]
        t_tag += [tag] * (1 << bits)
        t_val += list(range(1 << bits))
    assert len(t_tag) <= u, "lookup table does not fit"
    fixed.append(t_tag + [0] * (n - len(t_tag)))
    fixed.append(t_val + [0] * (n - len(t_val)))
    fixed.append([int(x) for x in rng[0]])   # s_composition (complex selector -> fixed column)
    fixed.append([int(x) for x in rng[2]])   # s_overflow
    assert len(fixed) == NUM_FIXED
    ka = {"k": k, "dom": dom, "table_len": len(t_tag)}
    ka["fixed_values"] = np.stack([ints_to_np(v) for v in fixed])
    mapping = build_permutation(n, layout["copies"])
    omega_pows, acc = [], 1
    for _ in range(n):
        omega_pows.append(acc)
        acc = acc * dom.omega % R
    delta_pows = [pow(O.DELTA, j, R) for j in range(len(PERM_COLUMNS))]
    ka["sigma_values"] = np.stack([ints_to_np([delta_pows[mc] * omega_pows[mr] % R for mc, mr in mapping[c]]) for c in range(len(PERM_COLUMNS))])
    for name in ("fixed", "sigma"):
        vals = ka[name + "_values"]
        polys = np.stack([CO.lagrange_to_coeff(v, k, threads) for v in vals])
        ka[name + "_polys"] = polys
        ka[name + "_cosets"] = np.stack([CO.coeff_to_extended(p, k, dom.ext_k, threads) for p in polys])
        ka[name + "_commitments"] = [g1_np_to_point(CO.best_multiexp(v, srs.g_lagrange, threads)) for v in vals]
    l0 = [0] * n
    l0[0] = 1
    lblind = [0] * n
    for i in range(n - BF, n):
        lblind[i] = 1
    llast = [0] * n
    llast[n - BF - 1] = 1
    ext = lambda v: CO.coeff_to_extended(CO.lagrange_to_coeff(ints_to_np(v), k, threads), k, dom.ext_k, threads)
    ka["l0"], ka["l_last"] = ext(l0), ext(llast)
    one_minus = ints_to_np([1] * dom.ext_n)
    ka["l_active"] = CO.fr_sub_arrays(CO.fr_sub_arrays(one_minus, ka["l_last"]), ext(lblind))
    h = hashlib.blake2b(digest_size=64, person=b"Halo2-Verify-Key")
    h.update(struct.pack("<I", k))
    for P in ka["fixed_commitments"] + ka["sigma_commitments"]:
        h.update(compress_point(P))
    ka["transcript_repr"] = int.from_bytes(h.digest(), "little") % R
    return ka


def keygen(layout, srs):
    """keygen_vk + keygen_pk as Python integer lists (what create_proof / verify_proof below work on)"""
    ka = keygen_arrays(layout, srs)
    pk = {"k": ka["k"], "dom": ka["dom"], "table_len": ka["table_len"], "transcript_repr": ka["transcript_repr"],
          "fixed_commitments": ka["fixed_commitments"], "sigma_commitments": ka["sigma_commitments"], "arrays": ka}
    for name in ("fixed_values", "fixed_polys", "fixed_cosets", "sigma_values", "sigma_polys", "sigma_cosets"):
        pk[name] = [np_to_ints(a) for a in ka[name]]
    for name in ("l0", "l_last", "l_active"):
        pk[name] = np_to_ints(ka[name])
    return pk


class _CKey(C.Structure):
    _fields_ = [("k", C.c_uint32), ("table_len", C.c_uint32)] + [(nm, C.c_void_p) for nm in (
        "fixed_values", "fixed_polys", "fixed_cosets", "sigma_values", "sigma_polys", "sigma_cosets", "l0", "l_last", "l_active",
        "g", "g_lagrange")] + [("transcript_repr", C.c_uint64 * 4)]


def create_proof_c(ka, srs, advice_np, seed, proof_index=0, nonce=0, threads=0, transcript_repr=None):
    """oracle/plonk_prover.c: the same create_proof in C (seconds at k = 17 / 18).  ka = keygen_arrays(..),
    advice_np = uint64[5, n, 4] Montgomery (CO.RsaTable.advice()).  -> (proof bytes, challenges dict)"""
    key = _CKey()
    key.k, key.table_len = ka["k"], ka["table_len"]
    keep = []
    for nm in ("fixed_values", "fixed_polys", "fixed_cosets", "sigma_values", "sigma_polys", "sigma_cosets", "l0", "l_last", "l_active"):
        a = np.ascontiguousarray(ka[nm], dtype=np.uint64)
        keep.append(a)
        setattr(key, nm, a.ctypes.data)
    g = np.ascontiguousarray(srs.g, dtype=np.uint64)
    gl = np.ascontiguousarray(srs.g_lagrange, dtype=np.uint64)
    key.g, key.g_lagrange = g.ctypes.data, gl.ctypes.data
    tr = ints_to_np([ka["transcript_repr"] if transcript_repr is None else transcript_repr])[0]
    for i in range(4):
        key.transcript_repr[i] = int(tr[i])
    adv = np.ascontiguousarray(advice_np, dtype=np.uint64)
    assert adv.shape == (NUM_ADVICE, 1 << ka["k"], 4)
    kb = struct.pack("<8I", *blind_key(seed))
    proof = np.zeros(proof_length(), dtype=np.uint8)
    chal = np.zeros((6, 4), dtype=np.uint64)
    L = CO.lib()
    L.orc_plonk_create_proof.restype = C.c_int
    rc = L.orc_plonk_create_proof(C.byref(key), _ptr(adv), kb, C.c_uint64(nonce), C.c_uint32(proof_index), C.c_int(threads or CO.ncores()),
                                  _ptr(proof), _ptr(chal))
    if rc == -1:
        raise ValueError("lookup input not in table (ConstraintSystemFailure)")
    assert rc == 0, rc
    return proof.tobytes(), dict(zip(("theta", "beta", "gamma", "y", "x", "v"), np_to_ints(chal)))


# ---- prover --------------------------------------------------------------------------------------------
def permute_expression_pair(A, S, u):
    """halo2 lookup::prover::permute_expression_pair on the first u rows"""
    a_sorted = sorted(A[:u])
    left = {}
    for v in S[:u]:
        left[v] = left.get(v, 0) + 1
    s_perm = [0] * u
    repeated = []
    for row, v in enumerate(a_sorted):
        if row == 0 or v != a_sorted[row - 1]:
            s_perm[row] = v
            if left.get(v, 0) == 0:
                raise ValueError("lookup input not in table (ConstraintSystemFailure)")
            left[v] -= 1
        else:
            repeated.append(row)
    for v in sorted(left):
        for _ in range(left[v]):
            s_perm[repeated.pop()] = v
    assert not repeated
    return a_sorted, s_perm


def batch_invert(xs):
    pref, acc = [], 1
    for x in xs:
        pref.append(acc)
        if x:
            acc = acc * x % R
    inv = pow(acc, -1, R)
    out = [0] * len(xs)
    for i in range(len(xs) - 1, -1, -1):
        if xs[i]:
            out[i] = inv * pref[i] % R
            inv = inv * xs[i] % R
    return out


def kate_division(coeffs, z):
    """(f(X) - f(z)) / (X - z)"""
    q = [0] * (len(coeffs) - 1)
    acc = 0
    for i in range(len(coeffs) - 1, 0, -1):
        acc = (coeffs[i] + acc * z) % R
        q[i - 1] = acc
    return q


def create_proof(pk, srs, advice, seed, proof_index=0, trace=None, nonce=0):
    """advice: 5 lists of n canonical values whose last BF+1 rows are ignored (re-blinded here).
    Returns the proof bytes.  `trace` (dict) receives intermediate values for stage-by-stage tests."""
    dom = pk["dom"]
    n, u, ext_n, step = dom.n, dom.n - (BF + 1), dom.ext_n, dom.step
    fx = pk["fixed_values"]
    tr = Transcript()
    tr.common_scalar(pk["transcript_repr"])
    # (no instance values: the sha-disabled circuit has an empty instance column)
    adv = [list(col[:u]) + blind_fes(seed, proof_index, ST_ADVICE + c, range(u, n), nonce) for c, col in enumerate(advice)]
    adv_polys = [dom.lagrange_to_coeff(v) for v in adv]
    for v in adv:
        tr.write_point(srs.commit_lagrange(v))
    theta = tr.squeeze()
    # lookups: permuted input / table
    lk = []
    table = [(fx[F_T_TAG][i] * theta + fx[F_T_VALUE][i]) % R for i in range(n)]
    for li, (acol, ftag, fsel) in enumerate(LOOKUPS):
        A = [(fx[ftag][i] * theta + fx[fsel][i] * adv[acol][i]) % R for i in range(n)]
        a_p, s_p = permute_expression_pair(A, table, u)
        a_p += blind_fes(seed, proof_index, ST_LOOKUP_A + li, range(u, n), nonce)
        s_p += blind_fes(seed, proof_index, ST_LOOKUP_S + li, range(u, n), nonce)
        tr.write_point(srs.commit_lagrange(a_p))
        tr.write_point(srs.commit_lagrange(s_p))
        lk.append({"A": A, "a_p": a_p, "s_p": s_p})
    beta = tr.squeeze()
    gamma = tr.squeeze()
    # permutation grand products
    inst = [0] * n
    colvals = adv + [inst]
    omega_pows, acc = [], 1
    for _ in range(n):
        omega_pows.append(acc)
        acc = acc * dom.omega % R
    perm_z = []
    last_z = 1
    for ch in range(0, len(PERM_COLUMNS), CHUNK):
        cols = list(range(ch, min(ch + CHUNK, len(PERM_COLUMNS))))
        den = [1] * n
        num = [1] * n
        for c in cols:
            sv, cv, dl = pk["sigma_values"][c], colvals[c], pow(O.DELTA, c, R)
            den = [d * ((beta * s + gamma + v) % R) % R for d, s, v in zip(den, sv, cv)]
            num = [m * ((dl * w % R * beta + gamma + v) % R) % R for m, w, v in zip(num, omega_pows, cv)]
        inv = batch_invert(den)
        z = [last_z]
        for row in range(1, n):
            z.append(z[-1] * num[row - 1] % R * inv[row - 1] % R)
        z[n - BF:] = blind_fes(seed, proof_index, ST_PERM_Z + ch // CHUNK, range(n - BF, n), nonce)
        last_z = z[n - BF - 1]
        perm_z.append(z)
    for z in perm_z:
        tr.write_point(srs.commit_lagrange(z))
    # lookup grand products
    for li, d in enumerate(lk):
        den = [((a + beta) % R) * ((s + gamma) % R) % R for a, s in zip(d["a_p"], d["s_p"])]
        inv = batch_invert(den)
        z = [1]
        for i in range(u):
            z.append(z[-1] * ((d["A"][i] + beta) % R) % R * ((table[i] + gamma) % R) % R * inv[i] % R)
        z = z[: n - BF] + blind_fes(seed, proof_index, ST_LOOKUP_Z + li, range(n - BF, n), nonce)
        d["z"] = z
        tr.write_point(srs.commit_lagrange(z))
    # vanishing argument: random polynomial
    random_poly = blind_fes(seed, proof_index, ST_RANDOM_POLY, range(n), nonce)
    tr.write_point(srs.commit(random_poly))
    y = tr.squeeze()
    # quotient on the extended domain
    perm_z_polys = [dom.lagrange_to_coeff(z) for z in perm_z]
    for d in lk:
        d["z_poly"] = dom.lagrange_to_coeff(d["z"])
        d["a_poly"] = dom.lagrange_to_coeff(d["a_p"])
        d["s_poly"] = dom.lagrange_to_coeff(d["s_p"])
    ae = [dom.coeff_to_extended(p) for p in adv_polys]
    pz = [dom.coeff_to_extended(p) for p in perm_z_polys]
    lz = [dom.coeff_to_extended(d["z_poly"]) for d in lk]
    la = [dom.coeff_to_extended(d["a_poly"]) for d in lk]
    ls = [dom.coeff_to_extended(d["s_poly"]) for d in lk]
    fe_, se_ = pk["fixed_cosets"], pk["sigma_cosets"]
    l0, l_last, l_active = pk["l0"], pk["l_last"], pk["l_active"]
    last_rot = -(BF + 1)
    nsets = len(perm_z)
    delta_pows = [pow(O.DELTA, j, R) for j in range(len(PERM_COLUMNS))]
    h = [0] * ext_n
    xpt = O.ZETA
    for i in range(ext_n):
        nx = (i + step) % ext_n
        pv = (i - step) % ext_n
        a, b, c, dd, e = ae[0][i], ae[1][i], ae[2][i], ae[3][i], ae[4][i]
        colv = (a, b, c, dd, e, 0)
        acc = (a * fe_[F_SA][i] + b * fe_[F_SB][i] + c * fe_[F_SC][i] + dd * fe_[F_SD][i] + e * fe_[F_SE][i]
               + a * b % R * fe_[F_MUL_AB][i] + c * dd % R * fe_[F_MUL_CD][i] + ae[4][nx] * fe_[F_SE_NEXT][i] + fe_[F_CONST][i]) % R
        # permutation
        acc = (acc * y + l0[i] * (1 - pz[0][i])) % R
        zl = pz[nsets - 1][i]
        acc = (acc * y + l_last[i] * (zl * zl - zl)) % R
        for s in range(1, nsets):
            acc = (acc * y + l0[i] * (pz[s][i] - pz[s - 1][(i + last_rot * step) % ext_n])) % R
        for s in range(nsets):
            left = pz[s][nx]
            right = pz[s][i]
            for cidx in range(s * CHUNK, min((s + 1) * CHUNK, len(PERM_COLUMNS))):
                left = left * ((colv[cidx] + beta * se_[cidx][i] + gamma) % R) % R
                right = right * ((colv[cidx] + beta * xpt % R * delta_pows[cidx] + gamma) % R) % R
            acc = (acc * y + l_active[i] * (left - right)) % R
        # lookups
        tbl = (fe_[F_T_TAG][i] * theta + fe_[F_T_VALUE][i]) % R
        for li, (acol, ftag, fsel) in enumerate(LOOKUPS):
            z_, zn_, ap, sp, apv = lz[li][i], lz[li][nx], la[li][i], ls[li][i], la[li][pv]
            inp = (fe_[ftag][i] * theta + fe_[fsel][i] * colv[acol]) % R
            acc = (acc * y + l0[i] * (1 - z_)) % R
            acc = (acc * y + l_last[i] * (z_ * z_ - z_)) % R
            acc = (acc * y + l_active[i] * (zn_ * ((ap + beta) % R) % R * ((sp + gamma) % R)
                                            - z_ * ((inp + beta) % R) % R * ((tbl + gamma) % R))) % R
            acc = (acc * y + l0[i] * (ap - sp)) % R
            acc = (acc * y + l_active[i] * ((ap - sp) % R * ((ap - apv) % R) % R)) % R
        h[i] = acc * dom.t_inv[i % step] % R
        xpt = xpt * dom.ext_omega % R
    h_coeffs = dom.extended_to_coeff(h)
    h_pieces = [h_coeffs[j * n:(j + 1) * n] for j in range(dom.qd)]
    for piece in h_pieces:
        tr.write_point(srs.commit(piece))
    x = tr.squeeze()
    xn = pow(x, n, R)
    # evaluations
    for col, rot in ADVICE_QUERIES:
        tr.write_scalar(eval_poly(adv_polys[col], dom.rotate_omega(x, rot)))
    for col, rot in FIXED_QUERIES:
        tr.write_scalar(eval_poly(pk["fixed_polys"][col], dom.rotate_omega(x, rot)))
    h_poly = [0] * n
    for piece in reversed(h_pieces):
        h_poly = [(a * xn + b) % R for a, b in zip(h_poly, piece)]
    tr.write_scalar(eval_poly(random_poly, x))
    for p in pk["sigma_polys"]:
        tr.write_scalar(eval_poly(p, x))
    x_next, x_inv, x_last = dom.rotate_omega(x, 1), dom.rotate_omega(x, -1), dom.rotate_omega(x, last_rot)
    for s, p in enumerate(perm_z_polys):
        tr.write_scalar(eval_poly(p, x))
        tr.write_scalar(eval_poly(p, x_next))
        if s != nsets - 1:
            tr.write_scalar(eval_poly(p, x_last))
    for d in lk:
        tr.write_scalar(eval_poly(d["z_poly"], x))
        tr.write_scalar(eval_poly(d["z_poly"], x_next))
        tr.write_scalar(eval_poly(d["a_poly"], x))
        tr.write_scalar(eval_poly(d["a_poly"], x_inv))
        tr.write_scalar(eval_poly(d["s_poly"], x))
    # multiopen (GWC): queries in halo2's chain order, grouped by point in order of first appearance
    queries = []
    for col, rot in ADVICE_QUERIES:
        queries.append((dom.rotate_omega(x, rot), adv_polys[col]))
    for s, p in enumerate(perm_z_polys):
        queries += [(x, p), (x_next, p)]
    for s in range(nsets - 2, -1, -1):
        queries.append((x_last, perm_z_polys[s]))
    for d in lk:
        queries += [(x, d["z_poly"]), (x, d["a_poly"]), (x, d["s_poly"]), (x_inv, d["a_poly"]), (x_next, d["z_poly"])]
    for col, rot in FIXED_QUERIES:
        queries.append((dom.rotate_omega(x, rot), pk["fixed_polys"][col]))
    for p in pk["sigma_polys"]:
        queries.append((x, p))
    queries += [(x, h_poly), (x, random_poly)]
    v = tr.squeeze()
    points = []
    for pt, _ in queries:
        if pt not in points:
            points.append(pt)
    for pt in points:
        batch = [0] * n
        for qp, poly in queries:
            if qp == pt:
                batch = [(a * v + b) % R for a, b in zip(batch, poly)]
        tr.write_point(srs.commit(kate_division(batch, pt)))
    if trace is not None:
        trace.update({"theta": theta, "beta": beta, "gamma": gamma, "y": y, "x": x, "v": v, "adv": adv, "lookups": lk,
                      "perm_z": perm_z, "h_pieces": h_pieces, "random_poly": random_poly})
    return bytes(tr.out)


# ---- verifier ----------------------------------------------------------------------------------------------
def _pt_add(P, Qp):
    return O.g1_add(P, Qp)


def _pt_mul(P, k):
    return O.g1_mul(P, k % R) if P is not None else None


def verify_proof(pk, srs_secret, proof):
    """verify_proof with VerifierGWC; the final pairing check is replaced by its equivalent under the
    known SRS secret.  Returns True / False (malformed input raises ValueError)."""
    dom = pk["dom"]
    n = dom.n
    tr = Transcript(proof)
    tr.common_scalar(pk["transcript_repr"])
    adv_c = [tr.read_point() for _ in range(NUM_ADVICE)]
    theta = tr.squeeze()
    lk_c = [{"a": tr.read_point(), "s": tr.read_point()} for _ in LOOKUPS]
    beta = tr.squeeze()
    gamma = tr.squeeze()
    nsets = (len(PERM_COLUMNS) + CHUNK - 1) // CHUNK
    pz_c = [tr.read_point() for _ in range(nsets)]
    for d in lk_c:
        d["z"] = tr.read_point()
    random_c = tr.read_point()
    y = tr.squeeze()
    h_c = [tr.read_point() for _ in range(dom.qd)]
    x = tr.squeeze()
    adv_e = [tr.read_scalar() for _ in ADVICE_QUERIES]
    fix_e = [tr.read_scalar() for _ in FIXED_QUERIES]
    random_e = tr.read_scalar()
    sig_e = [tr.read_scalar() for _ in PERM_COLUMNS]
    pz_e = []
    for s in range(nsets):
        d = {"z": tr.read_scalar(), "z_next": tr.read_scalar()}
        if s != nsets - 1:
            d["z_last"] = tr.read_scalar()
        pz_e.append(d)
    lk_e = [{"z": tr.read_scalar(), "z_next": tr.read_scalar(), "a": tr.read_scalar(), "a_inv": tr.read_scalar(), "s": tr.read_scalar()}
            for _ in LOOKUPS]
    # vanishing identity at x
    xn = pow(x, n, R)
    l0 = dom.l_i(x, 0)
    l_last = dom.l_i(x, -(BF + 1))
    l_blind = sum(dom.l_i(x, -r) for r in range(1, BF + 1)) % R
    l_active = (1 - l_last - l_blind) % R
    a, b, c, dd, e, e_next = adv_e
    colv = (a, b, c, dd, e, 0)  # instance column: no public inputs
    acc = (a * fix_e[F_SA] + b * fix_e[F_SB] + c * fix_e[F_SC] + dd * fix_e[F_SD] + e * fix_e[F_SE] + a * b % R * fix_e[F_MUL_AB]
           + c * dd % R * fix_e[F_MUL_CD] + e_next * fix_e[F_SE_NEXT] + fix_e[F_CONST]) % R
    acc = (acc * y + l0 * (1 - pz_e[0]["z"])) % R
    zl = pz_e[nsets - 1]["z"]
    acc = (acc * y + l_last * (zl * zl - zl)) % R
    for s in range(1, nsets):
        acc = (acc * y + l0 * (pz_e[s]["z"] - pz_e[s - 1]["z_last"])) % R
    for s in range(nsets):
        left, right = pz_e[s]["z_next"], pz_e[s]["z"]
        for cidx in range(s * CHUNK, min((s + 1) * CHUNK, len(PERM_COLUMNS))):
            left = left * ((colv[cidx] + beta * sig_e[cidx] + gamma) % R) % R
            right = right * ((colv[cidx] + beta * x % R * pow(O.DELTA, cidx, R) + gamma) % R) % R
        acc = (acc * y + l_active * (left - right)) % R
    tbl = (fix_e[F_T_TAG] * theta + fix_e[F_T_VALUE]) % R
    for (acol, ftag, fsel), ev in zip(LOOKUPS, lk_e):
        inp = (fix_e[ftag] * theta + fix_e[fsel] * colv[acol]) % R
        acc = (acc * y + l0 * (1 - ev["z"])) % R
        acc = (acc * y + l_last * (ev["z"] * ev["z"] - ev["z"])) % R
        acc = (acc * y + l_active * (ev["z_next"] * ((ev["a"] + beta) % R) % R * ((ev["s"] + gamma) % R)
                                     - ev["z"] * ((inp + beta) % R) % R * ((tbl + gamma) % R))) % R
        acc = (acc * y + l0 * (ev["a"] - ev["s"])) % R
        acc = (acc * y + l_active * ((ev["a"] - ev["s"]) % R * ((ev["a"] - ev["a_inv"]) % R) % R)) % R
    expected_h = acc * pow((xn - 1) % R, -1, R) % R
    h_commit = None
    for P in reversed(h_c):
        h_commit = _pt_add(_pt_mul(h_commit, xn), P)
    # queries in the prover's order: (point, commitment, eval)
    x_next, x_inv, x_last = dom.rotate_omega(x, 1), dom.rotate_omega(x, -1), dom.rotate_omega(x, -(BF + 1))
    queries = []
    for (col, rot), ev in zip(ADVICE_QUERIES, adv_e):
        queries.append((dom.rotate_omega(x, rot), adv_c[col], ev))
    for s in range(nsets):
        queries += [(x, pz_c[s], pz_e[s]["z"]), (x_next, pz_c[s], pz_e[s]["z_next"])]
    for s in range(nsets - 2, -1, -1):
        queries.append((x_last, pz_c[s], pz_e[s]["z_last"]))
    for d, ev in zip(lk_c, lk_e):
        queries += [(x, d["z"], ev["z"]), (x, d["a"], ev["a"]), (x, d["s"], ev["s"]), (x_inv, d["a"], ev["a_inv"]), (x_next, d["z"], ev["z_next"])]
    for (col, rot), ev in zip(FIXED_QUERIES, fix_e):
        queries.append((dom.rotate_omega(x, rot), pk["fixed_commitments"][col], ev))
    for cidx in range(len(PERM_COLUMNS)):
        queries.append((x, pk["sigma_commitments"][cidx], sig_e[cidx]))
    queries += [(x, h_commit, expected_h), (x, random_c, random_e)]
    v = tr.squeeze()
    points = []
    for pt, _, _ in queries:
        if pt not in points:
            points.append(pt)
    ok = True
    for pt in points:
        Cb, eb = None, 0
        for qp, cm, ev in queries:
            if qp == pt:
                Cb = _pt_add(_pt_mul(Cb, v), cm)
                eb = (eb * v + ev) % R
        W = tr.read_point()
        lhs = _pt_mul(W, (srs_secret - pt) % R)
        rhs = _pt_add(Cb, O.g1_neg(_pt_mul(O.G1_GEN, eb)) if eb else None)
        ok = ok and (lhs == rhs)
    if tr.pos != len(proof):
        return False
    return ok


def vk_from_commitments(k, fixed_commitments, sigma_commitments, transcript_repr):
    """the part of the key verify_proof needs, from an exported verifying key (points as (x, y) / None)"""
    return {"k": k, "dom": Domain(k), "fixed_commitments": list(fixed_commitments), "sigma_commitments": list(sigma_commitments),
            "transcript_repr": transcript_repr}


def proof_length(k=None):
    nsets = (len(PERM_COLUMNS) + CHUNK - 1) // CHUNK
    points = NUM_ADVICE + 2 * len(LOOKUPS) + nsets + len(LOOKUPS) + 1 + 4 + 4
    scalars = len(ADVICE_QUERIES) + len(FIXED_QUERIES) + 1 + len(PERM_COLUMNS) + (3 * nsets - 1) + 5 * len(LOOKUPS)
    return 32 * (points + scalars)
