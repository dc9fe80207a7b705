"""GPU parity: keygen and the full prover (b2r_rsa_keygen / b2r_rsa_prove_batch, through the C ABI) against the
oracle's independent restatement of halo2's keygen_vk / keygen_pk / create_proof / verify_proof (oracle/plonk.py).

  * RSA-512 at k = 14 (a size the Python oracle proves in seconds): verifying key and PROOF BYTES are identical
    for the same seeded randomness - every commitment, challenge, evaluation and opening witness of the proof.
  * RSA-2048 at k = 17 (BASELINE configs[0]/[1]) and RSA-4096 at k = 18 (configs[2]): verifying key and ALL 2848
    PROOF BYTES identical to the C restatement of the same prover (oracle/plonk_prover.c, itself pinned to plonk.py
    byte for byte at k = 14 by tests/test_oracle_plonk_c.py); the proofs are accepted by the oracle verifier, a
    tampered proof and a proof for a wrong signature are rejected.
  * batches that cross the quotient sub-batch (16), a ragged last sub-batch and several proof groups: every proof
    verified, the proofs at the seams compared byte for byte.
  * two contexts in one process, calls interleaved.
"""
import numpy as np
import pytest

import bn254 as O
import cpu_oracle as CO
import plonk as PL
import rsa_fixtures as RF
from util import fr_to_np, np_to_fr, np_to_g1

pytestmark = pytest.mark.gpu


def _setup(ctx, bits, k):
    prog = ctx.rsa_program(bits, k)
    g, gl = ctx.srs_setup(k, fr_to_np([O.srs_secret(k)])[0])
    pk = ctx.rsa_keygen(prog, g, gl)
    return prog, g, gl, pk


def _vk(pk):
    f, s, t = pk.export_vk()
    return PL.vk_from_commitments(pk.k, np_to_g1(f), np_to_g1(s), np_to_fr(t.reshape(1, 4))[0])


@pytest.fixture(scope="module")
def small(ctx):
    bits, k = 512, 14
    prog, g, gl, pk = _setup(ctx, bits, k)
    srs = PL.Srs(k)
    opk = PL.keygen(PL.circuit_layout(bits, k), srs)
    yield bits, k, pk, srs, opk
    pk.free(); g.free(); gl.free(); prog.free()


def test_keygen_matches_oracle(small):
    bits, k, pk, srs, opk = small
    vk = _vk(pk)
    assert pk.proof_bytes == PL.proof_length() == 2848
    assert vk["fixed_commitments"] == opk["fixed_commitments"]
    assert vk["sigma_commitments"] == opk["sigma_commitments"]
    assert vk["transcript_repr"] == opk["transcript_repr"]


def test_proof_bytes_match_oracle(small):
    bits, k, pk, srs, opk = small
    seed, batch = 0xB200, 2
    nl, sl, hl = RF.batch(bits, batch, start=1)
    proofs, status = pk.prove_batch(nl, sl, hl, seed, nonce=3)
    assert status.tolist() == [1] * batch
    for i in range(batch):
        v, adv, rows, bad, msg = CO.rsa_synthesize(bits, k, *RF.instance(bits, 1 + i))
        assert v == 1 and bad == 0
        want = PL.create_proof(opk, srs, [PL.np_to_ints(adv[c]) for c in range(5)], seed, proof_index=i, nonce=3)
        got = bytes(proofs[i])
        if got != want:   # name the first differing proof element
            first = next(j for j in range(0, len(want), 32) if got[j:j + 32] != want[j:j + 32]) // 32
            raise AssertionError(f"instance {i}: proof element {first} of {len(want) // 32} differs")
        assert PL.verify_proof(opk, srs.s, got)


_ORACLE_CACHE = {}


def _oracle_key(bits, k):
    """SRS + keygen of the oracle at (bits, k), cached for the session (tens of seconds at k = 17 / 18)"""
    if (bits, k) not in _ORACLE_CACHE:
        srs = PL.Srs(k)
        _ORACLE_CACHE[(bits, k)] = (srs, PL.keygen_arrays(PL.circuit_layout(bits, k), srs))
    return _ORACLE_CACHE[(bits, k)]


def _assert_same_proof(got, want, what):
    if got != want:   # name the first differing proof element
        first = next(j for j in range(0, len(want), 32) if got[j:j + 32] != want[j:j + 32]) // 32
        raise AssertionError(f"{what}: proof element {first} of {len(want) // 32} differs")


@pytest.mark.parametrize("bits,k,count", [(2048, 17, 2), (4096, 18, 1)])
def test_proof_bytes_match_oracle_at_baseline_sizes(ctx, bits, k, count):
    """BASELINE configs[1] (RSA-2048, k = 17) and configs[2] (RSA-4096, k = 18): vk commitments and all 2848 proof bytes
    equal to the oracle prover's for the same key / nonce / instance index"""
    prog, g, gl, pk = _setup(ctx, bits, k)
    srs, ka = _oracle_key(bits, k)
    vk = _vk(pk)
    assert vk["fixed_commitments"] == ka["fixed_commitments"]
    assert vk["sigma_commitments"] == ka["sigma_commitments"]
    assert vk["transcript_repr"] == ka["transcript_repr"]
    start, seed, nonce = 5, bytes(range(100, 132)), 0x1122334455
    nl, sl, hl = RF.batch(bits, count, start=start)
    proofs, status = pk.prove_batch(nl, sl, hl, seed, nonce=nonce)
    assert status.tolist() == [1] * count
    for i in range(count):
        v, adv, rows, bad, msg = CO.rsa_synthesize(bits, k, *RF.instance(bits, start + i))
        assert v == 1 and bad == 0
        want, _ = PL.create_proof_c(ka, srs, adv, seed, proof_index=i, nonce=nonce)
        _assert_same_proof(bytes(proofs[i]), want, f"RSA-{bits} k={k} instance {i}")
        assert PL.verify_proof(vk, srs.s, want)
    pk.free(); g.free(); gl.free(); prog.free()


@pytest.mark.parametrize("batch,group", [(20, None), (70, None), (9, 4)])
def test_sub_batches_and_groups(small, monkeypatch, batch, group):
    """batch 20: quotient sub-batches of 16 + a ragged one of 4 (the per-column coset fallback); batch 70: two proof
    groups (64 + 6, instance index offset p_base > 0); B2R_PROVE_GROUP=4: three groups.  Every proof must verify; the
    proofs on both sides of every seam are compared with the oracle prover byte for byte."""
    bits, k, pk, srs, opk = small
    if group:
        monkeypatch.setenv("B2R_PROVE_GROUP", str(group))
    nl, sl, hl = RF.batch(bits, batch)
    proofs, status = pk.prove_batch(nl, sl, hl, 0xABCD, nonce=9)
    assert status.tolist() == [1] * batch
    for i in range(batch):
        assert PL.verify_proof(opk, srs.s, bytes(proofs[i])), f"proof {i} of {batch} rejected"
    assert len({bytes(p) for p in proofs}) == batch
    seams = {0, batch - 1} | ({15, 16} if batch > 16 else set()) | ({63, 64} if batch > 64 else set()) | ({3, 4, 7, 8} if group else set())
    for i in sorted(seams):
        v, adv, rows, bad, msg = CO.rsa_synthesize(bits, k, *RF.instance(bits, i))
        want, _ = PL.create_proof_c(opk["arrays"], srs, adv, 0xABCD, proof_index=i, nonce=9)
        _assert_same_proof(bytes(proofs[i]), want, f"batch {batch} instance {i}")


def test_two_contexts_interleaved(ctx, small):
    """a second context in the same process (the header's one-context-per-thread model): programs, SRS, keys and proofs
    built on both with the calls interleaved give identical results, and the first context still works after the
    second is destroyed.  (Per-device function attributes / device binding: ctx.hpp DeviceGuard.)"""
    import b2rsa
    bits, k, pk, srs, opk = small
    other = b2rsa.Context(0)
    prog2 = other.rsa_program(bits, k)
    g2, gl2 = other.srs_setup(k, fr_to_np([O.srs_secret(k)])[0])
    nl, sl, hl = RF.batch(bits, 2, start=3)
    a1, _ = pk.prove_batch(nl[:1], sl[:1], hl[:1], 5, nonce=1)
    pk2 = other.rsa_keygen(prog2, g2, gl2)
    b1, _ = pk2.prove_batch(nl[:1], sl[:1], hl[:1], 5, nonce=1)
    a2, _ = pk.prove_batch(nl[1:], sl[1:], hl[1:], 5, nonce=2)
    b2, _ = pk2.prove_batch(nl[1:], sl[1:], hl[1:], 5, nonce=2)
    assert bytes(a1[0]) == bytes(b1[0]) and bytes(a2[0]) == bytes(b2[0])
    # uniform MSM (binned sort: 80 KB dynamic shared memory attribute) on both contexts
    from util import random_fr_np
    import torch
    n = 1 << k
    sc = torch.from_numpy(random_fr_np(n, 4).view(np.int64)).cuda()
    o1 = torch.zeros(8, dtype=torch.int64, device="cuda"); o2 = torch.zeros(8, dtype=torch.int64, device="cuda")
    gx, glx = ctx.srs_setup(k, fr_to_np([O.srs_secret(k)])[0])
    ctx.msm_batch_dev(glx, sc.data_ptr(), 1, n, o1.data_ptr(), uniform=True)
    other.msm_batch_dev(gl2, sc.data_ptr(), 1, n, o2.data_ptr(), uniform=True)
    ctx.sync(); other.sync()
    assert torch.equal(o1, o2)
    pk2.free(); prog2.free(); g2.free(); gl2.free(); gx.free(); glx.free()
    other.close()
    a3, st = pk.prove_batch(nl[:1], sl[:1], hl[:1], 5, nonce=1)
    assert bytes(a3[0]) == bytes(a1[0]) and PL.verify_proof(opk, srs.s, bytes(a3[0]))


def test_one_context_per_device_in_one_process(ctx, small):
    """the multi-GPU shape a Rust host uses: one context per device in ONE process, calls
    interleaved from one thread whose current device never changes.  Needs a 2-GPU lease (gpurun --gpus 2)."""
    import torch
    import b2rsa
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    bits, k, pk, srs, opk = small
    torch.cuda.set_device(0)
    other = b2rsa.Context(1)
    prog1 = other.rsa_program(bits, k)
    g1, gl1 = other.srs_setup(k, fr_to_np([O.srs_secret(k)])[0])
    pk1 = other.rsa_keygen(prog1, g1, gl1)
    assert torch.cuda.current_device() == 0            # the library restored the caller's device
    nl, sl, hl = RF.batch(bits, 3, start=11)
    a, sa = pk.prove_batch(nl, sl, hl, 21, nonce=7)     # device 0
    b, sb = pk1.prove_batch(nl, sl, hl, 21, nonce=7)    # device 1
    a2, _ = pk.prove_batch(nl[:1], sl[:1], hl[:1], 21, nonce=7)
    assert sa.tolist() == sb.tolist() == [1, 1, 1]
    assert a.tobytes() == b.tobytes() and bytes(a2[0]) == bytes(a[0])
    assert _vk(pk)["fixed_commitments"] == _vk(pk1)["fixed_commitments"]
    for i in range(3):
        assert PL.verify_proof(opk, srs.s, bytes(b[i]))
    pk1.free(); prog1.free(); g1.free(); gl1.free()
    other.close()
    assert torch.cuda.current_device() == 0


def test_keygen_rejects_wrong_srs_size(ctx):
    """a Lagrange basis of another domain must not be accepted silently"""
    import b2rsa
    prog = ctx.rsa_program(512, 14)
    g15, gl15 = ctx.srs_setup(15, fr_to_np([O.srs_secret(15)])[0])
    g13, gl13 = ctx.srs_setup(13, fr_to_np([O.srs_secret(13)])[0])
    g14, gl14 = ctx.srs_setup(14, fr_to_np([O.srs_secret(14)])[0])
    for g, gl in ((g15, gl15), (g14, gl15), (g13, gl14), (g14, gl13)):
        with pytest.raises(b2rsa.B2RError) as e:
            ctx.rsa_keygen(prog, g, gl)
        assert e.value.code == b2rsa.ERR_INVALID
    pk = ctx.rsa_keygen(prog, g15, gl14)     # a longer g is fine: commit only uses its first 2^k points
    pk.free()
    for b in (g13, gl13, g14, gl14, g15, gl15):
        b.free()
    prog.free()


def test_transcript_repr_override(small):
    """b2r_pk_set_transcript_repr (what a Rust host passes for the real vk): proofs follow the installed value"""
    import b2rsa
    bits, k, pk, srs, opk = small
    f, s_, t0 = pk.export_vk()
    nl, sl, hl = RF.batch(bits, 1)
    new_repr = 0x1234567890ABCDEF1234567890ABCDEF
    pk.set_transcript_repr(fr_to_np([new_repr])[0])
    try:
        assert np_to_fr(pk.export_vk()[2].reshape(1, 4))[0] == new_repr
        proofs, _ = pk.prove_batch(nl, sl, hl, 3, nonce=4)
        v, adv, rows, bad, msg = CO.rsa_synthesize(bits, k, *RF.instance(bits, 0))
        want, _ = PL.create_proof_c(opk["arrays"], srs, adv, 3, proof_index=0, nonce=4, transcript_repr=new_repr)
        _assert_same_proof(bytes(proofs[0]), want, "transcript_repr override")
        assert not PL.verify_proof(opk, srs.s, want) and PL.verify_proof(dict(opk, transcript_repr=new_repr), srs.s, want)
        with pytest.raises(b2rsa.B2RError):
            pk.set_transcript_repr(np.array([2**64 - 1] * 4, dtype=np.uint64))   # not reduced
    finally:
        pk.set_transcript_repr(t0)


def test_full_size_proofs_verify(ctx):
    bits, k = 2048, 17
    prog, g, gl, pk = _setup(ctx, bits, k)
    vk = _vk(pk)
    secret = O.srs_secret(k)
    nl, sl, hl = RF.batch(bits, 3)
    hl_bad = hl.copy()
    hl_bad[2, 0] ^= np.uint64(1)                      # third instance: wrong message hash
    proofs, status = pk.prove_batch(nl, sl, hl_bad, seed=7, nonce=11)
    assert status.tolist() == [1, 1, 0]
    assert PL.verify_proof(vk, secret, bytes(proofs[0]))
    assert PL.verify_proof(vk, secret, bytes(proofs[1]))
    assert not PL.verify_proof(vk, secret, bytes(proofs[2]))          # is_valid = 0 violates assert_one
    assert bytes(proofs[0]) != bytes(proofs[1])
    bad = bytearray(proofs[0])
    bad[32 * 40 + 3] ^= 1                                              # one evaluation
    assert not PL.verify_proof(vk, secret, bytes(bad))
    # same key, nonce and inputs -> same bytes; other nonce / other seed -> other blinding -> other proof, still valid
    again, _ = pk.prove_batch(nl[:1], sl[:1], hl[:1], seed=7, nonce=11)
    assert bytes(again[0]) == bytes(proofs[0])
    for kw in ({"seed": 7, "nonce": 12}, {"seed": 8, "nonce": 11}, {"seed": bytes(range(32)), "nonce": 11}):
        other, _ = pk.prove_batch(nl[:1], sl[:1], hl[:1], **kw)
        assert bytes(other[0]) != bytes(proofs[0]) and PL.verify_proof(vk, secret, bytes(other[0]))
    # the plain 64-bit-seed entry point takes a fresh nonce from the context on every call: a reused seed never
    # repeats a blinding stream
    a, _ = pk.prove_batch(nl[:1], sl[:1], hl[:1], seed=7)
    b, _ = pk.prove_batch(nl[:1], sl[:1], hl[:1], seed=7)
    assert bytes(a[0]) != bytes(b[0]) and PL.verify_proof(vk, secret, bytes(a[0])) and PL.verify_proof(vk, secret, bytes(b[0]))
    pk.free(); g.free(); gl.free(); prog.free()


@pytest.mark.parametrize("bits,k", [(1024, 15), (4096, 18)])
def test_other_key_sizes_verify(ctx, bits, k):
    """BASELINE configs[2] (RSA-4096, k=18) and the reference's own enabled bench size (RSA-1024, k=15)"""
    prog, g, gl, pk = _setup(ctx, bits, k)
    vk = _vk(pk)
    nl, sl, hl = RF.batch(bits, 2)
    proofs, status = pk.prove_batch(nl, sl, hl, seed=99)
    assert status.tolist() == [1, 1]
    for i in range(2):
        assert PL.verify_proof(vk, O.srs_secret(k), bytes(proofs[i]))
    pk.free(); g.free(); gl.free(); prog.free()


def test_prover_argument_errors(small):
    import b2rsa
    bits, k, pk, srs, opk = small
    nl, sl, hl = RF.batch(bits, 1)
    with pytest.raises(b2rsa.B2RError):
        pk.prove_batch(nl, sl, hl, seed=0)


def test_empty_batches_and_argument_errors(ctx, small):
    """edge cases at the ABI: empty batches are no-ops that return OK, every misuse comes back as a status code with a
    message (nothing aborts or throws across the boundary)"""
    import ctypes as C
    import b2rsa
    bits, k, pk, srs, opk = small
    lib, h = ctx.lib, ctx.h
    nl, sl, hl = RF.batch(bits, 1)
    p = lambda a: C.c_void_p(a.ctypes.data)
    proofs = np.zeros(pk.proof_bytes, dtype=np.uint8)
    status = np.full(1, 7, dtype=np.uint8)
    key = (C.c_uint8 * 32)(*range(32))
    # empty batch: OK, outputs untouched
    assert lib.b2r_rsa_prove_batch(h, pk.h, p(nl), p(sl), p(hl), 0, 5, p(proofs), p(status)) == 0
    assert lib.b2r_rsa_prove_batch_ex(h, pk.h, p(nl), p(sl), p(hl), 0, C.cast(key, C.c_void_p), 1, 0, p(proofs), p(status)) == 0
    assert status[0] == 7 and not proofs.any()
    adv = np.zeros((5, 1 << k, 4), dtype=np.uint64)
    assert lib.b2r_rsa_witness_batch(h, pk.prog.h, p(nl), p(sl), p(hl), 0, 0, p(adv), p(status)) == 0
    out = np.zeros(8, dtype=np.uint64)
    g, gl = ctx.srs_setup(10, fr_to_np([O.srs_secret(10)])[0])
    assert lib.b2r_msm_g1_batch(h, g.h, None, 0, 0, p(out)) == 0                       # m = 0
    sc = np.zeros((1, 4), dtype=np.uint64)
    assert lib.b2r_msm_g1_batch(h, g.h, p(sc), 1, 0, p(out)) == 0 and not out.any()    # n = 0: the identity
    # misuse
    assert lib.b2r_rsa_prove_batch_ex(h, pk.h, p(nl), p(sl), p(hl), 1, None, 1, 0, p(proofs), p(status)) == b2rsa.ERR_INVALID
    assert b"seed" in lib.b2r_last_error(h)
    assert lib.b2r_rsa_prove_batch_ex(h, pk.h, p(nl), p(sl), p(hl), 1, C.cast(key, C.c_void_p), 1, 64, p(proofs), p(status)) == b2rsa.ERR_INVALID
    zero = (C.c_uint8 * 32)()
    assert lib.b2r_rsa_prove_batch_ex(h, pk.h, p(nl), p(sl), p(hl), 1, C.cast(zero, C.c_void_p), 1, 2, p(proofs), p(status)) == b2rsa.ERR_INVALID
    assert lib.b2r_rsa_prove_batch(h, None, p(nl), p(sl), p(hl), 1, 5, p(proofs), p(status)) == b2rsa.ERR_INVALID
    assert lib.b2r_rsa_prove_batch(None, pk.h, p(nl), p(sl), p(hl), 1, 5, p(proofs), p(status)) == b2rsa.ERR_INVALID
    assert lib.b2r_msm_g1_batch(h, g.h, p(sc), 1, 1 << 11, p(out)) == b2rsa.ERR_INVALID   # more scalars than bases
    assert lib.b2r_pk_set_transcript_repr(None, p(sc)) == b2rsa.ERR_INVALID
    # a 32-byte all-zero KEY is a legitimate key (only the 64-bit seed form rejects zero)
    assert lib.b2r_rsa_prove_batch_ex(h, pk.h, p(nl), p(sl), p(hl), 1, C.cast(zero, C.c_void_p), 1, 0, p(proofs), p(status)) == 0
    assert status[0] == 1 and PL.verify_proof(opk, srs.s, proofs.tobytes())
    g.free(); gl.free()


def test_device_resident_commitment_block(ctx, small, monkeypatch):
    """b2r_last_commitments: the 31 commitments per proof that stay on the device for the multi-GPU all-gather
     are exactly the group elements of the proof stream - compressed, they equal the proof's first
    27 points and its last 4 - for a batch that crosses a proof group, and a device-to-device copy returns the
    same block as the host copy"""
    import torch
    bits, k, pk, srs, opk = small
    monkeypatch.setenv("B2R_PROVE_GROUP", "2")          # three proofs in groups of 2: the block spans proof groups
    nl, sl, hl = RF.batch(bits, 3, start=4)
    proofs, status = pk.prove_batch(nl, sl, hl, seed=77)
    assert status.tolist() == [1, 1, 1]
    assert ctx.last_commitments_info() == (3, 31)
    cm = ctx.last_commitments()
    for p in range(3):
        pts = np_to_g1(cm[p].reshape(-1, 8))
        raw = bytes(proofs[p])
        want = [raw[32 * i:32 * i + 32] for i in range(27)] + [raw[2848 - 128 + 32 * i:2848 - 96 + 32 * i] for i in range(4)]
        assert [PL.compress_point(P) for P in pts] == want, f"proof {p}"
    d = torch.zeros(3 * 31 * 8, dtype=torch.int64, device="cuda:0")
    ctx.last_commitments_dev(d.data_ptr(), 3 * 31)
    ctx.sync()
    assert np.array_equal(d.cpu().numpy().view(np.uint64).reshape(3, 31, 8), cm)
    import b2rsa
    with pytest.raises(b2rsa.B2RError):
        ctx.last_commitments_dev(d.data_ptr(), 3 * 31 - 1)   # destination too small
