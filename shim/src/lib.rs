//! Safe wrappers over libb2rsa.so for the reference's bench / example (uncompiled here: see Cargo.toml).
//!
//! What a maintainer of the reference gets:
//!  * `Gpu`            one context per device (b2r_ctx); `Gpu::thread_local()` for halo2's rayon workers
//!  * `best_multiexp`, `best_fft`   drop-in bodies for halo2_proofs::arithmetic (patches/halo2_proofs_arithmetic.patch)
//!  * `RsaProver`      keygen + create_proof for batches of pkcs1v15 instances: the call that replaces the
//!                     `create_proof` loop of benches/bench.rs:319-331; proofs verify with the stock `verify_proof`
//!                     against the Rust-side `keygen_vk` of the UNCHANGED circuit (tests/vk_matches.rs)
//!  * `vk_transcript_repr`   recovers halo2's private `VerifyingKey::transcript_repr` through `hash_into`
pub mod sys;

use halo2wrong::curves::bn256::{Bn256, Fr, G1Affine, G1};
use halo2wrong::curves::group::Curve;
use halo2wrong::halo2::arithmetic::{g_to_lagrange, Field};
use halo2wrong::halo2::plonk::VerifyingKey;
use halo2wrong::halo2::poly::commitment::{Params, ParamsProver};
use halo2wrong::halo2::poly::kzg::commitment::ParamsKZG;
use halo2wrong::halo2::transcript::{ChallengeScalar, EncodedChallenge, Transcript};
use std::cell::RefCell;
use std::collections::HashMap;
use std::ffi::CStr;
use std::io;
use std::ptr;

#[derive(Debug)]
pub struct Error(pub i32, pub String);
pub type Result<T> = std::result::Result<T, Error>;

/// One b2r_ctx: bound to one device and one stream; not Sync (one context per thread, include/b2rsa.h).
pub struct Gpu {
    raw: *mut sys::b2r_ctx,
    bases: RefCell<HashMap<(usize, usize), *mut sys::b2r_bases>>, // (ptr, len) of a &[G1Affine] -> resident table
}

impl Gpu {
    pub fn new(device: i32) -> Result<Self> {
        let mut raw = ptr::null_mut();
        let rc = unsafe { sys::b2r_ctx_create(device, &mut raw) };
        if rc != sys::B2R_OK {
            let msg = unsafe { CStr::from_ptr(sys::b2r_last_error(ptr::null())) }.to_string_lossy().into_owned();
            return Err(Error(rc, msg)); // B2R_ERR_NO_DEVICE: the library has no CPU fallback
        }
        Ok(Gpu { raw, bases: RefCell::new(HashMap::new()) })
    }
    pub fn raw(&self) -> *mut sys::b2r_ctx { self.raw }
    pub fn check(&self, rc: i32) -> Result<()> {
        if rc == sys::B2R_OK { return Ok(()); }
        Err(Error(rc, unsafe { CStr::from_ptr(sys::b2r_last_error(self.raw)) }.to_string_lossy().into_owned()))
    }
    /// `ParamsKZG::g` / `g_lagrange` never change: register a slice once, keyed by its address.
    pub fn bases_for(&self, bases: &[G1Affine]) -> Result<*mut sys::b2r_bases> {
        let key = (bases.as_ptr() as usize, bases.len());
        if let Some(b) = self.bases.borrow().get(&key) { return Ok(*b); }
        let mut out = ptr::null_mut();
        self.check(unsafe { sys::b2r_bases_register(self.raw, bases.as_ptr(), bases.len(), &mut out) })?;
        self.bases.borrow_mut().insert(key, out);
        Ok(out)
    }
    /// one context per thread for code that is called from halo2's rayon pool
    pub fn with_thread_local<T>(f: impl FnOnce(&Gpu) -> T) -> T {
        thread_local! { static CTX: Gpu = Gpu::new(std::env::var("B2R_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0)).expect("b2r_ctx_create"); }
        CTX.with(|g| f(g))
    }
}
impl Drop for Gpu {
    fn drop(&mut self) {
        for (_, b) in self.bases.borrow_mut().drain() { unsafe { sys::b2r_bases_free(self.raw, b) }; }
        unsafe { sys::b2r_ctx_destroy(self.raw) };
    }
}

/// Body for `halo2_proofs::arithmetic::best_multiexp::<G1Affine>` (same signature, same panics on misuse).
pub fn best_multiexp(coeffs: &[Fr], bases: &[G1Affine]) -> G1 {
    assert_eq!(coeffs.len(), bases.len());
    Gpu::with_thread_local(|gpu| {
        let reg = gpu.bases_for(bases).expect("b2r_bases_register");
        let mut out = G1::default();
        gpu.check(unsafe { sys::b2r_msm_g1(gpu.raw(), reg, coeffs.as_ptr(), coeffs.len(), &mut out) }).expect("b2r_msm_g1");
        out // normalised (z = 1); identity = (0, 1, 0) as G1::identity()
    })
}

/// Body for `halo2_proofs::arithmetic::best_fft::<Fr>`.
pub fn best_fft(a: &mut [Fr], omega: Fr, log_n: u32) {
    assert_eq!(a.len(), 1 << log_n);
    Gpu::with_thread_local(|gpu| gpu.check(unsafe { sys::b2r_ntt_fr(gpu.raw(), a.as_mut_ptr(), &omega, log_n) }).expect("b2r_ntt_fr"))
}

/// halo2 keeps `VerifyingKey::transcript_repr` private, but `hash_into` hands it to any transcript as the first (and
/// only) `common_scalar`: a transcript that records it recovers the value (oracle/EXT_ASSUMPTIONS.md D2).
pub fn vk_transcript_repr(vk: &VerifyingKey<G1Affine>) -> Fr {
    struct Probe(Option<Fr>);
    #[derive(Clone, Copy, Debug)]
    struct NoChallenge;
    impl EncodedChallenge<G1Affine> for NoChallenge {
        type Input = ();
        fn new(_: &()) -> Self { NoChallenge }
        fn get_scalar(&self) -> Fr { Fr::zero() }
    }
    impl Transcript<G1Affine, NoChallenge> for Probe {
        fn squeeze_challenge(&mut self) -> NoChallenge { NoChallenge }
        fn common_point(&mut self, _: G1Affine) -> io::Result<()> { Ok(()) }
        fn common_scalar(&mut self, s: Fr) -> io::Result<()> { self.0.get_or_insert(s); Ok(()) }
    }
    let mut p = Probe(None);
    vk.hash_into(&mut p).expect("hash_into");
    p.0.expect("vk.hash_into absorbed no scalar")
}
#[allow(dead_code)]
fn _challenge_type_check(_: ChallengeScalar<G1Affine, ()>) {}

/// keygen + batched create_proof of the reference's pkcs1v15 circuit on the GPU.
pub struct RsaProver<'a> {
    gpu: &'a Gpu,
    prog: *mut sys::b2r_prog,
    g: *mut sys::b2r_bases,
    g_lagrange: *mut sys::b2r_bases,
    pk: *mut sys::b2r_pk,
    pub bits_len: u32,
    pub k: u32,
    pub proof_bytes: usize,
    nonce: u64,
}

impl<'a> RsaProver<'a> {
    /// `params`: the reference's `ParamsKZG::<Bn256>::setup(k, OsRng)` (benches/bench.rs:235).  `e`: the fixed public
    /// exponent (`RSAPubE::Fix`, benches/bench.rs:79).
    pub fn new(gpu: &'a Gpu, params: &ParamsKZG<Bn256>, bits_len: u32, e: &num_bigint::BigUint) -> Result<Self> {
        let k = params.k();
        let g: &[G1Affine] = params.get_g();
        // ParamsKZG keeps g_lagrange private in this halo2 version: recompute it the way setup() does
        let g_lagrange: Vec<G1Affine> = g_to_lagrange(g.iter().map(|p| p.to_curve()).collect(), k);
        let (mut prog, mut rg, mut rgl, mut pk) = (ptr::null_mut(), ptr::null_mut(), ptr::null_mut(), ptr::null_mut());
        let e_le = e.to_bytes_le();
        gpu.check(unsafe { sys::b2r_rsa_program_build(gpu.raw(), bits_len, e_le.as_ptr(), e_le.len(), k, &mut prog) })?;
        gpu.check(unsafe { sys::b2r_bases_register(gpu.raw(), g.as_ptr(), g.len(), &mut rg) })?;
        gpu.check(unsafe { sys::b2r_bases_register(gpu.raw(), g_lagrange.as_ptr(), g_lagrange.len(), &mut rgl) })?;
        gpu.check(unsafe { sys::b2r_rsa_keygen(gpu.raw(), prog, rg, rgl, &mut pk) })?;
        let mut pb = 0u64;
        gpu.check(unsafe { sys::b2r_pk_info(pk, ptr::null_mut(), ptr::null_mut(), ptr::null_mut(), ptr::null_mut(), &mut pb) })?;
        Ok(RsaProver { gpu, prog, g: rg, g_lagrange: rgl, pk, bits_len, k, proof_bytes: pb as usize, nonce: 0 })
    }
    /// (fixed commitments, permutation commitments) as the GPU keygen computed them
    pub fn vk_commitments(&self) -> Result<(Vec<G1Affine>, Vec<G1Affine>)> {
        let (mut nf, mut ns) = (0u32, 0u32);
        self.gpu.check(unsafe { sys::b2r_pk_info(self.pk, ptr::null_mut(), ptr::null_mut(), &mut nf, &mut ns, ptr::null_mut()) })?;
        let mut f = vec![G1Affine::default(); nf as usize];
        let mut s = vec![G1Affine::default(); ns as usize];
        let mut t = Fr::zero();
        self.gpu.check(unsafe { sys::b2r_pk_export_vk(self.pk, f.as_mut_ptr(), s.as_mut_ptr(), &mut t) })?;
        Ok((f, s))
    }
    /// MUST be called with the Rust verifying key before proving for a stock verifier (include/b2rsa.h).
    pub fn bind_vk(&mut self, vk: &VerifyingKey<G1Affine>) -> Result<()> {
        let repr = vk_transcript_repr(vk);
        self.gpu.check(unsafe { sys::b2r_pk_set_transcript_repr(self.pk, &repr) })
    }
    /// `create_proof` for every instance: limbs are little-endian u64 words (decompose_big, benches/bench.rs:280-286),
    /// `hashed` = the 4 limbs of the SHA-256 digest as the bench feeds them.  `seed` = 32 bytes of OS entropy.
    /// -> (proofs: batch x proof_bytes, status: 1 valid / 0 invalid signature / 0xFF synthesize would panic / 0xFE)
    pub fn prove_batch(&mut self, n_limbs: &[u64], sig_limbs: &[u64], hashed: &[u64], seed: &[u8; 32]) -> Result<(Vec<u8>, Vec<u8>)> {
        let nl = (self.bits_len / 64) as usize;
        let batch = n_limbs.len() / nl;
        assert!(n_limbs.len() == batch * nl && sig_limbs.len() == batch * nl && hashed.len() == batch * 4);
        let mut proofs = vec![0u8; batch * self.proof_bytes];
        let mut status = vec![0u8; batch];
        self.nonce += 1;
        self.gpu.check(unsafe {
            sys::b2r_rsa_prove_batch_ex(self.gpu.raw(), self.pk, n_limbs.as_ptr(), sig_limbs.as_ptr(), hashed.as_ptr(), batch, seed.as_ptr(),
                                        self.nonce, 0, proofs.as_mut_ptr(), status.as_mut_ptr())
        })?;
        Ok((proofs, status))
    }
    /// `RSASignatureVerifier::verify_pkcs1v15_signature` from the message bytes on (src/lib.rs:183-248): SHA-256 of every
    /// message on the device, then `create_proof`.  -> (proofs, status, digests: batch x 32 bytes = `hashed_bytes`)
    pub fn prove_msgs(&mut self, n_limbs: &[u64], sig_limbs: &[u64], msgs: &[&[u8]], seed: &[u8; 32]) -> Result<(Vec<u8>, Vec<u8>, Vec<u8>)> {
        let nl = (self.bits_len / 64) as usize;
        let batch = msgs.len();
        assert!(n_limbs.len() == batch * nl && sig_limbs.len() == batch * nl);
        let (mut offs, mut buf) = (vec![0u64], Vec::<u8>::new());
        for m in msgs { buf.extend_from_slice(m); offs.push(buf.len() as u64); }
        let mut proofs = vec![0u8; batch * self.proof_bytes];
        let mut status = vec![0u8; batch];
        let mut digests = vec![0u8; batch * 32];
        self.nonce += 1;
        self.gpu.check(unsafe {
            sys::b2r_rsa_prove_msgs_batch(self.gpu.raw(), self.pk, n_limbs.as_ptr(), sig_limbs.as_ptr(), buf.as_ptr(), offs.as_ptr(), batch,
                                          seed.as_ptr(), self.nonce, 0, proofs.as_mut_ptr(), status.as_mut_ptr(), digests.as_mut_ptr())
        })?;
        Ok((proofs, status, digests))
    }
    /// The commitments of the last `prove_*` call (batch x 31 affine points in transcript order), copied from their
    /// device-resident block; a multi-GPU host passes a device pointer and `dst_on_device = 1` instead and hands that
    /// buffer to `ncclAllGather`.
    pub fn last_commitments(&self) -> Result<Vec<G1Affine>> {
        let (mut batch, mut per) = (0usize, 0u32);
        self.gpu.check(unsafe { sys::b2r_last_commitments(self.gpu.raw(), ptr::null_mut(), 0, 0, &mut batch, &mut per) })?;
        let mut out = vec![G1Affine::default(); batch * per as usize];
        self.gpu.check(unsafe { sys::b2r_last_commitments(self.gpu.raw(), out.as_mut_ptr(), out.len(), 0, ptr::null_mut(), ptr::null_mut()) })?;
        Ok(out)
    }
}
impl<'a> Drop for RsaProver<'a> {
    fn drop(&mut self) {
        unsafe {
            sys::b2r_pk_free(self.gpu.raw(), self.pk);
            sys::b2r_bases_free(self.gpu.raw(), self.g);
            sys::b2r_bases_free(self.gpu.raw(), self.g_lagrange);
            sys::b2r_prog_free(self.gpu.raw(), self.prog);
        }
    }
}
