//! `extern "C"` declarations of include/b2rsa.h, one for one.  Element types are halo2curves' own: `Fr`/`Fq` are
//! `[u64; 4]` Montgomery limbs and `G1Affine` is `{x, y}` (identity (0, 0)) in memory, which is the ABI's format, so
//! slices cross the boundary as raw pointers without conversion (oracle/EXT_ASSUMPTIONS.md A1, A2).
#![allow(non_camel_case_types)]
use halo2wrong::curves::bn256::{Fr, G1Affine, G1};
use std::os::raw::{c_char, c_void};

#[repr(C)] pub struct b2r_ctx { _p: [u8; 0] }
#[repr(C)] pub struct b2r_prog { _p: [u8; 0] }
#[repr(C)] pub struct b2r_bases { _p: [u8; 0] }
#[repr(C)] pub struct b2r_pk { _p: [u8; 0] }

pub const B2R_OK: i32 = 0;
pub const B2R_ERR_INVALID: i32 = -1;
pub const B2R_ERR_CUDA: i32 = -2;
pub const B2R_ERR_NO_DEVICE: i32 = -3;
pub const B2R_ERR_NOMEM: i32 = -4;
pub const B2R_ERR_LAYOUT: i32 = -5;
pub const B2R_ERR_SYNTH: i32 = -6;
pub const B2R_MSM_UNIFORM: u32 = 1;
pub const B2R_PROVE_INPUTS_ON_DEVICE: u32 = 1;
pub const B2R_PROVE_SEED64: u32 = 2;

#[link(name = "b2rsa")]
extern "C" {
    // ---- context
    pub fn b2r_ctx_create(device: i32, out: *mut *mut b2r_ctx) -> i32;
    pub fn b2r_ctx_destroy(ctx: *mut b2r_ctx) -> i32;
    pub fn b2r_ctx_set_stream(ctx: *mut b2r_ctx, cuda_stream: *mut c_void) -> i32;
    pub fn b2r_ctx_sync(ctx: *mut b2r_ctx) -> i32;
    pub fn b2r_last_error(ctx: *const b2r_ctx) -> *const c_char;
    pub fn b2r_version() -> *const c_char;
    pub fn b2r_launch_count(ctx: *const b2r_ctx) -> u64;
    pub fn b2r_dev_alloc(ctx: *mut b2r_ctx, bytes: usize, dptr: *mut *mut c_void) -> i32;
    pub fn b2r_dev_free(ctx: *mut b2r_ctx, dptr: *mut c_void) -> i32;
    pub fn b2r_h2d(ctx: *mut b2r_ctx, dst_dev: *mut c_void, src_host: *const c_void, bytes: usize) -> i32;
    pub fn b2r_d2h(ctx: *mut b2r_ctx, dst_host: *mut c_void, src_dev: *const c_void, bytes: usize) -> i32;
    // ---- halo2_proofs::arithmetic::best_fft and the EvaluationDomain wrappers
    pub fn b2r_ntt_fr(ctx: *mut b2r_ctx, a: *mut Fr, omega: *const Fr, log_n: u32) -> i32;
    pub fn b2r_ntt_fr_batch_dev(ctx: *mut b2r_ctx, a_dev: *mut Fr, batch: usize, omega_host: *const Fr, log_n: u32) -> i32;
    pub fn b2r_intt_fr(ctx: *mut b2r_ctx, a: *mut Fr, k: u32) -> i32;
    pub fn b2r_coset_ntt_fr(ctx: *mut b2r_ctx, coeffs: *const Fr, k: u32, ext_k: u32, out: *mut Fr) -> i32;
    pub fn b2r_coset_intt_fr(ctx: *mut b2r_ctx, a: *mut Fr, ext_k: u32) -> i32;
    // ---- halo2_proofs::arithmetic::best_multiexp over resident bases (ParamsKZG::g / g_lagrange)
    pub fn b2r_bases_register(ctx: *mut b2r_ctx, bases_host: *const G1Affine, n: usize, out: *mut *mut b2r_bases) -> i32;
    pub fn b2r_bases_free(ctx: *mut b2r_ctx, bases: *mut b2r_bases) -> i32;
    pub fn b2r_bases_download(ctx: *mut b2r_ctx, bases: *const b2r_bases, out_host: *mut G1Affine, n: usize) -> i32;
    pub fn b2r_msm_g1(ctx: *mut b2r_ctx, bases: *const b2r_bases, scalars: *const Fr, n: usize, out: *mut G1) -> i32;
    pub fn b2r_msm_g1_batch(ctx: *mut b2r_ctx, bases: *const b2r_bases, scalars: *const Fr, m: usize, n: usize, out: *mut G1Affine) -> i32;
    pub fn b2r_msm_g1_batch_dev_ex(ctx: *mut b2r_ctx, bases: *const b2r_bases, scalars_dev: *const Fr, m: usize, n: usize, flags: u32,
                                   out_dev: *mut G1Affine) -> i32;
    pub fn b2r_srs_setup(ctx: *mut b2r_ctx, k: u32, secret: *const Fr, g: *mut *mut b2r_bases, g_lagrange: *mut *mut b2r_bases) -> i32;
    // ---- Circuit::synthesize of the pkcs1v15 circuit (benches/bench.rs:132-225)
    pub fn b2r_rsa_program_build(ctx: *mut b2r_ctx, bits_len: u32, e_le: *const u8, e_len: usize, k: u32, out: *mut *mut b2r_prog) -> i32;
    pub fn b2r_rsa_program_build_sha_tail(ctx: *mut b2r_ctx, bits_len: u32, e_le: *const u8, e_len: usize, k: u32, out: *mut *mut b2r_prog) -> i32;
    pub fn b2r_rsa_program_build_var(ctx: *mut b2r_ctx, bits_len: u32, exp_limb_bits: u32, k: u32, out: *mut *mut b2r_prog) -> i32;
    pub fn b2r_prog_free(ctx: *mut b2r_ctx, prog: *mut b2r_prog) -> i32;
    pub fn b2r_prog_info(prog: *const b2r_prog, rows_used: *mut u64, num_values: *mut u64, num_levels: *mut u64) -> i32;
    pub fn b2r_prog_num_limbs(prog: *const b2r_prog) -> i32;
    pub fn b2r_prog_aux_words(prog: *const b2r_prog) -> i32;
    pub fn b2r_rsa_witness_batch(ctx: *mut b2r_ctx, prog: *const b2r_prog, n_limbs: *const u64, sig_limbs: *const u64, hash_limbs: *const u64,
                                 batch: usize, blind_seed: u64, advice: *mut Fr, is_valid: *mut u8) -> i32;
    // ---- keygen_vk / keygen_pk / create_proof for the batch (benches/bench.rs:236-237, 319-331)
    pub fn b2r_rsa_keygen(ctx: *mut b2r_ctx, prog: *const b2r_prog, g: *const b2r_bases, g_lagrange: *const b2r_bases, out: *mut *mut b2r_pk) -> i32;
    pub fn b2r_pk_free(ctx: *mut b2r_ctx, pk: *mut b2r_pk) -> i32;
    pub fn b2r_pk_info(pk: *const b2r_pk, k: *mut u32, ext_k: *mut u32, num_fixed: *mut u32, num_sigma: *mut u32, proof_bytes: *mut u64) -> i32;
    pub fn b2r_pk_export_vk(pk: *const b2r_pk, fixed: *mut G1Affine, sigma: *mut G1Affine, transcript_repr: *mut Fr) -> i32;
    pub fn b2r_pk_set_transcript_repr(pk: *mut b2r_pk, transcript_repr: *const Fr) -> i32;
    // diagnostic: device field arithmetic on host operands (field 0 = Fr, 1 = Fq; see include/b2rsa.h for the op codes)
    pub fn b2r_field_selftest(ctx: *mut b2r_ctx, field: u32, op: u32, a: *const Fr, b: *const Fr, c: *const Fr, d: *const Fr, out: *mut Fr, n: usize) -> i32;
    pub fn b2r_rsa_prove_batch(ctx: *mut b2r_ctx, pk: *const b2r_pk, n_limbs: *const u64, sig_limbs: *const u64, hash_limbs: *const u64,
                               batch: usize, seed: u64, proofs: *mut u8, status: *mut u8) -> i32;
    pub fn b2r_rsa_prove_batch_ex(ctx: *mut b2r_ctx, pk: *const b2r_pk, n_limbs: *const u64, sig_limbs: *const u64, hash_limbs: *const u64,
                                  batch: usize, seed32: *const u8, nonce: u64, flags: u32, proofs: *mut u8, status: *mut u8) -> i32;
    // device-resident commitments of the last prove call: batch x 31 affine points in transcript order (the all-gather payload)
    pub fn b2r_last_commitments(ctx: *mut b2r_ctx, dst: *mut G1Affine, capacity_points: usize, dst_on_device: u32, batch: *mut usize,
                                per_proof: *mut u32) -> i32;
    // ---- RSASignatureVerifier::verify_pkcs1v15_signature from the message bytes on (src/lib.rs:183-248): SHA-256 on the device
    pub fn b2r_sha256_batch(ctx: *mut b2r_ctx, msgs: *const u8, offsets: *const u64, batch: usize, hash_limbs: *mut u64, digests: *mut u8) -> i32;
    pub fn b2r_sha256_batch_dev(ctx: *mut b2r_ctx, msgs_dev: *const u8, offsets_dev: *const u64, batch: usize, hash_limbs_dev: *mut u64,
                                digests_dev: *mut u8) -> i32;
    pub fn b2r_rsa_prove_msgs_batch(ctx: *mut b2r_ctx, pk: *const b2r_pk, n_limbs: *const u64, sig_limbs: *const u64, msgs: *const u8,
                                    msg_offsets: *const u64, batch: usize, seed32: *const u8, nonce: u64, flags: u32, proofs: *mut u8,
                                    status: *mut u8, digests: *mut u8) -> i32;
}
