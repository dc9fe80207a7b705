/* libb2rsa - B200-native hot paths for the halo2-rsa prover, C ABI.
 *
 * The reference (SoraSuegami/halo2-rsa @ 86358fa) has no FFI of its own: its two
 * data-parallel hot paths are reached through Rust generics.  Each entry point below
 * names the reference interface it replaces; INTEGRATION.md shows the Rust `extern "C"`
 * binding and the ctypes binding used by this repo's tests.
 *
 * Conventions
 *   - every function returns int32_t: 0 = ok, < 0 = b2r_status error (never aborts,
 *     never unwinds); b2r_last_error(ctx) gives the message of the last failure.
 *   - element memory format = halo2curves bn256 in-memory format:
 *       b2r_fr / b2r_fq : 4 x uint64_t little-endian limbs, Montgomery form (R = 2^256)
 *       b2r_g1_affine   : { fq x; fq y }  64 bytes, identity = (0, 0)
 *       b2r_g1          : { fq x; fq y; fq z }  96 bytes, Jacobian; identity z = 0
 *     so a Rust &[Fr] / &[G1Affine] can be passed as a raw pointer.
 *   - functions without suffix take HOST pointers, copy in/out, and return when the
 *     result is in the caller's buffer (what the reference's synchronous calls do).
 *     `_dev` variants take DEVICE pointers, enqueue on the context's stream and return
 *     immediately (b2r_ctx_sync to wait).  They are what a resident prover pipeline and
 *     bench.py's device-timed leg use.
 *   - a context is bound to one device and one stream; calls on one context must not be
 *     made concurrently from several threads (use one context per thread).
 */
#ifndef B2RSA_H
#define B2RSA_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t l[4]; } b2r_fr;
typedef struct { uint64_t l[4]; } b2r_fq;
typedef struct { b2r_fq x, y; } b2r_g1_affine;
typedef struct { b2r_fq x, y, z; } b2r_g1;

typedef struct b2r_ctx b2r_ctx;
typedef struct b2r_prog b2r_prog;   /* a recorded witness program (static circuit layout) */
typedef struct b2r_bases b2r_bases; /* a resident, pre-processed MSM base set */
typedef struct b2r_pk b2r_pk;       /* proving key of the RSA circuit (fixed / permutation polynomials, cosets, vk) */

enum b2r_status {
    B2R_OK = 0,
    B2R_ERR_INVALID = -1,   /* bad argument (shape, null, range) */
    B2R_ERR_CUDA = -2,      /* CUDA runtime failure (message in b2r_last_error) */
    B2R_ERR_NO_DEVICE = -3, /* no usable sm_100 device: the library has no CPU fallback */
    B2R_ERR_NOMEM = -4,
    B2R_ERR_LAYOUT = -5,    /* circuit does not fit 2^k rows */
    B2R_ERR_SYNTH = -6      /* witness synthesis failed (reference would panic / Err) */
};

/* ---- context ------------------------------------------------------------------------ */
int32_t b2r_ctx_create(int32_t device, b2r_ctx** out);
int32_t b2r_ctx_destroy(b2r_ctx* ctx);
/* use an externally owned cudaStream_t (e.g. torch's current stream); NULL = own stream */
int32_t b2r_ctx_set_stream(b2r_ctx* ctx, void* cuda_stream);
int32_t b2r_ctx_sync(b2r_ctx* ctx);
const char* b2r_last_error(const b2r_ctx* ctx);
const char* b2r_version(void);
/* number of kernel launches this context has enqueued so far (bench.py `gpu_launches`) */
uint64_t b2r_launch_count(const b2r_ctx* ctx);

/* optional per-kernel timing: when enabled, selected kernel launches are bracketed by CUDA
 * events on the context's stream.  b2r_profile_read sums the launches recorded under one
 * kernel name (total device ms, launch count, work units); b2r_profile_dump writes
 * "name:ms:launches;" for all names and optionally clears the records. */
int32_t b2r_profile_enable(b2r_ctx* ctx, int32_t on);
int32_t b2r_profile_read(b2r_ctx* ctx, const char* name, double* total_ms, uint64_t* launches, double* units);
int32_t b2r_profile_dump(b2r_ctx* ctx, char* buf, size_t cap, int32_t clear);

/* Diagnostic: the device field arithmetic on caller-supplied operands, one thread per element (host buffers of n
 * elements in halo2curves' memory format, i.e. Montgomery limbs; operands must be < p).  field: 0 = Fr, 1 = Fq (the
 * arithmetic under halo2curves' bn256::{Fr, Fq}: Mul / Square / Add / Sub / invert).  op: 0 a*b, 1 a^2, 2 a*b + c*d,
 * 3 a*b - c*d, 4 a*b + c*d + a*c + b*d (the sums with ONE reduction the kernels use), 5 a + b, 6 a - b, 7 a^-1 (0 -> 0).
 * c, d may be NULL for ops that do not read them.  Used by tests/test_gpu_field.py against Python integers. */
int32_t b2r_field_selftest(b2r_ctx* ctx, uint32_t field, uint32_t op, const b2r_fr* a, const b2r_fr* b, const b2r_fr* c,
                           const b2r_fr* d, b2r_fr* out, size_t n);

/* plain device memory helpers so a host language needs no CUDA binding of its own */
int32_t b2r_dev_alloc(b2r_ctx* ctx, size_t bytes, void** dptr);
int32_t b2r_dev_free(b2r_ctx* ctx, void* dptr);
int32_t b2r_h2d(b2r_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int32_t b2r_d2h(b2r_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);

/* ---- NTT over Fr -------------------------------------------------------------------
 * Replaces halo2_proofs::arithmetic::best_fft(a: &mut [Fr], omega: Fr, log_n: u32)
 * (third-party crate, reached from reference benches/bench.rs:321-329 via create_proof ->
 * EvaluationDomain::{lagrange_to_coeff, coeff_to_extended, extended_to_coeff}).
 * Natural order in and out, in place, any primitive 2^log_n-th root `omega`. */
int32_t b2r_ntt_fr(b2r_ctx* ctx, b2r_fr* a, const b2r_fr* omega, uint32_t log_n);
int32_t b2r_ntt_fr_dev(b2r_ctx* ctx, b2r_fr* a_dev, const b2r_fr* omega_host, uint32_t log_n);
/* batched: `batch` independent vectors, contiguous, 2^log_n elements each */
int32_t b2r_ntt_fr_batch_dev(b2r_ctx* ctx, b2r_fr* a_dev, size_t batch, const b2r_fr* omega_host,
                             uint32_t log_n);

/* EvaluationDomain::lagrange_to_coeff: iFFT over the 2^k domain (omega_k^-1, then 1/n). */
int32_t b2r_intt_fr(b2r_ctx* ctx, b2r_fr* a, uint32_t k);
int32_t b2r_intt_fr_batch_dev(b2r_ctx* ctx, b2r_fr* a_dev, size_t batch, uint32_t k);
/* EvaluationDomain::coeff_to_extended: 2^k coefficients -> 2^ext_k evaluations on the
 * coset ZETA*<omega_ext> (coefficient i scaled by ZETA^(i mod 3), zero padded, FFT). */
int32_t b2r_coset_ntt_fr(b2r_ctx* ctx, const b2r_fr* coeffs, uint32_t k, uint32_t ext_k, b2r_fr* out);
int32_t b2r_coset_ntt_fr_batch_dev(b2r_ctx* ctx, const b2r_fr* coeffs_dev, size_t batch, uint32_t k,
                                   uint32_t ext_k, b2r_fr* out_dev);
/* EvaluationDomain::extended_to_coeff: inverse of the above on all 2^ext_k values. */
int32_t b2r_coset_intt_fr(b2r_ctx* ctx, b2r_fr* a, uint32_t ext_k);
int32_t b2r_coset_intt_fr_batch_dev(b2r_ctx* ctx, b2r_fr* a_dev, size_t batch, uint32_t ext_k);

/* ---- MSM over BN254 G1 -------------------------------------------------------------
 * Replaces halo2_proofs::arithmetic::best_multiexp(coeffs: &[Fr], bases: &[G1Affine]) -> G1
 * reached through ParamsKZG::{commit, commit_lagrange} (reference benches/bench.rs:235,
 * 321-329).  Bases are made resident once (ParamsKZG::g / g_lagrange never change):
 * registration builds the per-window multiples 2^(c*j) * P_i so that one MSM needs a
 * single bucket set and no window-combining doublings. */
int32_t b2r_bases_register(b2r_ctx* ctx, const b2r_g1_affine* bases_host, size_t n, b2r_bases** out);
int32_t b2r_bases_free(b2r_ctx* ctx, b2r_bases* bases);
/* out = sum_i scalars[i] * bases[i], i < n <= registered size.  The point is returned
 * normalised (z = 1), identity as z = 0: Jacobian coordinates are not canonical, affine
 * ones are, and the reference converts to affine right after (commit -> to_affine). */
int32_t b2r_msm_g1(b2r_ctx* ctx, const b2r_bases* bases, const b2r_fr* scalars, size_t n, b2r_g1* out);
/* `m` scalar vectors of length n sharing one base set -> m affine points
 * (one call per batch of advice/lookup/permutation columns) */
int32_t b2r_msm_g1_batch(b2r_ctx* ctx, const b2r_bases* bases, const b2r_fr* scalars, size_t m,
                         size_t n, b2r_g1_affine* out);
int32_t b2r_msm_g1_batch_dev(b2r_ctx* ctx, const b2r_bases* bases, const b2r_fr* scalars_dev,
                             size_t m, size_t n, b2r_g1_affine* out_dev);
/* The same with a caller's hint.  B2R_MSM_UNIFORM: the non-zero scalars are (pseudo-)random field elements - quotient
 * pieces, opening witnesses, blinded polynomials in coefficient form, i.e. what create_proof hands to ParamsKZG::commit
 * (benches/bench.rs:321-329) - so their digits fill the buckets evenly and the entries are sorted by the two-pass
 * binned counting sort instead of per-entry atomics (4x faster sort).  The hint never changes the result: a vector
 * that turns out not to be uniform overflows a bin, which is detected, and the call falls back to the general path. */
#define B2R_MSM_UNIFORM 1u
int32_t b2r_msm_g1_batch_dev_ex(b2r_ctx* ctx, const b2r_bases* bases, const b2r_fr* scalars_dev,
                                size_t m, size_t n, uint32_t flags, b2r_g1_affine* out_dev);

/* copies the registered affine points (window 0 of the table) back to the host */
int32_t b2r_bases_download(b2r_ctx* ctx, const b2r_bases* bases, b2r_g1_affine* out_host, size_t n);
/* Replaces ParamsKZG::<Bn256>::setup(k, rng) (reference benches/bench.rs:235) for a caller-chosen
 * secret: g[i] = s^i * G, g_lagrange[i] = L_i(s) * G over the 2^k domain; both sets are left
 * resident and MSM-ready.  Either output may be NULL. */
int32_t b2r_srs_setup(b2r_ctx* ctx, uint32_t k, const b2r_fr* secret, b2r_bases** g, b2r_bases** g_lagrange);

/* ---- RSA witness synthesis ---------------------------------------------------------
 * Replaces the witness side of Circuit::synthesize for the reference's pkcs1v15 circuit
 * (benches/bench.rs:132-225, SHA-disabled branch) = RSAChip::verify_pkcs1v15_signature
 * (src/chip.rs:128-199) -> BigIntChip::{assert_in_field, pow_mod_fixed_exp, mul_mod, mul,
 * is_equal_muled, ...} (src/big_integer/chip.rs).  The call sequence is data independent,
 * so it is recorded once into a program and replayed on the GPU for a batch. */
int32_t b2r_rsa_program_build(b2r_ctx* ctx, uint32_t bits_len, const uint8_t* e_le, size_t e_len,
                              uint32_t k, b2r_prog** out);
int32_t b2r_prog_free(b2r_ctx* ctx, b2r_prog* prog);
/* rows used by the layout (must be <= 2^k - blinding rows), number of recorded values */
int32_t b2r_prog_info(const b2r_prog* prog, uint64_t* rows_used, uint64_t* num_values,
                      uint64_t* num_levels);
int32_t b2r_prog_num_limbs(const b2r_prog* prog); /* bits_len / 64, < 0 on error */
/* 64-bit words per instance in the third input array of the witness / commit / prove entry points: 4 (the hash
 * limbs) for b2r_rsa_program_build, 5 (hash limbs, then the exponent) for b2r_rsa_program_build_var, num_limbs + 1
 * (n, then the exponent word) for b2r_bigint_program_build */
int32_t b2r_prog_aux_words(const b2r_prog* prog);
/* The pkcs1v15 circuit with RSAPubE::Var (reference src/lib.rs:58-63, src/chip.rs:58-70 and :99-114): the exponent is
 * an assigned one-limb integer and RSAChip::modpow_public_key runs BigIntChip::pow_mod over exp_limb_bits of its bits
 * (src/big_integer/chip.rs:664-696: one mul_mod, num_limbs selects and one square_mod per bit).  Input arrays as for
 * b2r_rsa_program_build with the third one holding hash[0..3], e per instance. */
int32_t b2r_rsa_program_build_var(b2r_ctx* ctx, uint32_t bits_len, uint32_t exp_limb_bits, uint32_t k, b2r_prog** out);
/* RSASignatureVerifier::verify_pkcs1v15_signature (reference src/lib.rs:183-248) from the digest bytes on - the part of
 * the SHA-256 front end that the reference itself holds: the 32 digest-byte cells (which halo2-dynamic-sha256's chip, an
 * unpinned external crate, would provide; stood in for by one assign_value row each) are composed into four 64-bit limbs
 * by assign_constant / mul_add (src/lib.rs:222-236) and verified in the same region; is_valid is returned, not asserted
 * (src/lib.rs:245).  Inputs as for b2r_rsa_program_build (the third array holds the digest as four little-endian
 * 64-bit limbs, i.e. the 32 bytes least significant first). */
int32_t b2r_rsa_program_build_sha_tail(b2r_ctx* ctx, uint32_t bits_len, const uint8_t* e_le, size_t e_len, uint32_t k,
                                       b2r_prog** out);
/* One BigIntInstructions method as the reference's in-file unit-test circuits drive it (src/big_integer/chip.rs:
 * 1861-1899 refresh, 1948-1986 add_mod, 2027-2070 sub_mod, 2229-2271 pow_mod), for the methods the pkcs1v15 circuit
 * does not call.  op: 6 = mul + refresh (both operand orders, assert_equal_fresh), 7 = add_mod, 8 = sub_mod,
 * 9 = pow_mod with an assigned one-limb exponent; the predicates of src/big_integer/chip.rs:754-1006 on (a, b), whose
 * one result cell is also what is_valid reports: 10 = is_zero(a), 11 = is_equal_fresh, 12 = is_less_than,
 * 13 = is_less_than_or_equal, 14 = is_greater_than, 15 = is_greater_than_or_equal, 16 = is_in_field(a, b);
 * 17 = square(a) (:431-437), 18 = square_mod(a, n) (:642-649).  Witness inputs: first array a, second array b, third array n
 * followed by the exponent word (b2r_prog_aux_words = num_limbs + 1).  is_valid of the witness entry points is 0xFF
 * where the reference would have panicked and otherwise not meaningful for these programs (it tests the last result
 * limb against 1): compare the advice columns. */
int32_t b2r_bigint_program_build(b2r_ctx* ctx, uint32_t op, uint32_t bits_len, uint32_t exp_limb_bits, uint32_t k, b2r_prog** out);
/* n_limbs, sig_limbs: batch x (bits_len/64) little-endian 64-bit limbs; hash_limbs:
 * batch x 4.  advice: batch x 5 x 2^k Fr, column-major per instance (HOST pointer);
 * is_valid: batch bytes (the value of the circuit's final is_valid cell).
 * blind_seed != 0 fills the last 6 rows of every column with the seeded blinding stream (the rows
 * halo2's create_proof fills from its RNG; nonce 0, i.e. reproducible: these entry points expose the
 * advice table, the complete prover is b2r_rsa_prove_batch*); 0 leaves them zero. */
int32_t b2r_rsa_witness_batch(b2r_ctx* ctx, const b2r_prog* prog, const uint64_t* n_limbs,
                              const uint64_t* sig_limbs, const uint64_t* hash_limbs, size_t batch,
                              uint64_t blind_seed, b2r_fr* advice, uint8_t* is_valid);
int32_t b2r_rsa_witness_batch_dev(b2r_ctx* ctx, const b2r_prog* prog, const uint64_t* n_limbs_dev,
                                  const uint64_t* sig_limbs_dev, const uint64_t* hash_limbs_dev,
                                  size_t batch, uint64_t blind_seed, b2r_fr* advice_dev,
                                  uint8_t* is_valid_dev);


/* ---- fused hot path ---------------------------------------------------------------------
 * What create_proof does with the advice columns of `batch` independent RSA instances
 * (reference benches/bench.rs:321-329; SURVEY.md 3 Stack 1 steps 2 and 6), in one call with
 * everything resident between the stages:
 *   witness synthesis -> commit_lagrange(5 advice columns) -> lagrange_to_coeff -> coeff_to_extended.
 * advice_dev: device buffer batch x 5 x 2^k Fr (holds coefficients on return when ext_dev is
 * given, Lagrange values otherwise); ext_dev: device buffer batch x 5 x 2^ext_k Fr or NULL to
 * stop after the commitments; commitments: batch x 5 affine points; is_valid: batch bytes
 * (1 valid, 0 invalid, 0xFF = the reference's synthesize would have panicked).
 * The non-_dev variant takes HOST inputs and returns HOST commitments / flags. */
int32_t b2r_rsa_commit_batch(b2r_ctx* ctx, const b2r_prog* prog, const b2r_bases* g_lagrange,
                             const uint64_t* n_limbs, const uint64_t* sig_limbs, const uint64_t* hash_limbs,
                             size_t batch, uint64_t blind_seed, uint32_t k, uint32_t ext_k,
                             b2r_fr* advice_dev, b2r_fr* ext_dev, b2r_g1_affine* commitments, uint8_t* is_valid);
int32_t b2r_rsa_commit_batch_dev(b2r_ctx* ctx, const b2r_prog* prog, const b2r_bases* g_lagrange,
                                 const uint64_t* n_limbs_dev, const uint64_t* sig_limbs_dev,
                                 const uint64_t* hash_limbs_dev, size_t batch, uint64_t blind_seed,
                                 uint32_t k, uint32_t ext_k, b2r_fr* advice_dev, b2r_fr* ext_dev,
                                 b2r_g1_affine* commitments_dev, uint8_t* is_valid_dev);

/* ---- keygen and the full prover ---------------------------------------
 * b2r_rsa_keygen replaces keygen_vk + keygen_pk for the pkcs1v15 circuit (reference benches/bench.rs:236-237):
 * fixed columns, lookup table, permutation (sigma) polynomials, their coefficient / extended-coset forms and
 * the verifying-key commitments, all resident on the GPU.  `g` / `g_lagrange` are the two SRS base sets
 * (b2r_srs_setup or b2r_bases_register) and must outlive the key, as must `prog`. */
int32_t b2r_rsa_keygen(b2r_ctx* ctx, const b2r_prog* prog, const b2r_bases* g, const b2r_bases* g_lagrange, b2r_pk** out);
int32_t b2r_pk_free(b2r_ctx* ctx, b2r_pk* pk);
int32_t b2r_pk_info(const b2r_pk* pk, uint32_t* k, uint32_t* ext_k, uint32_t* num_fixed, uint32_t* num_sigma, uint64_t* proof_bytes);
/* verifying key: num_fixed fixed-column commitments, num_sigma permutation commitments, vk transcript scalar */
int32_t b2r_pk_export_vk(const b2r_pk* pk, b2r_g1_affine* fixed_commitments, b2r_g1_affine* sigma_commitments, b2r_fr* transcript_repr);
/* halo2's VerifyingKey::transcript_repr is Blake2b("Halo2-Verify-Key") over the Rust Debug string of vk.pinned(),
 * which only the Rust crate can produce; b2r_rsa_keygen installs a stand-in (a hash of k and the vk commitments) that
 * only this repo's oracle verifier shares.  A Rust caller MUST pass the real value (vk.transcript_repr, private in
 * halo2_proofs: the shim in INTEGRATION.md section 3 obtains it by absorbing the vk into a probe transcript) before
 * proving, or stock verify_proof derives different challenges and rejects every proof.  Reduced Montgomery limbs. */
int32_t b2r_pk_set_transcript_repr(b2r_pk* pk, const b2r_fr* transcript_repr);
/* Replaces create_proof::<KZGCommitmentScheme<Bn256>, ProverGWC<_>, Challenge255<_>, _, Blake2bWrite<..>, _>
 * (reference benches/bench.rs:319-331) for `batch` independent instances: HOST inputs as in
 * b2r_rsa_witness_batch, HOST outputs: proofs = batch x proof_bytes (b2r_pk_info), status = batch bytes
 * (1 = proof of a valid signature; 0 = the witness does not satisfy the circuit, the proof will not verify;
 * 0xFF = the reference's synthesize would have panicked; 0xFE = a range-checked cell is out of range).
 * `seed` (non-zero) keys the blinding stream that stands in for the reference's OsRng: ChaCha20 blocks reduced mod r
 * (DESIGN.md section 4a).  These two entry points expand the 64-bit seed into the 256-bit ChaCha key (64 bits of
 * entropy: tests and benchmarks) and take a fresh nonce from the context on every call, so the same seed never yields
 * the same blinds twice; b2r_rsa_prove_batch_ex takes the full key and an explicit nonce. */
int32_t b2r_rsa_prove_batch(b2r_ctx* ctx, const b2r_pk* pk, const uint64_t* n_limbs, const uint64_t* sig_limbs,
                            const uint64_t* hash_limbs, size_t batch, uint64_t seed, uint8_t* proofs, uint8_t* status);
/* same with the instance inputs already resident in HBM (DEVICE pointers); proofs / status are HOST buffers: the proof
 * bytes are assembled by the per-proof transcripts on the host. */
int32_t b2r_rsa_prove_batch_dev(b2r_ctx* ctx, const b2r_pk* pk, const uint64_t* n_limbs_dev, const uint64_t* sig_limbs_dev,
                                const uint64_t* hash_limbs_dev, size_t batch, uint64_t seed, uint8_t* proofs, uint8_t* status);

/* The same with caller-owned randomness: seed32 = 32 bytes from the host's CSPRNG (the ChaCha20 key; what OsRng is to
 * the reference), nonce = any value the caller never repeats under one key (a proof-batch counter).  (key, nonce,
 * instance index) determine every blinding value, so a call is reproducible - the parity tests rely on that - and a
 * caller who repeats a (key, nonce) pair for different witnesses gives up zero knowledge for them.
 * flags: B2R_PROVE_INPUTS_ON_DEVICE = the three input arrays are device pointers; B2R_PROVE_SEED64 = only the first 8
 * bytes of seed32 are used, expanded as the 64-bit-seed entry points do (test vectors). */
#define B2R_PROVE_INPUTS_ON_DEVICE 1u
#define B2R_PROVE_SEED64 2u
int32_t b2r_rsa_prove_batch_ex(b2r_ctx* ctx, const b2r_pk* pk, const uint64_t* n_limbs, const uint64_t* sig_limbs,
                               const uint64_t* hash_limbs, size_t batch, const uint8_t seed32[32], uint64_t nonce,
                               uint32_t flags, uint8_t* proofs, uint8_t* status);

/* The commitments of the last completed b2r_rsa_prove_batch* / b2r_rsa_prove_msgs_batch call on this context, DEVICE
 * RESIDENT: batch x per_proof (= 31) affine points (Montgomery coordinates, identity (0, 0)) in the order the transcript
 * absorbs them and the proof stream carries them compressed - 5 advice, 10 permuted lookup columns (A'_0, S'_0, A'_1, ..),
 * 2 permutation + 5 lookup grand products, the random polynomial, 4 quotient pieces, 4 GWC witnesses.  This is the block
 * a multi-GPU host all-gathers over NCCL without a host round trip (reference benches/bench.rs:319-331
 * produces one proof per instance, the commitments are its first 27 and last 4 group elements).
 * dst: capacity_points affine points, a DEVICE pointer if dst_on_device != 0 (device-to-device copy on the context's
 * stream, not synchronised) else a HOST pointer; dst == NULL only queries batch / per_proof.  The block is overwritten by
 * the next prove call. */
int32_t b2r_last_commitments(b2r_ctx* ctx, b2r_g1_affine* dst, size_t capacity_points, uint32_t dst_on_device, size_t* batch,
                             uint32_t* per_proof);

/* ---- SHA-256 front end -----------------------------------------------------------
 * RSASignatureVerifier::verify_pkcs1v15_signature (reference src/lib.rs:183-248) starts from the signed MESSAGE:
 * step 1 hashes it (`sha256.finalize`, `decompose_digest_to_bytes`, `hashed_bytes.reverse()`, :204-211), step 2
 * composes the digest bytes into four limbs and verifies (:218-245, built as b2r_rsa_program_build_sha_tail).
 * These entry points are step 1 at the value level, on the GPU: SHA-256 (FIPS 180-4; the reference's bench and tests
 * take the same digest from the `sha2` crate, benches/bench.rs:255-268) of `batch` messages, message i being
 * msgs[offsets[i] .. offsets[i + 1]) (offsets: batch + 1 non-decreasing entries; empty messages allowed).
 * hash_limbs: batch x 4 words, the digest read as a big-endian integer in little-endian 64-bit limbs - the third
 * input array of b2r_rsa_witness_batch / b2r_rsa_prove_batch for the pkcs1v15 and sha_tail programs; digests:
 * batch x 32 bytes in SHA order (the `hashed_bytes` the reference returns, :246-247).  Either output may be NULL.
 * The constraint layout of the compression function belongs to halo2-dynamic-sha256 (unpinned external crate, not
 * vendored) and is not reproduced: see DESIGN.md section 7. */
int32_t b2r_sha256_batch(b2r_ctx* ctx, const uint8_t* msgs, const uint64_t* offsets, size_t batch, uint64_t* hash_limbs,
                         uint8_t* digests);
int32_t b2r_sha256_batch_dev(b2r_ctx* ctx, const uint8_t* msgs_dev, const uint64_t* offsets_dev, size_t batch,
                             uint64_t* hash_limbs_dev, uint8_t* digests_dev);
/* message bytes -> proofs in one call for a key over a program that takes a 4-limb digest (b2r_rsa_program_build or
 * b2r_rsa_program_build_sha_tail): HOST inputs, the messages are hashed on the device and never leave it; seed32 /
 * nonce / flags as in b2r_rsa_prove_batch_ex (B2R_PROVE_SEED64 only); digests (batch x 32 bytes, may be NULL) returns
 * the `hashed_bytes` of every instance. */
int32_t b2r_rsa_prove_msgs_batch(b2r_ctx* ctx, const b2r_pk* pk, const uint64_t* n_limbs, const uint64_t* sig_limbs,
                                 const uint8_t* msgs, const uint64_t* msg_offsets, size_t batch, const uint8_t seed32[32],
                                 uint64_t nonce, uint32_t flags, uint8_t* proofs, uint8_t* status, uint8_t* digests);

#ifdef __cplusplus
}
#endif
#endif /* B2RSA_H */
