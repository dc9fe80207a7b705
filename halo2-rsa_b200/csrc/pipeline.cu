// KZG SRS setup on the GPU and the fused hot-path entry point.
//
//  b2r_srs_setup       restates halo2_proofs ParamsKZG::<Bn256>::setup(k, rng) for a caller-chosen
//                      secret s (reference call site benches/bench.rs:235): g[i] = s^i * G and
//                      g_lagrange[i] = L_i(s) * G, L_i(s) = omega^i (s^n - 1) / (n (s - omega^i)).
//                      Both base sets are left resident and pre-processed for the MSM.
//  b2r_rsa_commit_batch  the prover's hot path for a batch of RSA instances in one call
//                      (create_proof steps 2 and 6 for the advice columns, SURVEY.md 3 Stack 1):
//                      witness synthesis -> commit_lagrange of the 5 advice columns ->
//                      lagrange_to_coeff -> coeff_to_extended, everything resident in HBM between
//                      the stages; only the inputs go in and the commitments come out.
#include <cuda_runtime.h>

#include <vector>

#include "ctx.hpp"
#include "ec.cuh"

namespace b2r {
fe_t fr_omega(uint32_t k);
fe_t fr_from_u64(uint64_t v);
int32_t msm_batch_dev(b2r_ctx* ctx, const b2r_bases* bs, const fe_t* scalars_dev, size_t m, size_t n, affine_t* out_dev, bool uniform);

// scalar * G by double-and-add over the canonical bits (one thread per point)
__device__ affine_t g1_mul_generator(const fe_t& scalar_mont) {
    fe_t k = Fr::from_mont(scalar_mont);
    affine_t g;
    g.x = Fq::one();
    g.y = Fq::dbl(Fq::one());
    xyzz_t acc = xyzz_identity();
    bool started = false;
    for (int w = 7; w >= 0; w--) {
        for (int b = 31; b >= 0; b--) {
            if (started) acc = xyzz_double(acc);
            if ((k.l[w] >> b) & 1u) {
                xyzz_madd(acc, g, false);
                started = true;
            }
        }
    }
    return xyzz_to_affine(acc);
}

// mode 0: scalar_i = s^i ; mode 1: scalar_i = L_i(s)
__global__ void __launch_bounds__(128) k_srs_points(affine_t* out, fe_t s, fe_t omega, fe_t zn_over_n, uint32_t n, int mode) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t e[8] = {i, 0, 0, 0, 0, 0, 0, 0};
    fe_t sc;
    if (mode == 0) {
        sc = Fr::pow(s, e);
    } else {
        fe_t wi = Fr::pow(omega, e);
        fe_t d = Fr::inv(Fr::sub(s, wi));
        sc = Fr::mul(Fr::mul(wi, zn_over_n), d);
    }
    out[i] = g1_mul_generator(sc);
}

// b2r_field_selftest: one field operation per thread on caller-supplied operands
template <class F>
__global__ void __launch_bounds__(128) k_field_selftest(uint32_t op, const fe_t* a, const fe_t* b, const fe_t* c, const fe_t* d, fe_t* out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fe_t x = a[i], y = b ? b[i] : F::zero(), z = c ? c[i] : F::zero(), w = d ? d[i] : F::zero();
    fe_t r = F::zero();
    switch (op) {
        case 0: r = F::mul(x, y); break;
        case 1: r = F::sqr(x); break;
        case 2: r = F::mul_add_mul(x, y, z, w); break;
        case 3: r = F::mul_sub_mul(x, y, z, w); break;
        case 4: r = F::dot4(x, y, z, w, x, z, y, w); break;
        case 5: r = F::add(x, y); break;
        case 6: r = F::sub(x, y); break;
        case 7: r = F::inv_vartime(x); break;
    }
    out[i] = r;
}
}  // namespace b2r

using namespace b2r;

extern "C" {

int32_t b2r_srs_setup(b2r_ctx* ctx, uint32_t k, const b2r_fr* secret, b2r_bases** g, b2r_bases** g_lagrange) try {
    B2R_ENTER(ctx);
    if (!secret || (!g && !g_lagrange)) return fail(ctx, B2R_ERR_INVALID, "srs_setup: null argument");
    if (k > 24) return fail(ctx, B2R_ERR_INVALID, "srs_setup: k > 24");
    const uint32_t n = 1u << k;
    fe_t s;
    for (int i = 0; i < 4; i++) {
        s.l[2 * i] = (uint32_t)secret->l[i];
        s.l[2 * i + 1] = (uint32_t)(secret->l[i] >> 32);
    }
    fe_t omega = fr_omega(k);
    uint32_t ne[8] = {n, 0, 0, 0, 0, 0, 0, 0};
    fe_t zn = Fr::sub(Fr::pow(s, ne), Fr::one());
    fe_t zn_over_n = Fr::mul(zn, Fr::inv(fr_from_u64(n)));
    if (Fr::is_zero(zn)) return fail(ctx, B2R_ERR_INVALID, "srs_setup: secret lies in the evaluation domain");
    affine_t* d = nullptr;
    B2R_TRY(scratch_get(ctx, SC_MISC, (size_t)n * sizeof(affine_t), (void**)&d));
    std::vector<affine_t> host(n);
    for (int mode = 0; mode < 2; mode++) {
        b2r_bases** dst = mode == 0 ? g : g_lagrange;
        if (!dst) continue;
        k_srs_points<<<(n + 127) / 128, 128, 0, ctx->stream>>>(d, s, omega, zn_over_n, n, mode);
        B2R_LAUNCH_CHECK(ctx);
        B2R_CUDA(ctx, cudaMemcpyAsync(host.data(), d, (size_t)n * sizeof(affine_t), cudaMemcpyDeviceToHost, ctx->stream));
        B2R_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        B2R_TRY(b2r_bases_register(ctx, (const b2r_g1_affine*)host.data(), n, dst));
    }
    return 0;
} B2R_ABI_CATCH(ctx)

int32_t b2r_rsa_commit_batch_dev(b2r_ctx* ctx, const b2r_prog* prog, const b2r_bases* g_lagrange, const uint64_t* n_limbs_dev,
                                 const uint64_t* sig_limbs_dev, const uint64_t* hash_limbs_dev, size_t batch, uint64_t blind_seed,
                                 uint32_t k, uint32_t ext_k, b2r_fr* advice_dev, b2r_fr* ext_dev, b2r_g1_affine* commitments_dev,
                                 uint8_t* is_valid_dev) try {
    B2R_ENTER(ctx);
    if (!prog || !g_lagrange || !advice_dev || !commitments_dev || !is_valid_dev) return fail(ctx, B2R_ERR_INVALID, "rsa_commit: null pointer");
    const size_t n = (size_t)1 << k;
    // (a) witness: 5 advice columns per instance, Lagrange basis, Montgomery form
    B2R_TRY(b2r_rsa_witness_batch_dev(ctx, prog, n_limbs_dev, sig_limbs_dev, hash_limbs_dev, batch, blind_seed, advice_dev, is_valid_dev));
    // (b) commit_lagrange of every column: one batched MSM over the resident g_lagrange table
    B2R_TRY(msm_batch_dev(ctx, g_lagrange, (const fe_t*)advice_dev, batch * 5, n, (affine_t*)commitments_dev, false));
    if (ext_dev) {
        // lagrange_to_coeff in place, then coeff_to_extended into the extended-domain buffer
        B2R_TRY(b2r_intt_fr_batch_dev(ctx, advice_dev, batch * 5, k));
        B2R_TRY(b2r_coset_ntt_fr_batch_dev(ctx, advice_dev, batch * 5, k, ext_k, ext_dev));
    }
    return 0;
} B2R_ABI_CATCH(ctx)

int32_t b2r_rsa_commit_batch(b2r_ctx* ctx, const b2r_prog* prog, const b2r_bases* g_lagrange, const uint64_t* n_limbs,
                             const uint64_t* sig_limbs, const uint64_t* hash_limbs, size_t batch, uint64_t blind_seed, uint32_t k,
                             uint32_t ext_k, b2r_fr* advice_dev, b2r_fr* ext_dev, b2r_g1_affine* commitments, uint8_t* is_valid) try {
    B2R_ENTER(ctx);
    if (!prog || !g_lagrange || !n_limbs || !sig_limbs || !hash_limbs || !advice_dev || !commitments || !is_valid)
        return fail(ctx, B2R_ERR_INVALID, "rsa_commit: null pointer");
    if (batch == 0) return 0;
    const size_t limbs = (size_t)b2r_prog_num_limbs(prog);
    const size_t aux = (size_t)b2r_prog_aux_words(prog);
    const size_t in_words = batch * (2 * limbs + aux);
    char* d = nullptr;
    size_t in_al = (in_words * 8 + 255) & ~(size_t)255;
    size_t cm_al = (batch * 5 * sizeof(b2r_g1_affine) + 255) & ~(size_t)255;
    B2R_TRY(scratch_get(ctx, SC_MISC, in_al + cm_al + batch + 256, (void**)&d));
    uint64_t* d_n = (uint64_t*)d;
    uint64_t* d_s = d_n + batch * limbs;
    uint64_t* d_h = d_s + batch * limbs;
    b2r_g1_affine* d_cm = (b2r_g1_affine*)(d + in_al);
    uint8_t* d_valid = (uint8_t*)(d + in_al + cm_al);
    B2R_CUDA(ctx, cudaMemcpyAsync(d_n, n_limbs, batch * limbs * 8, cudaMemcpyHostToDevice, ctx->stream));
    B2R_CUDA(ctx, cudaMemcpyAsync(d_s, sig_limbs, batch * limbs * 8, cudaMemcpyHostToDevice, ctx->stream));
    B2R_CUDA(ctx, cudaMemcpyAsync(d_h, hash_limbs, batch * aux * 8, cudaMemcpyHostToDevice, ctx->stream));
    B2R_TRY(b2r_rsa_commit_batch_dev(ctx, prog, g_lagrange, d_n, d_s, d_h, batch, blind_seed, k, ext_k, advice_dev, ext_dev, d_cm, d_valid));
    B2R_CUDA(ctx, cudaMemcpyAsync(commitments, d_cm, batch * 5 * sizeof(b2r_g1_affine), cudaMemcpyDeviceToHost, ctx->stream));
    B2R_CUDA(ctx, cudaMemcpyAsync(is_valid, d_valid, batch, cudaMemcpyDeviceToHost, ctx->stream));
    B2R_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
} B2R_ABI_CATCH(ctx)


int32_t b2r_field_selftest(b2r_ctx* ctx, uint32_t field, uint32_t op, const b2r_fr* a, const b2r_fr* b, const b2r_fr* c,
                           const b2r_fr* d, b2r_fr* out, size_t n) try {
    B2R_ENTER(ctx);
    if (!a || !out || field > 1 || op > 7) return fail(ctx, B2R_ERR_INVALID, "field_selftest: bad argument");
    const bool need_b = op == 0 || (op >= 2 && op <= 6), need_cd = op >= 2 && op <= 4;
    if ((need_b && !b) || (need_cd && (!c || !d))) return fail(ctx, B2R_ERR_INVALID, "field_selftest: operand missing");
    if (n == 0) return 0;
    fe_t* buf = nullptr;
    B2R_TRY(scratch_get(ctx, SC_STAGE, 5 * n * sizeof(fe_t), (void**)&buf));
    const b2r_fr* src[4] = {a, need_b ? b : nullptr, need_cd ? c : nullptr, need_cd ? d : nullptr};
    fe_t* dev[4] = {buf, nullptr, nullptr, nullptr};
    for (int k = 0; k < 4; k++) {
        if (!src[k]) continue;
        dev[k] = buf + (size_t)k * n;
        B2R_CUDA(ctx, cudaMemcpyAsync(dev[k], src[k], n * sizeof(fe_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    fe_t* o = buf + 4 * n;
    const unsigned grid = (unsigned)((n + 127) / 128);
    if (field == 0) k_field_selftest<Fr><<<grid, 128, 0, ctx->stream>>>(op, dev[0], dev[1], dev[2], dev[3], o, n);
    else k_field_selftest<Fq><<<grid, 128, 0, ctx->stream>>>(op, dev[0], dev[1], dev[2], dev[3], o, n);
    B2R_LAUNCH_CHECK(ctx);
    B2R_CUDA(ctx, cudaMemcpyAsync(out, o, n * sizeof(fe_t), cudaMemcpyDeviceToHost, ctx->stream));
    B2R_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
} B2R_ABI_CATCH(ctx)

}  // extern "C"
