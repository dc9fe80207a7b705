// Keygen and the full prover for the RSA pkcs1v15 circuit, batched over independent instances.
//
// Replaces, for a whole batch, what the reference's bench does around the two hot paths
// (benches/bench.rs:228-239 keygen_vk / keygen_pk, :319-331 create_proof with KZG + GWC + Blake2b):
// SURVEY.md 8f rows 1-3.  The protocol is halo2_proofs' (third-party crate, 2022-10 era; restated
// independently in oracle/plonk.py, which documents the constraint system and the proof layout):
//
//   phase 1  witness -> 5 advice columns (+ blinding rows) -> commit_lagrange            | theta
//   phase 2  per lookup: compress, permute (A', S')                                      | beta, gamma
//   phase 3  permutation and lookup grand products Z, random polynomial                  | y
//   phase 4  lagrange_to_coeff, coeff_to_extended, quotient on the 2^(k+2) coset, h      | x
//   phase 5  58 evaluations                                                              | v
//   phase 6  GWC witness polynomials W (linear combination + division by X - z) -> commit
//
// Everything a phase produces stays in HBM; between phases only commitments / evaluations go to the
// host, which owns the per-proof Blake2b transcripts (transcript.hpp) and sends the challenges back.
// Polynomials live in one arena P[slot][proof][row] so that every phase's commit / NTT is ONE batched
// call over contiguous vectors.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <thread>
#include <cstring>
#include <vector>

#include "ctx.hpp"
#include "devutil.cuh"
#include "ec.cuh"
#include "prog.hpp"
#include "transcript.hpp"

namespace b2r {
fe_t fr_omega(uint32_t k);
fe_t fr_zeta();
fe_t fr_from_u64(uint64_t v);
int32_t msm_batch_dev(b2r_ctx* ctx, const b2r_bases* bs, const fe_t* scalars_dev, size_t m, size_t n, affine_t* out_dev, bool uniform);
int32_t bases_register_suffix_sums(b2r_ctx* ctx, const b2r_bases* src, b2r_bases** out);
int32_t bases_register_rewindowed(b2r_ctx* ctx, const b2r_bases* src, uint32_t window, b2r_bases** out);
int32_t coset_ntt_grouped_dev(b2r_ctx* ctx, const fe_t* coeffs, size_t outer, uint64_t outer_stride, size_t inner, uint32_t k, uint32_t ext_k,
                              fe_t* out);
void bases_destroy(b2r_bases* bs);
size_t bases_count(const b2r_bases* bs);
int32_t witness_run(b2r_ctx* ctx, const b2r_prog* prog, const uint64_t* n_limbs_dev, const uint64_t* sig_limbs_dev,
                    const uint64_t* hash_limbs_dev, size_t batch, const BlindKey& bkey, b2r_fr* advice_dev, uint8_t* is_valid_dev,
                    size_t p_base, size_t p_stride, size_t col_stride);
}  // namespace b2r

using namespace b2r;

// ---- constraint system constants (mirrors oracle/plonk.py) -----------------------------------------
static constexpr int NADV = 5, NFIXED = 15, NPERM = 6, NLOOK = 5, NSETS = 2, CHUNK = 3, BF = 5, QD = 4;
enum { FX_SA = 0, FX_SB, FX_SC, FX_SD, FX_SE, FX_MUL_AB, FX_MUL_CD, FX_SE_NEXT, FX_CONST, FX_TAG_COMP, FX_TAG_OVER, FX_T_TAG,
       FX_T_VALUE, FX_S_COMP, FX_S_OVER };
// polynomial slots of the arena
enum { SL_ADV = 0, SL_LA = 5 /* A'_l = 5+2l, S'_l = 6+2l */, SL_PZ = 15, SL_LZ = 17, SL_RAND = 22, NSLOT = 23, NTRANS = 22 };
static constexpr int NZ = NSETS + NLOOK;  // grand products per proof
static constexpr int NEVAL = 58, NPOINTS = 4, MAXTERMS = 52;
static constexpr uint32_t MAX_TABLE = 1024;
#ifndef B2R_GP_CH
#define B2R_GP_CH 256
#endif
static constexpr int CH = B2R_GP_CH;  // rows per thread in the grand-product kernels (one Fermat inversion per CH rows: 1.5 products per row at 256)

__host__ __device__ inline int lookup_acol(int l) { return l < 4 ? l : 0; }
__host__ __device__ inline int lookup_ftag(int l) { return l < 4 ? FX_TAG_COMP : FX_TAG_OVER; }
__host__ __device__ inline int lookup_fsel(int l) { return l < 4 ? FX_S_COMP : FX_S_OVER; }

struct b2r_pk {
    const b2r_prog* prog = nullptr;
    const b2r_bases *g = nullptr, *gl = nullptr;
    b2r_bases* gl_sfx = nullptr;  // suffix sums of g_lagrange: commits the run-structured grand-product columns (msm.cu k_sfx_local)
    // narrow-window copies for the sparse commitments, whose cost is the bucket reduction (2^15 buckets per vector at
    // c = 16) rather than the additions: advice columns at c = 13, permuted lookup columns S' and the first differences
    // of A' at c = 10
    b2r_bases* gl_c13 = nullptr;
    b2r_bases* gl_c10 = nullptr;
    b2r_bases* gl_sfx_c10 = nullptr;
    uint32_t k = 0, ext_k = 0, n = 0, ext_n = 0, u = 0, T = 0;
    fe_t *fixed_values = nullptr, *fixed_polys = nullptr, *fixed_cosets = nullptr;
    fe_t *sigma_values = nullptr, *sigma_polys = nullptr, *sigma_cosets = nullptr;
    fe_t* l_cosets = nullptr;       // [3][ext_n]: l0, l_last, l_active
    uint8_t* range_tags = nullptr;  // [4][n]: s_comp, tag_comp, s_over, tag_over
    uint32_t* table = nullptr;      // [T]: tag << 16 | value
    uint32_t tag_base[16] = {0};
    uint8_t tag_bits[16] = {0};
    std::vector<affine_t> fixed_commitments, sigma_commitments;
    fe_t transcript_repr;
    fe_t delta_pows[NPERM];
    fe_t t_inv[4];
};

namespace b2r {

struct DevConsts {
    fe_t delta_pows[NPERM];
    fe_t t_inv[4];
    fe_t zeta;
    uint32_t tag_base[16];
    uint8_t tag_bits[16];
};

__device__ __forceinline__ fe_t omega_pow(const fe_t* tw, uint32_t n, uint32_t i) {
    // tw[e] = omega^e for e < n/2 ; omega^(n/2) = -1
    const uint32_t half = n >> 1;
    return i < half ? ldv_nc(tw + i) : Fr::neg(ldv_nc(tw + (i - half)));
}

// ---- keygen kernels ------------------------------------------------------------------------------------
__global__ void k_sigma(const uint32_t* __restrict__ mapping, fe_t* __restrict__ sigma, const fe_t* __restrict__ tw, uint32_t n, DevConsts C) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NPERM * n) return;
    uint32_t m = mapping[i], col = m / n, row = m % n;
    stv(sigma + i, Fr::mul(C.delta_pows[col], omega_pow(tw, n, row)));
}
__global__ void k_l_active(fe_t* l_last_to_active /* in: l_last copy, out: l_active */, const fe_t* __restrict__ l_blind, uint32_t ext_n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ext_n) return;
    fe_t v = Fr::sub(Fr::sub(Fr::one(), ldv(l_last_to_active + i)), ldv(l_blind + i));
    stv(l_last_to_active + i, v);
}

// ---- phase 2: lookups -----------------------------------------------------------------------------------
// histogram of the compressed input over table entries; the input of row i is (tag_i, sel_i * advice_i), so its
// table index follows from the tag and the (small) advice value without field arithmetic
__global__ void __launch_bounds__(256)
k_lookup_hist(const fe_t* __restrict__ P, const uint8_t* __restrict__ rt, uint32_t n, uint32_t u, uint32_t B, uint32_t T,
              uint32_t* __restrict__ hist, uint32_t* __restrict__ err, DevConsts C) {
    __shared__ uint32_t sh[MAX_TABLE];
    const uint32_t li = blockIdx.y, p = blockIdx.z, tid = threadIdx.x;
    for (uint32_t i = tid; i < T; i += 256) sh[i] = 0;
    __syncthreads();
    const uint32_t row = blockIdx.x * 256 + tid;
    if (row < u) {
        const uint8_t sel = rt[(li < 4 ? 0 : 2) * (size_t)n + row], tag = rt[(li < 4 ? 1 : 3) * (size_t)n + row];
        uint32_t idx = 0;
        bool bad = false;
        if (sel) {
            fe_t v = Fr::from_mont(ldv(P + ((size_t)(SL_ADV + lookup_acol(li)) * B + p) * n + row));
            uint32_t hi = v.l[1] | v.l[2] | v.l[3] | v.l[4] | v.l[5] | v.l[6] | v.l[7];
            uint32_t bits = C.tag_bits[tag & 15];
            if (hi || bits == 0 || (v.l[0] >> bits)) bad = true;
            else idx = C.tag_base[tag & 15] + v.l[0];
        } else if (tag) {
            bad = true;
        }
        if (bad) atomicOr(err + p, 1u);
        else atomicAdd(&sh[idx], 1u);
    }
    __syncthreads();
    uint32_t* h = hist + ((size_t)p * NLOOK + li) * MAX_TABLE;
    for (uint32_t i = tid; i < T; i += 256)
        if (sh[i]) atomicAdd(h + i, sh[i]);
}

// per proof: compressed table values theta * tag + value, sorted by canonical value (rank sort in one CTA)
__global__ void __launch_bounds__(1024)
k_table_sort(const uint32_t* __restrict__ table, uint32_t T, const fe_t* __restrict__ chal /* [B][8]: theta first */,
             fe_t* __restrict__ sorted_cv /* [B][MAX_TABLE] Montgomery */, uint32_t* __restrict__ order /* [B][MAX_TABLE] */) {
    __shared__ uint32_t canon[MAX_TABLE][8];
    const uint32_t p = blockIdx.x, t = threadIdx.x;
    fe_t cvm = Fr::zero();
    if (t < T) {
        const uint32_t e = table[t];
        fe_t tag = Fr::zero(), val = Fr::zero();
        tag.l[0] = e >> 16;
        val.l[0] = e & 0xffffu;
        cvm = Fr::add(Fr::mul(Fr::to_mont(tag), ldv(chal + (size_t)p * 8)), Fr::to_mont(val));
        fe_t c = Fr::from_mont(cvm);
        for (int i = 0; i < 8; i++) canon[t][i] = c.l[i];
    }
    __syncthreads();
    if (t < T) {
        uint32_t rank = 0;
        for (uint32_t o = 0; o < T; o++) {
            bool less = false, decided = false;
            for (int w = 7; w >= 0; w--) {
                uint32_t a = canon[o][w], b = canon[t][w];
                if (!decided && a != b) {
                    less = a < b;
                    decided = true;
                }
            }
            if (!decided) less = o < t;
            rank += less ? 1u : 0u;
        }
        stv(sorted_cv + (size_t)p * MAX_TABLE + rank, cvm);
        order[(size_t)p * MAX_TABLE + rank] = t;
    }
}

struct LookupPlan {
    uint32_t run_start[MAX_TABLE + 1];  // first A' row of the run of rank r (counts in rank order, prefix summed)
    uint32_t ne_prefix[MAX_TABLE];      // non-empty runs before rank r
    uint32_t unused[MAX_TABLE];         // ranks >= 1 with an empty run, ascending
    uint32_t z0, R, pad0, pad1;         // leftover zeros; number of repeated rows
};

__device__ __forceinline__ uint32_t block_excl_scan_1024(uint32_t v, uint32_t* sh /*1024*/, uint32_t* total) {
    const uint32_t t = threadIdx.x;
    sh[t] = v;
    __syncthreads();
    for (uint32_t d = 1; d < 1024; d <<= 1) {
        uint32_t o = t >= d ? sh[t - d] : 0;
        __syncthreads();
        sh[t] += o;
        __syncthreads();
    }
    uint32_t incl = sh[t];
    *total = sh[1023];
    __syncthreads();
    return incl - v;
}

__global__ void __launch_bounds__(1024)
k_lookup_plan(const uint32_t* __restrict__ hist, const uint32_t* __restrict__ order, uint32_t T, uint32_t u, LookupPlan* __restrict__ plans,
              uint32_t* __restrict__ err) {
    __shared__ uint32_t sh[1024];
    const uint32_t li = blockIdx.x, p = blockIdx.y, r = threadIdx.x;
    LookupPlan* pl = plans + (size_t)p * NLOOK + li;
    const uint32_t* h = hist + ((size_t)p * NLOOK + li) * MAX_TABLE;
    const uint32_t c = r < T ? h[order[(size_t)p * MAX_TABLE + r]] : 0;
    uint32_t total, ne_total, un_total;
    uint32_t start = block_excl_scan_1024(c, sh, &total);
    uint32_t ne = block_excl_scan_1024(c ? 1u : 0u, sh, &ne_total);
    const bool is_unused = r >= 1 && r < T && c == 0;
    uint32_t up = block_excl_scan_1024(is_unused ? 1u : 0u, sh, &un_total);
    if (r < T) {
        pl->run_start[r] = start;
        pl->ne_prefix[r] = ne;
        if (is_unused) pl->unused[up] = r;
    }
    if (r == 0) {
        pl->run_start[T] = total;
        const uint32_t m0 = u - (T - 1);  // multiplicity of the value 0 in the table column (unassigned rows are 0)
        pl->z0 = m0 - (c ? 1u : 0u);
        pl->R = u - ne_total;
        if (total != u || order[(size_t)p * MAX_TABLE] != 0) atomicOr(err + p, 2u);
    }
}

__global__ void __launch_bounds__(256)
k_lookup_fill(fe_t* __restrict__ P, const LookupPlan* __restrict__ plans, const fe_t* __restrict__ sorted_cv, uint32_t T, uint32_t n,
              uint32_t u, uint32_t B, const BlindKey seed, uint32_t p_base) {
    const uint32_t li = blockIdx.y, p = blockIdx.z, row = blockIdx.x * 256 + threadIdx.x;
    if (row >= n) return;
    fe_t* outA = P + ((size_t)(SL_LA + 2 * li) * B + p) * n;
    fe_t* outS = P + ((size_t)(SL_LA + 2 * li + 1) * B + p) * n;
    if (row >= u) {
        stv(outA + row, blind_value(seed, p_base + p, ST_LOOKUP_A + li, row));
        stv(outS + row, blind_value(seed, p_base + p, ST_LOOKUP_S + li, row));
        return;
    }
    const LookupPlan* pl = plans + (size_t)p * NLOOK + li;
    const fe_t* cv = sorted_cv + (size_t)p * MAX_TABLE;
    // largest r with run_start[r] <= row (empty runs share their start with the next run: take the last one)
    uint32_t lo = 0, hi = T;  // invariant run_start[lo] <= row < run_start[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (pl->run_start[mid] <= row) lo = mid; else hi = mid;
    }
    const uint32_t r = lo;
    const fe_t a = ldv(cv + r);
    stv(outA + row, a);
    fe_t s;
    if (row == pl->run_start[r]) {
        s = a;
    } else {
        const uint32_t rr = row - pl->ne_prefix[r] - 1;  // rank of this row among the repeated rows
        const uint32_t j = pl->R - 1 - rr;               // leftover table elements are handed out from the last repeated row down
        s = j < pl->z0 ? Fr::zero() : ldv(cv + pl->unused[j - pl->z0]);
    }
    stv(outS + row, s);
}

// ---- phase 3: grand products ------------------------------------------------------------------------------
// chal[p] = {theta, beta, gamma, y, x, v, -, -}
__global__ void __launch_bounds__(256)
k_perm_numden(const fe_t* __restrict__ P, const fe_t* __restrict__ sigma_values, const fe_t* __restrict__ tw, const fe_t* __restrict__ chal,
              uint32_t n, uint32_t B, fe_t* __restrict__ num, fe_t* __restrict__ den, DevConsts C) {
    const uint32_t s = blockIdx.y, p = blockIdx.z, row = blockIdx.x * 256 + threadIdx.x;
    // beta * delta^c of this proof's columns, once per CTA (one product per row and column less)
    __shared__ uint4 bd_raw[CHUNK * 2];
    fe_t* bd = reinterpret_cast<fe_t*>(bd_raw);
    const fe_t beta = ldv(chal + (size_t)p * 8 + 1), gamma = ldv(chal + (size_t)p * 8 + 2);
    if (threadIdx.x < CHUNK && s * CHUNK + threadIdx.x < NPERM) stv(bd + threadIdx.x, Fr::mul(C.delta_pows[s * CHUNK + threadIdx.x], beta));
    __syncthreads();
    if (row >= n) return;
    const fe_t w = omega_pow(tw, n, row);
    fe_t nu = Fr::one(), de = Fr::one();
    for (int c = s * CHUNK; c < (int)(s + 1) * CHUNK && c < NPERM; c++) {
        fe_t v = c < NADV ? ldv(P + ((size_t)(SL_ADV + c) * B + p) * n + row) : Fr::zero();  // instance column: empty
        fe_t vg = Fr::add(v, gamma);
        const fe_t fn = Fr::add(Fr::mul(ldv(bd + (c - s * CHUNK)), w), vg);
        const fe_t fd = Fr::add(Fr::mul(beta, ldv_nc(sigma_values + (size_t)c * n + row)), vg);
        nu = c == (int)(s * CHUNK) ? fn : Fr::mul(nu, fn);   // the first factor needs no product
        de = c == (int)(s * CHUNK) ? fd : Fr::mul(de, fd);
    }
    const size_t o = ((size_t)s * B + p) * n + row;
    stv(num + o, nu);
    stv(den + o, de);
}
__global__ void __launch_bounds__(256)
k_lookup_numden(const fe_t* __restrict__ P, const fe_t* __restrict__ fixed_values, const fe_t* __restrict__ chal, uint32_t n, uint32_t B,
                fe_t* __restrict__ num, fe_t* __restrict__ den) {
    const uint32_t li = blockIdx.y, p = blockIdx.z, row = blockIdx.x * 256 + threadIdx.x;
    if (row >= n) return;
    const fe_t theta = ldv(chal + (size_t)p * 8), beta = ldv(chal + (size_t)p * 8 + 1), gamma = ldv(chal + (size_t)p * 8 + 2);
    const fe_t adv = ldv(P + ((size_t)(SL_ADV + lookup_acol(li)) * B + p) * n + row);
    const fe_t A = Fr::mul_add_mul(ldv_nc(fixed_values + (size_t)lookup_ftag(li) * n + row), theta,
                                   ldv_nc(fixed_values + (size_t)lookup_fsel(li) * n + row), adv);
    const fe_t S = Fr::add(Fr::mul(ldv_nc(fixed_values + (size_t)FX_T_TAG * n + row), theta), ldv_nc(fixed_values + (size_t)FX_T_VALUE * n + row));
    const fe_t ap = ldv(P + ((size_t)(SL_LA + 2 * li) * B + p) * n + row), sp = ldv(P + ((size_t)(SL_LA + 2 * li + 1) * B + p) * n + row);
    const size_t o = ((size_t)(NSETS + li) * B + p) * n + row;
    stv(num + o, Fr::mul(Fr::add(A, beta), Fr::add(S, gamma)));
    stv(den + o, Fr::mul(Fr::add(ap, beta), Fr::add(sp, gamma)));
}
// thread = CH consecutive rows of one grand product: Montgomery batch inversion of den, ratio = num / den written over
// den, product of the chunk's ratios (rows < u only) to chunk_prod
__global__ void __launch_bounds__(128)
k_batch_inv_ratio(const fe_t* __restrict__ num, fe_t* __restrict__ den, fe_t* __restrict__ pref, uint32_t n, uint32_t u, uint32_t nz_total,
                  fe_t* __restrict__ chunk_prod) {
    const uint32_t nch = n / CH;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nz_total * nch) return;
    const uint32_t z = t / nch, c = t % nch;
    const size_t base = (size_t)z * n + (size_t)c * CH;
    fe_t acc = Fr::one();
    for (int i = 0; i < CH; i++) {
        stv(pref + base + i, acc);
        fe_t d = ldv(den + base + i);
        if (!Fr::is_zero(d)) acc = Fr::mul(acc, d);
    }
    fe_t inv = Fr::inv(acc);
    fe_t prod = Fr::one();
    for (int i = CH - 1; i >= 0; i--) {
        fe_t d = ldv(den + base + i);
        fe_t r = Fr::zero();
        if (!Fr::is_zero(d)) {
            r = Fr::mul(Fr::mul(inv, ldv(pref + base + i)), ldv(num + base + i));
            inv = Fr::mul(inv, d);
        }
        stv(den + base + i, r);
        if (c * CH + i < u) prod = Fr::mul(prod, r);
    }
    stv(chunk_prod + (size_t)z * nch + c, prod);
}
// exclusive prefix products of the chunk products of one grand product (one CTA each)
__global__ void __launch_bounds__(1024) k_chunk_scan(fe_t* __restrict__ chunk_prod, uint32_t nch) {
    extern __shared__ uint4 smem_raw[];
    fe_t* sh = reinterpret_cast<fe_t*>(smem_raw);
    const uint32_t z = blockIdx.x, t = threadIdx.x;
    const uint32_t per = (nch + 1023) / 1024;
    fe_t* cp = chunk_prod + (size_t)z * nch;
    fe_t local = Fr::one();
    for (uint32_t i = 0; i < per; i++) {
        uint32_t c = t * per + i;
        if (c < nch) local = Fr::mul(local, ldv(cp + c));
    }
    fe_t run = local;
    stv(sh + t, run);
    __syncthreads();
    for (uint32_t d = 1; d < 1024; d <<= 1) {
        fe_t o = Fr::one();
        if (t >= d) o = ldv(sh + t - d);
        __syncthreads();
        run = Fr::mul(run, o);
        stv(sh + t, run);
        __syncthreads();
    }
    fe_t excl = t ? ldv(sh + t - 1) : Fr::one();
    for (uint32_t i = 0; i < per; i++) {
        uint32_t c = t * per + i;
        if (c < nch) {
            fe_t v = ldv(cp + c);
            stv(cp + c, excl);
            excl = Fr::mul(excl, v);
        }
    }
}
__global__ void __launch_bounds__(128)
k_z_write(fe_t* __restrict__ P, const fe_t* __restrict__ ratio, const fe_t* __restrict__ chunk_prefix, uint32_t n, uint32_t B, uint32_t nz_total,
          const BlindKey seed, uint32_t p_base) {
    const uint32_t nch = n / CH;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nz_total * nch) return;
    const uint32_t z = t / nch, c = t % nch;  // z = zslot * B + p
    const uint32_t zslot = z / B, p = z % B;
    const size_t base = (size_t)z * n + (size_t)c * CH;
    fe_t* out = P + ((size_t)(SL_PZ + zslot) * B + p) * n + (size_t)c * CH;
    const uint32_t stream = zslot < NSETS ? ST_PERM_Z + zslot : ST_LOOKUP_Z + (zslot - NSETS);
    fe_t cur = ldv(chunk_prefix + (size_t)z * nch + c);
    for (int i = 0; i < CH; i++) {
        const uint32_t row = c * CH + i;
        if (row >= n - BF) {
            stv(out + i, blind_value(seed, p_base + p, stream, row));
        } else {
            stv(out + i, cur);
            cur = Fr::mul(cur, ldv(ratio + base + i));
        }
    }
}
// second permutation set continues the first: z_1[row] *= z_0[u] for row <= u
__global__ void __launch_bounds__(256) k_perm_chain(fe_t* __restrict__ P, uint32_t n, uint32_t u, uint32_t B, uint32_t set) {
    const uint32_t p = blockIdx.y, row = blockIdx.x * 256 + threadIdx.x;
    if (row > u) return;
    const fe_t last = ldv(P + ((size_t)(SL_PZ + set - 1) * B + p) * n + u);
    fe_t* z = P + ((size_t)(SL_PZ + set) * B + p) * n + row;
    stv(z, Fr::mul(ldv(z), last));
}
__global__ void __launch_bounds__(256) k_random_poly(fe_t* __restrict__ P, uint32_t n, uint32_t B, const BlindKey seed, uint32_t p_base) {
    const uint32_t p = blockIdx.y, row = blockIdx.x * 256 + threadIdx.x;
    if (row >= n) return;
    stv(P + ((size_t)SL_RAND * B + p) * n + row, blind_value(seed, p_base + p, ST_RANDOM_POLY, row));
}

// first differences of the grand-product columns: D[z][row] = Z[row] - Z[row - 1].  Z is constant wherever the ratio is 1
// (rows in no copy constraint / lookup, all rows behind the circuit), so D is zero there and
// commit(Z) = sum_row D[row] * (sum_{i >= row} g_lagrange[i]) needs bucket entries only for the rows where Z moves.
__global__ void __launch_bounds__(256) k_run_diff(const fe_t* __restrict__ Z /* [nz][n] */, fe_t* __restrict__ D, uint32_t n) {
    const uint32_t z = blockIdx.y, row = blockIdx.x * 256 + threadIdx.x;
    if (row >= n) return;
    const fe_t* a = Z + (size_t)z * n;
    const fe_t cur = ldv(a + row);
    stv(D + (size_t)z * n + row, row ? Fr::sub(cur, ldv(a + row - 1)) : cur);
}

// ---- device-resident commitment block ---------------------------------------------------------------------------
// log[(p_base + p) * NCOMMIT + off + j] = the j-th commitment of this phase for proof p, read from the phase's MSM output:
// mode 0: src[j * B + p] (column-major phases), 1: src[p * J + j] (per-proof phases), 2: the permuted-lookup phase, whose MSM
// output holds all A' then all S' while the transcript interleaves them (A'_0, S'_0, A'_1, ...)
static constexpr uint32_t NCOMMIT = NADV + 2 * NLOOK + (NZ + 1) + QD + NPOINTS;
__global__ void __launch_bounds__(256) k_log_commitments(const affine_t* __restrict__ src, affine_t* __restrict__ log, uint32_t B, uint32_t J, uint32_t off,
                                                        uint32_t p_base, uint32_t mode) {
    const uint32_t t = blockIdx.x * 256 + threadIdx.x;
    if (t >= B * J) return;
    const uint32_t p = t / J, j = t % J;
    const uint32_t si = mode == 0 ? j * B + p : mode == 1 ? p * J + j : ((j & 1u) * (J / 2) + (j >> 1)) * B + p;
    log[(size_t)(p_base + p) * NCOMMIT + off + j] = src[si];
}

// ---- phase 4: quotient on the extended coset --------------------------------------------------------------------
struct QuotArgs {
    const fe_t* E;        // [NTRANS][QB][ext_n]
    const fe_t* fixed_c;  // [NFIXED][ext_n]
    const fe_t* sigma_c;  // [NPERM][ext_n]
    const fe_t* l_c;      // [3][ext_n]
    const fe_t* tw_ext;   // omega_ext^e, e < ext_n / 2
    const fe_t* chal;     // [QB][8] (already offset to the sub-batch)
    const fe_t* ypow;     // [QB][NCONS]: y^(NCONS - 1 - k) for constraint k (already offset to the sub-batch)
    fe_t* h;              // [QB][ext_n]
    uint32_t ext_n, step, QB;
};
// constraints in halo2's order: gate, 3 permutation boundary terms, NSETS product terms, 5 per lookup
static constexpr int NCONS = 1 + 3 + NSETS + 5 * NLOOK;
static_assert(NSETS == 2, "constraint numbering below assumes two permutation sets");
// h = sum_k C_k y^(NCONS-1-k) / (X^n - 1).  halo2 folds the constraints by Horner in y (two products per constraint);
// every C_k here is l(X) * e_k with l one of {1, l0, l_last, l_active}, so the sum is regrouped as
// gate y^30 + l0 sum(e_k y^..) + l_last sum(..) + l_active sum(..): one product per constraint plus three.
#ifndef B2R_QMINB
#define B2R_QMINB 3
#endif
__global__ void __launch_bounds__(128, B2R_QMINB) k_quotient(const QuotArgs A, const DevConsts C) {
    const uint32_t q = blockIdx.y, i = blockIdx.x * 128 + threadIdx.x;
    if (i >= A.ext_n) return;
    const uint32_t mask = A.ext_n - 1, nx = (i + A.step) & mask, pv = (i - A.step) & mask, lastr = (i - (BF + 1) * A.step) & mask;
    const fe_t theta = ldv(A.chal + (size_t)q * 8), beta = ldv(A.chal + (size_t)q * 8 + 1), gamma = ldv(A.chal + (size_t)q * 8 + 2);
    const fe_t* yp = A.ypow + (size_t)q * NCONS;
    auto ext = [&](int slot, uint32_t idx) { return ldv(A.E + ((size_t)slot * A.QB + q) * A.ext_n + idx); };
    auto fx = [&](int col) { return ldv_nc(A.fixed_c + (size_t)col * A.ext_n + i); };
    auto wy = [&](const fe_t& e, int k) { return Fr::mul(e, ldv_nc(yp + k)); };
    // the advice values are re-read where they are used (L1 hits) instead of being held in 40 registers across the kernel
    auto adv = [&](int c) { return c < NADV ? ext(SL_ADV + c, i) : Fr::zero(); };
    // Sums of products share ONE Montgomery reduction (Fr::dot4 / mul_add_mul / mul_sub_mul: 64 wide MADs per product
    // plus 64 per sum instead of 128 per product) - 28 of the kernel's 110 reductions go, values unchanged.
    auto yw = [&](int k) { return ldv_nc(yp + k); };
    // main gate
    fe_t gate;
    {
        const fe_t a = adv(0), b = adv(1), c = adv(2);
        gate = Fr::dot4(a, fx(FX_SA), b, fx(FX_SB), Fr::mul(a, b), fx(FX_MUL_AB), c, fx(FX_SC));
        const fe_t d = adv(3);
        gate = Fr::add(gate, Fr::dot4(d, fx(FX_SD), Fr::mul(c, d), fx(FX_MUL_CD), adv(4), fx(FX_SE), ext(SL_ADV + 4, nx), fx(FX_SE_NEXT)));
    }
    gate = Fr::add(gate, fx(FX_CONST));
    const fe_t one = Fr::one();
    // permutation argument
    fe_t pz[NSETS];
    for (int s = 0; s < NSETS; s++) pz[s] = ext(SL_PZ + s, i);
    fe_t g0 = Fr::mul_add_mul(Fr::sub(one, pz[0]), yw(1), Fr::sub(pz[1], ext(SL_PZ, lastr)), yw(3));   // l0 group
    fe_t glast = wy(Fr::sub(Fr::sqr(pz[NSETS - 1]), pz[NSETS - 1]), 2);                               // l_last group
    // X at this point of the coset: zeta * omega_ext^i
    const uint32_t half = A.ext_n >> 1;
    fe_t xi = i < half ? ldv_nc(A.tw_ext + i) : Fr::neg(ldv_nc(A.tw_ext + (i - half)));
    const fe_t beta_x = Fr::mul(beta, Fr::mul(xi, C.zeta));
    fe_t gact;                                                                              // l_active group
    {
        fe_t pd[NSETS];
        for (int s = 0; s < NSETS; s++) {
            fe_t left = ext(SL_PZ + s, nx), right = pz[s];
            const int c_end = (s + 1) * CHUNK < NPERM ? (s + 1) * CHUNK : NPERM;
            for (int c = s * CHUNK; c < c_end; c++) {
                const fe_t vg = Fr::add(adv(c), gamma);
                const fe_t fl = Fr::add(Fr::mul(beta, ldv_nc(A.sigma_c + (size_t)c * A.ext_n + i)), vg);
                const fe_t fr = Fr::add(Fr::mul(beta_x, C.delta_pows[c]), vg);
                if (c + 1 < c_end) {
                    left = Fr::mul(left, fl);
                    right = Fr::mul(right, fr);
                } else {
                    pd[s] = Fr::mul_sub_mul(left, fl, right, fr);   // left - right of the set
                }
            }
        }
        gact = Fr::mul_add_mul(pd[0], yw(4), pd[1], yw(5));
    }
    // lookup arguments
    const fe_t tbl_g = Fr::add(Fr::add(Fr::mul(fx(FX_T_TAG), theta), fx(FX_T_VALUE)), gamma);
    const fe_t tag_c = Fr::mul(fx(FX_TAG_COMP), theta), tag_o = Fr::mul(fx(FX_TAG_OVER), theta);
    const fe_t s_c = fx(FX_S_COMP), s_o = fx(FX_S_OVER);
#pragma unroll 1
    for (int l = 0; l < NLOOK; l++) {
        const int k0 = 4 + NSETS + 5 * l;
        const fe_t z = ext(SL_LZ + l, i), ap = ext(SL_LA + 2 * l, i), sp = ext(SL_LA + 2 * l + 1, i);
        const fe_t d = Fr::sub(ap, sp);
        g0 = Fr::add(g0, Fr::mul_add_mul(Fr::sub(one, z), yw(k0), d, yw(k0 + 3)));
        glast = Fr::add(glast, wy(Fr::sub(Fr::sqr(z), z), k0 + 1));
        const fe_t inp = l < 4 ? Fr::add(tag_c, Fr::mul(s_c, adv(l))) : Fr::add(tag_o, Fr::mul(s_o, adv(0)));
        const fe_t lr = Fr::mul_sub_mul(Fr::mul(ext(SL_LZ + l, nx), Fr::add(ap, beta)), Fr::add(sp, gamma), Fr::mul(z, Fr::add(inp, beta)), tbl_g);
        gact = Fr::add(gact, Fr::mul_add_mul(lr, yw(k0 + 2), Fr::mul(d, Fr::sub(ap, ext(SL_LA + 2 * l, pv))), yw(k0 + 4)));
    }
    const fe_t acc = Fr::dot4(gate, yw(0), ldv_nc(A.l_c + i), g0, ldv_nc(A.l_c + (size_t)A.ext_n + i), glast, ldv_nc(A.l_c + 2 * (size_t)A.ext_n + i), gact);
    stv(A.h + (size_t)q * A.ext_n + i, Fr::mul(acc, C.t_inv[i & (A.step - 1)]));
}

// ---- phase 5 / 6: evaluations, linear combinations, division by (X - z) -------------------------------------------
// polynomial reference: kind 0 = arena slot (per proof), 1 = shared proving-key polynomial, 2 = h piece (per proof)
struct PolyRef {
    uint32_t kind, idx;
};
struct PolyTable {
    const fe_t* P;      // arena, [NSLOT][B][n]
    const fe_t* fixedp; // [NFIXED][n]
    const fe_t* sigmap; // [NPERM][n]
    const fe_t* hbuf;   // [B][QD][n]
    uint32_t n, B;
};
__device__ __forceinline__ const fe_t* poly_ptr(const PolyTable& T, PolyRef r, uint32_t p) {
    if (r.kind == 0) return T.P + ((size_t)r.idx * T.B + p) * T.n;
    if (r.kind == 1) return r.idx < NFIXED ? T.fixedp + (size_t)r.idx * T.n : T.sigmap + (size_t)(r.idx - NFIXED) * T.n;
    return T.hbuf + ((size_t)p * QD + r.idx) * T.n;
}
struct EvalPlan {
    PolyRef poly[NEVAL];
    uint8_t point[NEVAL];
};
__device__ __forceinline__ fe_t fr_pow_u32(fe_t a, uint32_t e) {
    fe_t acc = Fr::one();
    while (e) {
        if (e & 1u) acc = Fr::mul(acc, a);
        a = Fr::sqr(a);
        e >>= 1;
    }
    return acc;
}
// evals[p][e] = poly_e(point_e).  Thread t sums coefficients t, t + 256, ...: S_t = sum_j c[256 j + t] w^j with w = z^256, then
// S_t z^t and a block tree sum.  The powers w^j (at most 512 of them, 16 KB) are built once per CTA in shared memory, so the
// inner sum is a dot product taken four terms per Montgomery reduction (Fr::dot4: 80 instead of the 128 wide MADs per
// coefficient of a Horner step); blocks of 512 powers are chained by Horner in w^512 for larger n.
static constexpr uint32_t EVAL_WB = 512;
__global__ void __launch_bounds__(256)
k_eval(const PolyTable T, const EvalPlan* __restrict__ plan, const fe_t* __restrict__ points /* [B][NPOINTS] */, fe_t* __restrict__ evals) {
    __shared__ uint4 smem_raw[256 * 2];
    __shared__ uint4 pw_raw[EVAL_WB * 2];
    fe_t* sh = reinterpret_cast<fe_t*>(smem_raw);
    fe_t* W = reinterpret_cast<fe_t*>(pw_raw);
    const uint32_t e = blockIdx.x, p = blockIdx.y, t = threadIdx.x;
    const fe_t* c = poly_ptr(T, plan->poly[e], p);
    const fe_t z = ldv(points + (size_t)p * NPOINTS + plan->point[e]);
    fe_t w = z;
    for (int i = 0; i < 8; i++) w = Fr::sqr(w);  // z^256
    const uint32_t m = T.n / 256;                // coefficients per thread
    const uint32_t wb = m < EVAL_WB ? m : EVAL_WB;
    for (uint32_t j = t; j < wb; j += 256) stv(W + j, fr_pow_u32(w, j));
    const fe_t w_blk = fr_pow_u32(w, wb);
    __syncthreads();
    fe_t acc = Fr::zero();
    for (int blk = (int)(m / wb) - 1; blk >= 0; blk--) {
        const fe_t* cb = c + (size_t)blk * wb * 256 + t;
        fe_t s = Fr::zero();
        uint32_t j = 0;
        for (; j + 4 <= wb; j += 4)
            s = Fr::add(s, Fr::dot4(ldv(cb + (size_t)j * 256), ldv(W + j), ldv(cb + (size_t)(j + 1) * 256), ldv(W + j + 1),
                                    ldv(cb + (size_t)(j + 2) * 256), ldv(W + j + 2), ldv(cb + (size_t)(j + 3) * 256), ldv(W + j + 3)));
        for (; j < wb; j++) s = Fr::add(s, Fr::mul(ldv(cb + (size_t)j * 256), ldv(W + j)));
        acc = Fr::add(Fr::mul(acc, w_blk), s);
    }
    acc = Fr::mul(acc, fr_pow_u32(z, t));
    stv(sh + t, acc);
    __syncthreads();
    for (uint32_t s = 128; s > 0; s >>= 1) {
        if (t < s) stv(sh + t, Fr::add(ldv(sh + t), ldv(sh + t + s)));
        __syncthreads();
    }
    if (t == 0) stv(evals + (size_t)p * NEVAL + e, ldv(sh));
}
struct LincombPlan {
    PolyRef poly[NPOINTS][MAXTERMS];
    uint32_t nterms[NPOINTS];
};
// out[p][g][row] = sum_j scal[p][g][j] * poly_{g,j}[row].  The scalars and polynomial base pointers of the (proof, point)
// pair are staged in shared memory once per CTA (the inner loop then has no dependent global loads for them), and the sum
// takes four terms per Montgomery reduction.
__global__ void __launch_bounds__(256)
k_lincomb(const PolyTable T, const LincombPlan* __restrict__ plan, const fe_t* __restrict__ scal /* [B][NPOINTS][MAXTERMS] */, fe_t* __restrict__ out) {
    __shared__ uint4 sc_raw[MAXTERMS * 2];
    __shared__ const fe_t* base[MAXTERMS];
    fe_t* sc = reinterpret_cast<fe_t*>(sc_raw);
    const uint32_t g = blockIdx.y, p = blockIdx.z, row = blockIdx.x * 256 + threadIdx.x;
    const uint32_t m = plan->nterms[g];
    if (threadIdx.x < m) {
        stv(sc + threadIdx.x, ldv_nc(scal + ((size_t)p * NPOINTS + g) * MAXTERMS + threadIdx.x));
        base[threadIdx.x] = poly_ptr(T, plan->poly[g][threadIdx.x], p);
    }
    __syncthreads();
    if (row >= T.n) return;
    auto term = [&](uint32_t j) { return ldv(base[j] + row); };
    fe_t acc = Fr::zero();
    uint32_t j = 0;
    for (; j + 4 <= m; j += 4)   // four terms per Montgomery reduction
        acc = Fr::add(acc, Fr::dot4(ldv(sc + j), term(j), ldv(sc + j + 1), term(j + 1), ldv(sc + j + 2), term(j + 2), ldv(sc + j + 3), term(j + 3)));
    for (; j < m; j++) acc = Fr::add(acc, Fr::mul(ldv(sc + j), term(j)));
    stv(out + ((size_t)p * NPOINTS + g) * T.n + row, acc);
}
// q = (f(X) - f(z)) / (X - z): q_{i-1} = c_i + z q_i.  One CTA per polynomial; thread = n/256 consecutive coefficients:
// local Horner, suffix scan across threads with ratio z^m, second Horner pass that writes the quotient.
__global__ void __launch_bounds__(256)
k_kate(const fe_t* __restrict__ f /* [B][NPOINTS][n] */, const fe_t* __restrict__ points, uint32_t n, fe_t* __restrict__ qout) {
    __shared__ uint4 smem_raw[256 * 2];
    fe_t* sh = reinterpret_cast<fe_t*>(smem_raw);
    const uint32_t g = blockIdx.x, p = blockIdx.y, t = threadIdx.x;
    const uint32_t m = n / 256;
    const fe_t* c = f + ((size_t)p * NPOINTS + g) * n;
    fe_t* q = qout + ((size_t)p * NPOINTS + g) * n;
    const fe_t z = ldv(points + (size_t)p * NPOINTS + g);
    const uint32_t start = t * m;
    fe_t L = Fr::zero();
    for (int i = (int)m - 1; i >= 0; i--) L = Fr::add(Fr::mul(L, z), ldv(c + start + i));
    // S_t = sum_{d >= 1} L_{t+d} w^(d-1), w = z^m
    stv(sh + t, L);
    __syncthreads();
    fe_t S = t + 1 < 256 ? ldv(sh + t + 1) : Fr::zero();
    __syncthreads();
    fe_t w = fr_pow_u32(z, m);
    stv(sh + t, S);
    __syncthreads();
    for (uint32_t d = 1; d < 256; d <<= 1) {
        fe_t o = t + d < 256 ? ldv(sh + t + d) : Fr::zero();
        __syncthreads();
        S = Fr::add(S, Fr::mul(w, o));
        stv(sh + t, S);
        __syncthreads();
        w = Fr::sqr(w);
    }
    fe_t acc = S;
    for (int i = (int)m - 1; i >= 0; i--) {
        acc = Fr::add(Fr::mul(acc, z), ldv(c + start + i));
        if (start + i >= 1) stv(q + start + i - 1, acc);
    }
    if (t == 255) stv(q + n - 1, Fr::zero());
}

// ---- host helpers ---------------------------------------------------------------------------------------------
static fe_t u256_to_mont_fr(const U256& v) {
    fe_t c;
    for (int i = 0; i < 4; i++) {
        c.l[2 * i] = (uint32_t)v.l[i];
        c.l[2 * i + 1] = (uint32_t)(v.l[i] >> 32);
    }
    return Fr::to_mont(c);
}
static fe_t fr_delta() {
    // DELTA = 7^(2^28): 0x09226b6e22c6f0ca64ec26aad4c86e715b5f898e5e963f25870e56bbe533e9a2
    fe_t c;
    const uint32_t w[8] = {0xe533e9a2u, 0x870e56bbu, 0x5e963f25u, 0x5b5f898eu, 0xd4c86e71u, 0x64ec26aau, 0x22c6f0cau, 0x09226b6eu};
    for (int i = 0; i < 8; i++) c.l[i] = w[i];
    return Fr::to_mont(c);
}
static fe_t fr_pow_host(fe_t a, uint64_t e) {
    fe_t acc = Fr::one();
    while (e) {
        if (e & 1) acc = Fr::mul(acc, a);
        a = Fr::sqr(a);
        e >>= 1;
    }
    return acc;
}
static void compress_point_bytes(const affine_t& pt, uint8_t out[32]) {
    fe_t x = Fq::from_mont(pt.x), y = Fq::from_mont(pt.y);
    fe_to_bytes(x, out);
    out[31] |= (uint8_t)((y.l[0] & 1u) << 7);
}

static DevConsts make_consts(const b2r_pk* pk) {
    DevConsts C;
    for (int i = 0; i < NPERM; i++) C.delta_pows[i] = pk->delta_pows[i];
    for (int i = 0; i < 4; i++) C.t_inv[i] = pk->t_inv[i];
    C.zeta = fr_zeta();
    for (int i = 0; i < 16; i++) C.tag_base[i] = pk->tag_base[i], C.tag_bits[i] = pk->tag_bits[i];
    return C;
}

// halo2 permutation::keygen::Assembly::copy: merge the smaller cycle into the larger, relabel, splice
static void build_permutation(uint32_t n, const std::vector<std::array<uint32_t, 4>>& copies, std::vector<uint32_t>& mapping) {
    const size_t cells = (size_t)NPERM * n;
    mapping.resize(cells);
    std::vector<uint32_t> aux(cells), sizes(cells, 1);
    for (size_t i = 0; i < cells; i++) mapping[i] = aux[i] = (uint32_t)i;
    for (const auto& cp : copies) {
        const uint32_t left = cp[0] * n + cp[1], right = cp[2] * n + cp[3];
        uint32_t lc = aux[left], rc = aux[right];
        if (lc == rc) continue;
        if (sizes[lc] < sizes[rc]) std::swap(lc, rc);
        sizes[lc] += sizes[rc];
        uint32_t i = rc;
        for (;;) {
            aux[i] = lc;
            i = mapping[i];
            if (i == rc) break;
        }
        std::swap(mapping[left], mapping[right]);
    }
}

}  // namespace b2r

extern "C" int32_t b2r_intt_fr_batch_dev(b2r_ctx* ctx, b2r_fr* a_dev, size_t batch, uint32_t k);
extern "C" int32_t b2r_coset_ntt_fr_batch_dev(b2r_ctx* ctx, const b2r_fr* coeffs_dev, size_t batch, uint32_t k, uint32_t ext_k, b2r_fr* out_dev);
extern "C" int32_t b2r_coset_intt_fr_batch_dev(b2r_ctx* ctx, b2r_fr* a_dev, size_t batch, uint32_t ext_k);

static void pk_release(b2r_pk* pk) {
    cudaFree(pk->fixed_values); cudaFree(pk->fixed_polys); cudaFree(pk->fixed_cosets);
    cudaFree(pk->sigma_values); cudaFree(pk->sigma_polys); cudaFree(pk->sigma_cosets);
    cudaFree(pk->l_cosets); cudaFree(pk->range_tags); cudaFree(pk->table);
    bases_destroy(pk->gl_sfx);
    bases_destroy(pk->gl_c13);
    bases_destroy(pk->gl_c10);
    bases_destroy(pk->gl_sfx_c10);
    delete pk;
}

extern "C" {

int32_t b2r_rsa_keygen(b2r_ctx* ctx, const b2r_prog* prog, const b2r_bases* g, const b2r_bases* g_lagrange, b2r_pk** out) try {
    B2R_ENTER(ctx);
    if (!prog || !g || !g_lagrange || !out) return fail(ctx, B2R_ERR_INVALID, "rsa_keygen: null argument");
    *out = nullptr;
    // the two base sets must be the SRS of THIS circuit size: g needs >= 2^k points (commit of degree < 2^k polynomials),
    // g_lagrange exactly 2^k (a Lagrange basis of another domain commits to something else without any error)
    if (bases_count(g) < ((size_t)1 << prog->k)) return fail(ctx, B2R_ERR_INVALID, "rsa_keygen: g holds fewer than 2^k points");
    if (bases_count(g_lagrange) != ((size_t)1 << prog->k)) return fail(ctx, B2R_ERR_INVALID, "rsa_keygen: g_lagrange must hold exactly 2^k points");
    b2r_pk* pk = new b2r_pk();
    pk->prog = prog;
    pk->g = g;
    pk->gl = g_lagrange;
    pk->k = prog->k;
    pk->ext_k = prog->k + 2;  // degree 5 -> quotient degree 4 -> 2^(k+2)
    const uint32_t n = pk->n = 1u << pk->k, ext_n = pk->ext_n = 1u << pk->ext_k;
    pk->u = n - (BF + 1);
    cudaStream_t st = ctx->stream;
    auto bail = [&](int32_t rc) { pk_release(pk); return rc; };
#define KG_CUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return bail(cuda_fail(ctx, _e, #call)); } while (0)
#define KG_TRY(expr) do { int32_t _r = (expr); if (_r) return bail(_r); } while (0)
    // lookup table (RangeChip::load_table): (0, 0), then every value of every tag in ascending tag order
    std::vector<uint32_t> table(1, 0);
    for (int tag = 1; tag < 16; tag++) {
        pk->tag_bits[tag] = prog->tag_bits[tag];
        if (!prog->tag_bits[tag]) continue;
        if (prog->tag_bits[tag] > 15) return bail(fail(ctx, B2R_ERR_INVALID, "rsa_keygen: lookup tag wider than 15 bits"));
        pk->tag_base[tag] = (uint32_t)table.size();
        for (uint32_t v = 0; v < (1u << prog->tag_bits[tag]); v++) table.push_back(((uint32_t)tag << 16) | v);
    }
    pk->T = (uint32_t)table.size();
    if (pk->T > MAX_TABLE || pk->T - 1 > pk->u) return bail(fail(ctx, B2R_ERR_LAYOUT, "rsa_keygen: lookup table does not fit"));
    if (n < 1024) return bail(fail(ctx, B2R_ERR_INVALID, "rsa_keygen: k < 10"));
    // fixed columns (Lagrange values)
    std::vector<fe_t> cm(prog->constants.size());
    for (size_t i = 0; i < cm.size(); i++) cm[i] = u256_to_mont_fr(prog->constants[i]);
    fe_t small[16];
    for (int i = 0; i < 16; i++) small[i] = fr_from_u64(i);
    std::vector<fe_t> fixed((size_t)NFIXED * n, Fr::zero());
    std::vector<uint8_t> rt((size_t)4 * n, 0);
    for (size_t r = 0; r < prog->fixed.size(); r++) {
        for (int c = 0; c < 9; c++) fixed[(size_t)c * n + r] = cm[prog->fixed[r][c]];
        const auto& t = prog->range_tags[r];
        rt[r] = t[0]; rt[(size_t)n + r] = t[1]; rt[2 * (size_t)n + r] = t[2]; rt[3 * (size_t)n + r] = t[3];
        fixed[(size_t)FX_TAG_COMP * n + r] = small[t[1] & 15];
        fixed[(size_t)FX_TAG_OVER * n + r] = small[t[3] & 15];
        fixed[(size_t)FX_S_COMP * n + r] = small[t[0] & 1];
        fixed[(size_t)FX_S_OVER * n + r] = small[t[2] & 1];
    }
    for (uint32_t i = 0; i < pk->T; i++) {
        fixed[(size_t)FX_T_TAG * n + i] = small[table[i] >> 16];
        fixed[(size_t)FX_T_VALUE * n + i] = fr_from_u64(table[i] & 0xffffu);
    }
    // constants
    fe_t delta = fr_delta();
    pk->delta_pows[0] = Fr::one();
    for (int i = 1; i < NPERM; i++) pk->delta_pows[i] = Fr::mul(pk->delta_pows[i - 1], delta);
    {
        fe_t zn = fr_pow_host(fr_zeta(), n), wn = fr_pow_host(fr_omega(pk->ext_k), n), cur = Fr::one();
        for (int i = 0; i < 4; i++) {
            pk->t_inv[i] = Fr::inv(Fr::sub(Fr::mul(zn, cur), Fr::one()));
            cur = Fr::mul(cur, wn);
        }
    }
    DevConsts C = make_consts(pk);
    // uploads
    KG_CUDA(cudaMalloc(&pk->fixed_values, (size_t)NFIXED * n * 32));
    KG_CUDA(cudaMalloc(&pk->fixed_polys, (size_t)NFIXED * n * 32));
    KG_CUDA(cudaMalloc(&pk->fixed_cosets, (size_t)NFIXED * ext_n * 32));
    KG_CUDA(cudaMalloc(&pk->sigma_values, (size_t)NPERM * n * 32));
    KG_CUDA(cudaMalloc(&pk->sigma_polys, (size_t)NPERM * n * 32));
    KG_CUDA(cudaMalloc(&pk->sigma_cosets, (size_t)NPERM * ext_n * 32));
    KG_CUDA(cudaMalloc(&pk->l_cosets, (size_t)3 * ext_n * 32));
    KG_CUDA(cudaMalloc(&pk->range_tags, (size_t)4 * n));
    KG_CUDA(cudaMalloc(&pk->table, (size_t)pk->T * 4));
    KG_CUDA(cudaMemcpyAsync(pk->fixed_values, fixed.data(), fixed.size() * 32, cudaMemcpyHostToDevice, st));
    KG_CUDA(cudaMemcpyAsync(pk->range_tags, rt.data(), rt.size(), cudaMemcpyHostToDevice, st));
    KG_CUDA(cudaMemcpyAsync(pk->table, table.data(), table.size() * 4, cudaMemcpyHostToDevice, st));
    // permutation
    std::vector<uint32_t> mapping;
    build_permutation(n, prog->copies, mapping);
    uint32_t* d_map = nullptr;
    KG_TRY(scratch_get(ctx, SC_MISC, mapping.size() * 4 + (size_t)3 * n * 32 + 64 * sizeof(affine_t), (void**)&d_map));
    KG_CUDA(cudaMemcpyAsync(d_map, mapping.data(), mapping.size() * 4, cudaMemcpyHostToDevice, st));
    const fe_t* tw = nullptr;
    KG_TRY(ntt_get_twiddles(ctx, fr_omega(pk->k), pk->k, &tw));
    k_sigma<<<(NPERM * n + 255) / 256, 256, 0, st>>>(d_map, pk->sigma_values, tw, n, C);
    ctx->launches++;
    // polys and cosets
    KG_CUDA(cudaMemcpyAsync(pk->fixed_polys, pk->fixed_values, (size_t)NFIXED * n * 32, cudaMemcpyDeviceToDevice, st));
    KG_CUDA(cudaMemcpyAsync(pk->sigma_polys, pk->sigma_values, (size_t)NPERM * n * 32, cudaMemcpyDeviceToDevice, st));
    KG_TRY(b2r_intt_fr_batch_dev(ctx, (b2r_fr*)pk->fixed_polys, NFIXED, pk->k));
    KG_TRY(b2r_intt_fr_batch_dev(ctx, (b2r_fr*)pk->sigma_polys, NPERM, pk->k));
    KG_TRY(b2r_coset_ntt_fr_batch_dev(ctx, (b2r_fr*)pk->fixed_polys, NFIXED, pk->k, pk->ext_k, (b2r_fr*)pk->fixed_cosets));
    KG_TRY(b2r_coset_ntt_fr_batch_dev(ctx, (b2r_fr*)pk->sigma_polys, NPERM, pk->k, pk->ext_k, (b2r_fr*)pk->sigma_cosets));
    // l0, l_last, l_blind -> cosets; l_active = 1 - l_last - l_blind
    {
        std::vector<fe_t> lv((size_t)3 * n, Fr::zero());
        lv[0] = Fr::one();                                        // l0
        lv[(size_t)n + (n - BF - 1)] = Fr::one();                 // l_last
        for (uint32_t i = n - BF; i < n; i++) lv[2 * (size_t)n + i] = Fr::one();  // l_blind
        fe_t* d_l = (fe_t*)((char*)d_map + ((mapping.size() * 4 + 255) & ~(size_t)255));
        KG_CUDA(cudaMemcpyAsync(d_l, lv.data(), lv.size() * 32, cudaMemcpyHostToDevice, st));
        KG_TRY(b2r_intt_fr_batch_dev(ctx, (b2r_fr*)d_l, 3, pk->k));
        fe_t* d_ext = nullptr;
        KG_TRY(scratch_get(ctx, SC_STAGE, (size_t)3 * ext_n * 32, (void**)&d_ext));
        KG_TRY(b2r_coset_ntt_fr_batch_dev(ctx, (b2r_fr*)d_l, 3, pk->k, pk->ext_k, (b2r_fr*)d_ext));
        KG_CUDA(cudaMemcpyAsync(pk->l_cosets, d_ext, (size_t)2 * ext_n * 32, cudaMemcpyDeviceToDevice, st));  // l0, l_last
        KG_CUDA(cudaMemcpyAsync(pk->l_cosets + 2 * (size_t)ext_n, d_ext + (size_t)ext_n, (size_t)ext_n * 32, cudaMemcpyDeviceToDevice, st));
        k_l_active<<<(ext_n + 255) / 256, 256, 0, st>>>(pk->l_cosets + 2 * (size_t)ext_n, d_ext + 2 * (size_t)ext_n, ext_n);
        ctx->launches++;
        KG_CUDA(cudaStreamSynchronize(st));  // lv goes out of scope
    }
    KG_TRY(bases_register_suffix_sums(ctx, g_lagrange, &pk->gl_sfx));
    KG_TRY(bases_register_rewindowed(ctx, g_lagrange, 13, &pk->gl_c13));
    KG_TRY(bases_register_rewindowed(ctx, g_lagrange, 10, &pk->gl_c10));
    KG_TRY(bases_register_rewindowed(ctx, pk->gl_sfx, 10, &pk->gl_sfx_c10));
    // commitments of the verifying key
    {
        affine_t* d_cm = nullptr;
        KG_TRY(scratch_get(ctx, SC_STAGE, 64 * sizeof(affine_t), (void**)&d_cm));
        KG_TRY(msm_batch_dev(ctx, g_lagrange, pk->fixed_values, NFIXED, n, d_cm, false));
        KG_TRY(msm_batch_dev(ctx, g_lagrange, pk->sigma_values, NPERM, n, d_cm + NFIXED, false));
        pk->fixed_commitments.resize(NFIXED);
        pk->sigma_commitments.resize(NPERM);
        KG_CUDA(cudaMemcpyAsync(pk->fixed_commitments.data(), d_cm, NFIXED * sizeof(affine_t), cudaMemcpyDeviceToHost, st));
        KG_CUDA(cudaMemcpyAsync(pk->sigma_commitments.data(), d_cm + NFIXED, NPERM * sizeof(affine_t), cudaMemcpyDeviceToHost, st));
        KG_CUDA(cudaStreamSynchronize(st));
    }
    {
        Blake2b h;
        h.init("Halo2-Verify-Key");
        uint8_t kb[4] = {(uint8_t)pk->k, (uint8_t)(pk->k >> 8), (uint8_t)(pk->k >> 16), (uint8_t)(pk->k >> 24)};
        h.update(kb, 4);
        uint8_t b[32];
        for (const auto& pt : pk->fixed_commitments) { compress_point_bytes(pt, b); h.update(b, 32); }
        for (const auto& pt : pk->sigma_commitments) { compress_point_bytes(pt, b); h.update(b, 32); }
        uint8_t d[64];
        h.digest(d);
        pk->transcript_repr = fr_from_wide(d);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return bail(cuda_fail(ctx, e, "keygen kernels"));
    *out = pk;
    return 0;
#undef KG_CUDA
#undef KG_TRY
} B2R_ABI_CATCH(ctx)

int32_t b2r_pk_free(b2r_ctx* ctx, b2r_pk* pk) try {
    B2R_ENTER(ctx);
    if (!ctx || !pk) return B2R_ERR_INVALID;
    cudaStreamSynchronize(ctx->stream);
    pk_release(pk);
    return 0;
} B2R_ABI_CATCH(ctx)

int32_t b2r_pk_info(const b2r_pk* pk, uint32_t* k, uint32_t* ext_k, uint32_t* num_fixed, uint32_t* num_sigma, uint64_t* proof_bytes) {
    if (!pk) return B2R_ERR_INVALID;
    if (k) *k = pk->k;
    if (ext_k) *ext_k = pk->ext_k;
    if (num_fixed) *num_fixed = NFIXED;
    if (num_sigma) *num_sigma = NPERM;
    if (proof_bytes) *proof_bytes = 32 * (NADV + 2 * NLOOK + NSETS + NLOOK + 1 + QD + NPOINTS + NEVAL);
    return 0;
}

int32_t b2r_pk_set_transcript_repr(b2r_pk* pk, const b2r_fr* transcript_repr) {
    if (!pk || !transcript_repr) return B2R_ERR_INVALID;
    // must be a reduced Montgomery representation (what a Rust Fr holds in memory)
    fe_t v;
    for (int i = 0; i < 4; i++) {
        v.l[2 * i] = (uint32_t)transcript_repr->l[i];
        v.l[2 * i + 1] = (uint32_t)(transcript_repr->l[i] >> 32);
    }
    uint32_t m[8], t[8];
    for (int i = 0; i < 8; i++) m[i] = FrP::MOD(i);
    if (!sub8(t, v.l, m)) return B2R_ERR_INVALID;  // >= r
    pk->transcript_repr = v;
    return 0;
}

int32_t b2r_pk_export_vk(const b2r_pk* pk, b2r_g1_affine* fixed_commitments, b2r_g1_affine* sigma_commitments, b2r_fr* transcript_repr) {
    if (!pk || !fixed_commitments || !sigma_commitments || !transcript_repr) return B2R_ERR_INVALID;
    memcpy(fixed_commitments, pk->fixed_commitments.data(), NFIXED * sizeof(affine_t));
    memcpy(sigma_commitments, pk->sigma_commitments.data(), NPERM * sizeof(affine_t));
    for (int i = 0; i < 4; i++) transcript_repr->l[i] = (uint64_t)pk->transcript_repr.l[2 * i] | ((uint64_t)pk->transcript_repr.l[2 * i + 1] << 32);
    return 0;
}

}  // extern "C"

// ---- the prover ---------------------------------------------------------------------------------------------------
namespace b2r {

struct ProveScratch {
    fe_t *P, *hbuf, *num, *den, *pref, *chunk_prod, *E, *hext, *lc, *wq, *sorted_cv, *chal, *ypow, *points, *scal, *evals;
    uint32_t *hist, *order, *err;
    LookupPlan* plans;
    affine_t* cm;
    EvalPlan* eplan;
    LincombPlan* lplan;
    uint64_t* inputs;
    uint8_t* valid;
};

static void build_plans(EvalPlan& ep, LincombPlan& lp) {
    int e = 0;
    auto add_eval = [&](uint32_t kind, uint32_t idx, uint8_t point) { ep.poly[e] = {kind, idx}; ep.point[e] = point; e++; };
    // points: 0 = x, 1 = omega x, 2 = omega^last x, 3 = omega^-1 x (order of first appearance among the queries)
    for (int c = 0; c < NADV; c++) add_eval(0, SL_ADV + c, 0);
    add_eval(0, SL_ADV + 4, 1);
    for (int c = 0; c < NFIXED; c++) add_eval(1, c, 0);
    add_eval(0, SL_RAND, 0);
    for (int c = 0; c < NPERM; c++) add_eval(1, NFIXED + c, 0);
    for (int s = 0; s < NSETS; s++) {
        add_eval(0, SL_PZ + s, 0);
        add_eval(0, SL_PZ + s, 1);
        if (s != NSETS - 1) add_eval(0, SL_PZ + s, 2);
    }
    for (int l = 0; l < NLOOK; l++) {
        add_eval(0, SL_LZ + l, 0);
        add_eval(0, SL_LZ + l, 1);
        add_eval(0, SL_LA + 2 * l, 0);
        add_eval(0, SL_LA + 2 * l, 3);
        add_eval(0, SL_LA + 2 * l + 1, 0);
    }
    // GWC groups in query order
    for (int g = 0; g < NPOINTS; g++) lp.nterms[g] = 0;
    auto add_q = [&](int g, uint32_t kind, uint32_t idx) { lp.poly[g][lp.nterms[g]++] = {kind, idx}; };
    for (int c = 0; c < NADV; c++) add_q(0, 0, SL_ADV + c);
    add_q(1, 0, SL_ADV + 4);
    for (int s = 0; s < NSETS; s++) { add_q(0, 0, SL_PZ + s); add_q(1, 0, SL_PZ + s); }
    for (int s = NSETS - 2; s >= 0; s--) add_q(2, 0, SL_PZ + s);
    for (int l = 0; l < NLOOK; l++) {
        add_q(0, 0, SL_LZ + l); add_q(0, 0, SL_LA + 2 * l); add_q(0, 0, SL_LA + 2 * l + 1); add_q(3, 0, SL_LA + 2 * l); add_q(1, 0, SL_LZ + l);
    }
    for (int c = 0; c < NFIXED; c++) add_q(0, 1, c);
    for (int c = 0; c < NPERM; c++) add_q(0, 1, NFIXED + c);
    for (int j = 0; j < QD; j++) add_q(0, 2, j);  // h(X) = sum_j x^(n j) piece_j: QD terms sharing one power of v
    add_q(0, 0, SL_RAND);
}

// The per-proof host work between the GPU phases (Blake2b transcripts, Montgomery conversions, challenge powers) is
// independent across proofs: spread it over a few host threads so that the GPU does not idle behind one core
// (about 0.1 ms of host field arithmetic per proof and step).
template <class F>
static void parallel_for_proofs(uint32_t count, F body) {
    uint32_t nt = std::min<uint32_t>(8, std::max<uint32_t>(1, std::thread::hardware_concurrency()));
    nt = std::min<uint32_t>(nt, std::max<uint32_t>(1, count / 4));
    if (nt <= 1) {
        for (uint32_t p = 0; p < count; p++) body(p);
        return;
    }
    std::vector<std::thread> th;
    auto range = [&](uint32_t t) {
        const uint32_t lo = (uint64_t)count * t / nt, hi = (uint64_t)count * (t + 1) / nt;
        for (uint32_t p = lo; p < hi; p++) body(p);
    };
    for (uint32_t t = 1; t < nt; t++) th.emplace_back(range, t);
    range(0);
    for (auto& x : th) x.join();
}

// one group of at most `B` proofs, inputs already on the device
static int32_t prove_group(b2r_ctx* ctx, const b2r_pk* pk, const uint64_t* d_n, const uint64_t* d_s, const uint64_t* d_h, uint32_t B,
                           const BlindKey& seed, uint32_t p_base, uint8_t* proofs_host, uint8_t* status_host, size_t proof_bytes) {
    const uint32_t n = pk->n, ext_n = pk->ext_n, u = pk->u, k = pk->k, T = pk->T;
    const uint32_t QB = std::min<uint32_t>(B, 16);
    const uint32_t nch = n / CH;
    cudaStream_t st = ctx->stream;
    const DevConsts C = make_consts(pk);
    // ---- arena
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t r = off; off += (bytes + 255) & ~(size_t)255; return r; };
    const size_t oP = carve((size_t)NSLOT * B * n * 32), oH = carve((size_t)B * QD * n * 32);
    const size_t oNum = carve((size_t)NZ * B * n * 32), oDen = carve((size_t)NZ * B * n * 32), oPref = carve((size_t)NZ * B * n * 32);
    const size_t oCp = carve((size_t)NZ * B * nch * 32);
    const size_t oE = carve((size_t)NTRANS * QB * ext_n * 32), oHext = carve((size_t)QB * ext_n * 32);
    const size_t oLc = carve((size_t)B * NPOINTS * n * 32), oWq = carve((size_t)B * NPOINTS * n * 32);
    const size_t oCv = carve((size_t)B * MAX_TABLE * 32), oChal = carve((size_t)B * 8 * 32), oYp = carve((size_t)B * NCONS * 32), oPts = carve((size_t)B * NPOINTS * 32);
    const size_t oScal = carve((size_t)B * NPOINTS * MAXTERMS * 32), oEv = carve((size_t)B * NEVAL * 32);
    const size_t oHist = carve((size_t)B * NLOOK * MAX_TABLE * 4), oOrd = carve((size_t)B * MAX_TABLE * 4), oErr = carve((size_t)B * 4);
    const size_t oPlans = carve((size_t)B * NLOOK * sizeof(LookupPlan));
    const size_t oCm = carve((size_t)B * 16 * sizeof(affine_t));
    const size_t oEp = carve(sizeof(EvalPlan)), oLp = carve(sizeof(LincombPlan));
    const size_t oValid = carve(B);
    char* base = nullptr;
    B2R_TRY(scratch_get(ctx, SC_MSM_C, off, (void**)&base));
    ProveScratch S;
    S.P = (fe_t*)(base + oP); S.hbuf = (fe_t*)(base + oH); S.num = (fe_t*)(base + oNum); S.den = (fe_t*)(base + oDen);
    S.pref = (fe_t*)(base + oPref); S.chunk_prod = (fe_t*)(base + oCp); S.E = (fe_t*)(base + oE); S.hext = (fe_t*)(base + oHext);
    S.lc = (fe_t*)(base + oLc); S.wq = (fe_t*)(base + oWq); S.sorted_cv = (fe_t*)(base + oCv); S.chal = (fe_t*)(base + oChal); S.ypow = (fe_t*)(base + oYp);
    S.points = (fe_t*)(base + oPts); S.scal = (fe_t*)(base + oScal); S.evals = (fe_t*)(base + oEv);
    S.hist = (uint32_t*)(base + oHist); S.order = (uint32_t*)(base + oOrd); S.err = (uint32_t*)(base + oErr);
    S.plans = (LookupPlan*)(base + oPlans); S.cm = (affine_t*)(base + oCm); S.eplan = (EvalPlan*)(base + oEp);
    S.lplan = (LincombPlan*)(base + oLp); S.valid = (uint8_t*)(base + oValid);

    EvalPlan ep;
    LincombPlan lp;
    memset(&ep, 0, sizeof ep);
    memset(&lp, 0, sizeof lp);
    build_plans(ep, lp);
    B2R_CUDA(ctx, cudaMemcpyAsync(S.eplan, &ep, sizeof ep, cudaMemcpyHostToDevice, st));
    B2R_CUDA(ctx, cudaMemcpyAsync(S.lplan, &lp, sizeof lp, cudaMemcpyHostToDevice, st));
    B2R_CUDA(ctx, cudaMemsetAsync(S.hist, 0, (size_t)B * NLOOK * MAX_TABLE * 4, st));
    B2R_CUDA(ctx, cudaMemsetAsync(S.err, 0, (size_t)B * 4, st));

    std::vector<Transcript> tr(B);
    std::vector<fe_t> chal((size_t)B * 8, Fr::zero()), ypow((size_t)B * NCONS);
    std::vector<affine_t> cm((size_t)B * 16);
    std::vector<uint8_t> valid(B);
    std::vector<uint32_t> err(B);
    auto fetch_points = [&](size_t count) -> int32_t {
        B2R_CUDA(ctx, cudaMemcpyAsync(cm.data(), S.cm, count * sizeof(affine_t), cudaMemcpyDeviceToHost, st));
        B2R_CUDA(ctx, cudaStreamSynchronize(st));
        return 0;
    };
    // every phase also files its commitments in the device-resident block (no host round trip: a copy kernel on the stream)
    auto log_points = [&](uint32_t J, uint32_t off, uint32_t mode) -> int32_t {
        if (!ctx->commit_log) return 0;
        k_log_commitments<<<(B * J + 255) / 256, 256, 0, st>>>(S.cm, (affine_t*)ctx->commit_log, B, J, off, p_base, mode);
        B2R_LAUNCH_CHECK(ctx);
        return 0;
    };
    auto push_chal = [&]() -> int32_t {
        B2R_CUDA(ctx, cudaMemcpyAsync(S.chal, chal.data(), chal.size() * 32, cudaMemcpyHostToDevice, st));
        return 0;
    };
    const fe_t* tw = nullptr;
    const fe_t* tw_ext = nullptr;
    B2R_TRY(ntt_get_twiddles(ctx, fr_omega(k), k, &tw));
    B2R_TRY(ntt_get_twiddles(ctx, fr_omega(pk->ext_k), pk->ext_k, &tw_ext));

    // ---- phase 1: witness + advice commitments
    B2R_TRY(witness_run(ctx, pk->prog, d_n, d_s, d_h, B, seed, (b2r_fr*)S.P, S.valid, p_base, /*p_stride=*/n, /*col_stride=*/(size_t)B * n));
    B2R_TRY(msm_batch_dev(ctx, pk->gl_c13, S.P + (size_t)SL_ADV * B * n, (size_t)NADV * B, n, S.cm, false));
    B2R_CUDA(ctx, cudaMemcpyAsync(valid.data(), S.valid, B, cudaMemcpyDeviceToHost, st));
    B2R_TRY(log_points(NADV, 0, 0));
    B2R_TRY(fetch_points((size_t)NADV * B));
    parallel_for_proofs(B, [&](uint32_t p) {
        tr[p].common_scalar(pk->transcript_repr);
        for (int c = 0; c < NADV; c++) tr[p].write_point(cm[(size_t)c * B + p].x, cm[(size_t)c * B + p].y);
        chal[(size_t)p * 8 + 0] = tr[p].squeeze();  // theta
    });
    B2R_TRY(push_chal());
    // ---- phase 2: lookups
    {
        KTimer kt(ctx, "lookup_permute", (double)B);
        k_lookup_hist<<<dim3((u + 255) / 256, NLOOK, B), 256, 0, st>>>(S.P, pk->range_tags, n, u, B, T, S.hist, S.err, C);
        B2R_LAUNCH_CHECK(ctx);
        k_table_sort<<<B, 1024, 0, st>>>(pk->table, T, S.chal, S.sorted_cv, S.order);
        B2R_LAUNCH_CHECK(ctx);
        k_lookup_plan<<<dim3(NLOOK, B), 1024, 0, st>>>(S.hist, S.order, T, u, S.plans, S.err);
        B2R_LAUNCH_CHECK(ctx);
        k_lookup_fill<<<dim3((n + 255) / 256, NLOOK, B), 256, 0, st>>>(S.P, S.plans, S.sorted_cv, T, n, u, B, seed, p_base);
        B2R_LAUNCH_CHECK(ctx);
    }
    // A' is sorted, i.e. constant over at most T + 1 runs: commit it through its first differences against the suffix-sum
    // bases like the grand products (a few hundred non-zero scalars instead of every looked-up row); S' is non-zero only
    // at run starts and is committed directly.  S.num .. S.den (contiguous, free until phase 3) hold D(A'_l) then S'_l.
    {
        fe_t* DA = S.num;
        fe_t* SS = S.num + (size_t)NLOOK * B * n;
        if (S.den != S.num + (size_t)NZ * B * n) return fail(ctx, B2R_ERR_INVALID, "prove: scratch layout");
        for (int l = 0; l < NLOOK; l++) {
            k_run_diff<<<dim3((n + 255) / 256, B), 256, 0, st>>>(S.P + (size_t)(SL_LA + 2 * l) * B * n, DA + (size_t)l * B * n, n);
            B2R_LAUNCH_CHECK(ctx);
        }
        B2R_CUDA(ctx, cudaMemcpy2DAsync(SS, (size_t)B * n * 32, S.P + (size_t)(SL_LA + 1) * B * n, (size_t)2 * B * n * 32, (size_t)B * n * 32, NLOOK,
                                        cudaMemcpyDeviceToDevice, st));
        B2R_TRY(msm_batch_dev(ctx, pk->gl_sfx_c10, DA, (size_t)NLOOK * B, n, S.cm, false));
        B2R_TRY(msm_batch_dev(ctx, pk->gl_c10, SS, (size_t)NLOOK * B, n, S.cm + (size_t)NLOOK * B, false));
    }
    B2R_CUDA(ctx, cudaMemcpyAsync(err.data(), S.err, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    B2R_TRY(log_points(2 * NLOOK, NADV, 2));
    B2R_TRY(fetch_points((size_t)2 * NLOOK * B));
    parallel_for_proofs(B, [&](uint32_t p) {
        for (int j = 0; j < 2 * NLOOK; j++) {   // halo2 writes A'_l, S'_l per lookup
            const affine_t& c = cm[((size_t)(j & 1) * NLOOK + (j >> 1)) * B + p];
            tr[p].write_point(c.x, c.y);
        }
        chal[(size_t)p * 8 + 1] = tr[p].squeeze();  // beta
        chal[(size_t)p * 8 + 2] = tr[p].squeeze();  // gamma
    });
    B2R_TRY(push_chal());
    // ---- phase 3: grand products + random polynomial
    {
        KTimer kt(ctx, "grand_products", (double)B);
        k_perm_numden<<<dim3((n + 255) / 256, NSETS, B), 256, 0, st>>>(S.P, pk->sigma_values, tw, S.chal, n, B, S.num, S.den, C);
        B2R_LAUNCH_CHECK(ctx);
        k_lookup_numden<<<dim3((n + 255) / 256, NLOOK, B), 256, 0, st>>>(S.P, pk->fixed_values, S.chal, n, B, S.num, S.den);
        B2R_LAUNCH_CHECK(ctx);
        const uint32_t nzt = NZ * B, threads = nzt * nch;
        k_batch_inv_ratio<<<(threads + 127) / 128, 128, 0, st>>>(S.num, S.den, S.pref, n, u, nzt, S.chunk_prod);
        B2R_LAUNCH_CHECK(ctx);
        k_chunk_scan<<<nzt, 1024, 1024 * 32, st>>>(S.chunk_prod, nch);
        B2R_LAUNCH_CHECK(ctx);
        k_z_write<<<(threads + 127) / 128, 128, 0, st>>>(S.P, S.den, S.chunk_prod, n, B, nzt, seed, p_base);
        B2R_LAUNCH_CHECK(ctx);
        for (int s = 1; s < NSETS; s++) {
            k_perm_chain<<<dim3((u + 256) / 256, B), 256, 0, st>>>(S.P, n, u, B, s);
            B2R_LAUNCH_CHECK(ctx);
        }
        k_random_poly<<<dim3((n + 255) / 256, B), 256, 0, st>>>(S.P, n, B, seed, p_base);
        B2R_LAUNCH_CHECK(ctx);
    }
    // Z columns through their run structure (first differences against the suffix-sum bases); S.num is free again
    k_run_diff<<<dim3((n + 255) / 256, NZ * B), 256, 0, st>>>(S.P + (size_t)SL_PZ * B * n, S.num, n);
    B2R_LAUNCH_CHECK(ctx);
    B2R_TRY(msm_batch_dev(ctx, pk->gl_sfx, S.num, (size_t)NZ * B, n, S.cm, true));   // non-zero differences of a grand product: uniform
    B2R_TRY(msm_batch_dev(ctx, pk->g, S.P + (size_t)SL_RAND * B * n, B, n, S.cm + (size_t)NZ * B, true));
    B2R_TRY(log_points(NZ + 1, NADV + 2 * NLOOK, 0));
    B2R_TRY(fetch_points((size_t)(NZ + 1) * B));
    parallel_for_proofs(B, [&](uint32_t p) {
        for (int j = 0; j < NZ + 1; j++) tr[p].write_point(cm[(size_t)j * B + p].x, cm[(size_t)j * B + p].y);
        chal[(size_t)p * 8 + 3] = tr[p].squeeze();  // y
        fe_t* yp = ypow.data() + (size_t)p * NCONS;   // y^(NCONS-1-k): the weight of constraint k in h
        yp[NCONS - 1] = Fr::one();
        for (int k = NCONS - 2; k >= 0; k--) yp[k] = Fr::mul(yp[k + 1], chal[(size_t)p * 8 + 3]);
    });
    B2R_TRY(push_chal());
    B2R_CUDA(ctx, cudaMemcpyAsync(S.ypow, ypow.data(), ypow.size() * 32, cudaMemcpyHostToDevice, st));
    // ---- phase 4: coefficient forms, extended coset, quotient
    B2R_TRY(b2r_intt_fr_batch_dev(ctx, (b2r_fr*)S.P, (size_t)NTRANS * B, k));
    for (uint32_t q0 = 0; q0 < B; q0 += QB) {
        const uint32_t qb = std::min(QB, B - q0);
        // all 22 columns of the sub-batch in one launch per pass (slot-major arena: slot s of proof q0 + j at (s * B + q0 + j) * n)
        if (qb == QB) {
            B2R_TRY(coset_ntt_grouped_dev(ctx, S.P + (size_t)q0 * n, NTRANS, (uint64_t)B * n, qb, k, pk->ext_k, S.E));
        } else {
            for (int s = 0; s < NTRANS; s++)
                B2R_TRY(b2r_coset_ntt_fr_batch_dev(ctx, (b2r_fr*)(S.P + ((size_t)s * B + q0) * n), qb, k, pk->ext_k, (b2r_fr*)(S.E + (size_t)s * QB * ext_n)));
        }
        QuotArgs A;
        A.E = S.E; A.fixed_c = pk->fixed_cosets; A.sigma_c = pk->sigma_cosets; A.l_c = pk->l_cosets; A.tw_ext = tw_ext;
        A.chal = S.chal + (size_t)q0 * 8; A.ypow = S.ypow + (size_t)q0 * NCONS; A.h = S.hext; A.ext_n = ext_n; A.step = ext_n / n; A.QB = QB;
        { KTimer kt(ctx, "quotient", (double)qb);
        k_quotient<<<dim3(ext_n / 128, qb), 128, 0, st>>>(A, C); }
        B2R_LAUNCH_CHECK(ctx);
        B2R_TRY(b2r_coset_intt_fr_batch_dev(ctx, (b2r_fr*)S.hext, qb, pk->ext_k));
        B2R_CUDA(ctx, cudaMemcpy2DAsync(S.hbuf + (size_t)q0 * QD * n, (size_t)QD * n * 32, S.hext, (size_t)ext_n * 32, (size_t)QD * n * 32, qb,
                                        cudaMemcpyDeviceToDevice, st));
    }
    B2R_TRY(msm_batch_dev(ctx, pk->g, S.hbuf, (size_t)QD * B, n, S.cm, true));
    B2R_TRY(log_points(QD, NADV + 2 * NLOOK + NZ + 1, 1));
    B2R_TRY(fetch_points((size_t)QD * B));
    std::vector<fe_t> points((size_t)B * NPOINTS), scal((size_t)B * NPOINTS * MAXTERMS, Fr::zero()), xn(B);
    {
        const fe_t omega = fr_omega(k), omega_inv = Fr::inv(omega);
        const fe_t w_last = fr_pow_host(omega_inv, BF + 1);
        parallel_for_proofs(B, [&](uint32_t p) {
            for (int j = 0; j < QD; j++) tr[p].write_point(cm[(size_t)p * QD + j].x, cm[(size_t)p * QD + j].y);
            const fe_t x = tr[p].squeeze();
            chal[(size_t)p * 8 + 4] = x;
            points[(size_t)p * NPOINTS + 0] = x;
            points[(size_t)p * NPOINTS + 1] = Fr::mul(x, omega);
            points[(size_t)p * NPOINTS + 2] = Fr::mul(x, w_last);
            points[(size_t)p * NPOINTS + 3] = Fr::mul(x, omega_inv);
            xn[p] = fr_pow_host(x, n);
        });
    }
    B2R_CUDA(ctx, cudaMemcpyAsync(S.points, points.data(), points.size() * 32, cudaMemcpyHostToDevice, st));
    // ---- phase 5: evaluations
    PolyTable PT;
    PT.P = S.P; PT.fixedp = pk->fixed_polys; PT.sigmap = pk->sigma_polys; PT.hbuf = S.hbuf; PT.n = n; PT.B = B;
    { KTimer kt(ctx, "evaluate", (double)B);
    k_eval<<<dim3(NEVAL, B), 256, 0, st>>>(PT, S.eplan, S.points, S.evals); }
    B2R_LAUNCH_CHECK(ctx);
    std::vector<fe_t> evals((size_t)B * NEVAL);
    B2R_CUDA(ctx, cudaMemcpyAsync(evals.data(), S.evals, evals.size() * 32, cudaMemcpyDeviceToHost, st));
    B2R_CUDA(ctx, cudaStreamSynchronize(st));
    parallel_for_proofs(B, [&](uint32_t p) {
        for (int e = 0; e < NEVAL; e++) tr[p].write_scalar(evals[(size_t)p * NEVAL + e]);
        const fe_t v = tr[p].squeeze();
        // scalars of the GWC linear combinations: term j of m gets v^(m-1-j); the QD h pieces share one v power
        for (int g = 0; g < NPOINTS; g++) {
            const uint32_t mt = lp.nterms[g];
            // number of distinct "queries": h pieces count once
            uint32_t nq = 0;
            for (uint32_t j = 0; j < mt; j++) nq += (lp.poly[g][j].kind == 2 && lp.poly[g][j].idx != 0) ? 0 : 1;
            fe_t vp = Fr::one();
            std::vector<fe_t> pw(nq);
            for (uint32_t j = 0; j < nq; j++) { pw[nq - 1 - j] = vp; vp = Fr::mul(vp, v); }
            uint32_t qi = 0;
            fe_t xnp = Fr::one();
            for (uint32_t j = 0; j < mt; j++) {
                fe_t s;
                if (lp.poly[g][j].kind == 2) {
                    if (lp.poly[g][j].idx == 0) { xnp = Fr::one(); s = pw[qi]; }
                    else { xnp = Fr::mul(xnp, xn[p]); s = Fr::mul(pw[qi], xnp); }
                    if (lp.poly[g][j].idx == QD - 1) qi++;
                } else {
                    s = pw[qi++];
                }
                scal[((size_t)p * NPOINTS + g) * MAXTERMS + j] = s;
            }
        }
    });
    B2R_CUDA(ctx, cudaMemcpyAsync(S.scal, scal.data(), scal.size() * 32, cudaMemcpyHostToDevice, st));
    // ---- phase 6: GWC witnesses
    { KTimer kt(ctx, "multiopen", (double)B);
    k_lincomb<<<dim3((n + 255) / 256, NPOINTS, B), 256, 0, st>>>(PT, S.lplan, S.scal, S.lc);
    B2R_LAUNCH_CHECK(ctx);
    k_kate<<<dim3(NPOINTS, B), 256, 0, st>>>(S.lc, S.points, n, S.wq); }
    B2R_LAUNCH_CHECK(ctx);
    B2R_TRY(msm_batch_dev(ctx, pk->g, S.wq, (size_t)NPOINTS * B, n, S.cm, true));
    B2R_TRY(log_points(NPOINTS, NADV + 2 * NLOOK + NZ + 1 + QD, 1));
    B2R_TRY(fetch_points((size_t)NPOINTS * B));
    for (uint32_t p = 0; p < B; p++) {
        for (int g = 0; g < NPOINTS; g++) tr[p].write_point(cm[(size_t)p * NPOINTS + g].x, cm[(size_t)p * NPOINTS + g].y);
        if (tr[p].out.size() != proof_bytes) return fail(ctx, B2R_ERR_INVALID, "prove: internal proof size mismatch");
        memcpy(proofs_host + (size_t)p * proof_bytes, tr[p].out.data(), proof_bytes);
        // status: 1 = proof of a valid signature, 0 = witness does not satisfy the circuit (is_valid = 0; the proof will not
        // verify), 0xFF = the reference's synthesize would have panicked, 0xFE = a lookup input is not in the table
        uint8_t stt = valid[p];
        if (stt != 0xFF && err[p]) stt = 0xFE;
        status_host[p] = stt;
    }
    return 0;
}

}  // namespace b2r

extern "C" {

static int32_t prove_all(b2r_ctx* ctx, const b2r_pk* pk, const uint64_t* d_n, const uint64_t* d_s, const uint64_t* d_h, size_t batch,
                         const BlindKey& bkey, uint8_t* proofs, uint8_t* status) {
    uint64_t proof_bytes = 0;
    b2r_pk_info(pk, nullptr, nullptr, nullptr, nullptr, &proof_bytes);
    const size_t nl = pk->prog->num_limbs;
    // group size bounded by a memory budget (about 230 MiB of arena per proof at k = 17)
    const size_t per_proof = ((size_t)NSLOT + QD + 3 * NZ + 2 * NPOINTS) * pk->n * 32 + 4096;
    size_t G = std::max<size_t>(1, std::min<size_t>(64, ((size_t)48 << 30) / per_proof));
    if (const char* ov = getenv("B2R_PROVE_GROUP")) G = std::max<size_t>(1, std::min<size_t>(G, (size_t)atoi(ov)));  // tests: force several groups
    // device-resident commitment block of this call (b2r_last_commitments)
    ctx->commit_log_batch = 0;
    if (ctx->commit_log_cap < batch) {
        if (ctx->commit_log) {
            B2R_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            B2R_CUDA(ctx, cudaFree(ctx->commit_log));
            ctx->commit_log = nullptr;
            ctx->commit_log_cap = 0;
        }
        B2R_CUDA(ctx, cudaMalloc(&ctx->commit_log, batch * NCOMMIT * sizeof(affine_t)));
        ctx->commit_log_cap = batch;
    }
    for (size_t p0 = 0; p0 < batch; p0 += G) {
        const uint32_t g = (uint32_t)std::min(G, batch - p0);
        B2R_TRY(prove_group(ctx, pk, d_n + p0 * nl, d_s + p0 * nl, d_h + p0 * pk->prog->aux_words, g, bkey, (uint32_t)p0, proofs + p0 * proof_bytes, status + p0,
                            (size_t)proof_bytes));
    }
    ctx->commit_log_batch = batch;
    return 0;
}

static int32_t prove_entry(b2r_ctx* ctx, const b2r_pk* pk, const uint64_t* n_limbs, const uint64_t* sig_limbs, const uint64_t* hash_limbs,
                           size_t batch, const BlindKey& bkey, bool inputs_on_device, uint8_t* proofs, uint8_t* status) {
    if (!pk || !n_limbs || !sig_limbs || !hash_limbs || !proofs || !status) return fail(ctx, B2R_ERR_INVALID, "rsa_prove: null pointer");
    if (batch == 0) return 0;
    if (batch > 0xffffffffull) return fail(ctx, B2R_ERR_INVALID, "rsa_prove: batch too large");
    if (inputs_on_device) return prove_all(ctx, pk, n_limbs, sig_limbs, hash_limbs, batch, bkey, proofs, status);
    const size_t nl = pk->prog->num_limbs;
    uint64_t* d_in = nullptr;
    const size_t aw = pk->prog->aux_words;
    B2R_TRY(scratch_get(ctx, SC_MISC, batch * (2 * nl + aw) * 8 + 256, (void**)&d_in));
    uint64_t *d_n = d_in, *d_s = d_in + batch * nl, *d_h = d_in + 2 * batch * nl;
    B2R_CUDA(ctx, cudaMemcpyAsync(d_n, n_limbs, batch * nl * 8, cudaMemcpyHostToDevice, ctx->stream));
    B2R_CUDA(ctx, cudaMemcpyAsync(d_s, sig_limbs, batch * nl * 8, cudaMemcpyHostToDevice, ctx->stream));
    B2R_CUDA(ctx, cudaMemcpyAsync(d_h, hash_limbs, batch * aw * 8, cudaMemcpyHostToDevice, ctx->stream));
    return prove_all(ctx, pk, d_n, d_s, d_h, batch, bkey, proofs, status);
}

// the 64-bit-seed entry points draw a fresh nonce from the context for every call, so that reusing a seed never
// repeats a blinding stream (two witnesses under identical blinds would leak their difference)
int32_t b2r_rsa_prove_batch_dev(b2r_ctx* ctx, const b2r_pk* pk, const uint64_t* n_limbs_dev, const uint64_t* sig_limbs_dev,
                                const uint64_t* hash_limbs_dev, size_t batch, uint64_t seed, uint8_t* proofs, uint8_t* status) try {
    B2R_ENTER(ctx);
    if (seed == 0) return fail(ctx, B2R_ERR_INVALID, "rsa_prove: seed must be non-zero (blinding rows)");
    return prove_entry(ctx, pk, n_limbs_dev, sig_limbs_dev, hash_limbs_dev, batch, blind_key_from_seed64(seed, ctx->prove_calls++), true, proofs, status);
} B2R_ABI_CATCH(ctx)

int32_t b2r_rsa_prove_batch(b2r_ctx* ctx, const b2r_pk* pk, const uint64_t* n_limbs, const uint64_t* sig_limbs, const uint64_t* hash_limbs,
                            size_t batch, uint64_t seed, uint8_t* proofs, uint8_t* status) try {
    B2R_ENTER(ctx);
    if (seed == 0) return fail(ctx, B2R_ERR_INVALID, "rsa_prove: seed must be non-zero (blinding rows)");
    return prove_entry(ctx, pk, n_limbs, sig_limbs, hash_limbs, batch, blind_key_from_seed64(seed, ctx->prove_calls++), false, proofs, status);
} B2R_ABI_CATCH(ctx)

int32_t b2r_rsa_prove_batch_ex(b2r_ctx* ctx, const b2r_pk* pk, const uint64_t* n_limbs, const uint64_t* sig_limbs, const uint64_t* hash_limbs,
                               size_t batch, const uint8_t seed32[32], uint64_t nonce, uint32_t flags, uint8_t* proofs, uint8_t* status) try {
    B2R_ENTER(ctx);
    if (!seed32) return fail(ctx, B2R_ERR_INVALID, "rsa_prove: null seed");
    if (flags & ~(uint32_t)(B2R_PROVE_INPUTS_ON_DEVICE | B2R_PROVE_SEED64)) return fail(ctx, B2R_ERR_INVALID, "rsa_prove: unknown flag");
    BlindKey bkey;
    if (flags & B2R_PROVE_SEED64) {
        uint64_t s64 = 0;
        for (int i = 0; i < 8; i++) s64 |= (uint64_t)seed32[i] << (8 * i);
        if (s64 == 0) return fail(ctx, B2R_ERR_INVALID, "rsa_prove: seed must be non-zero (blinding rows)");
        bkey = blind_key_from_seed64(s64, nonce);
    } else {
        bkey = blind_key_from_bytes(seed32, nonce);
    }
    return prove_entry(ctx, pk, n_limbs, sig_limbs, hash_limbs, batch, bkey, (flags & B2R_PROVE_INPUTS_ON_DEVICE) != 0, proofs, status);
} B2R_ABI_CATCH(ctx)

int32_t b2r_last_commitments(b2r_ctx* ctx, b2r_g1_affine* dst, size_t capacity_points, uint32_t dst_on_device, size_t* batch, uint32_t* per_proof) try {
    B2R_ENTER(ctx);
    if (batch) *batch = ctx->commit_log_batch;
    if (per_proof) *per_proof = NCOMMIT;
    if (!dst) return 0;   // query only
    const size_t pts = ctx->commit_log_batch * NCOMMIT;
    if (pts == 0) return fail(ctx, B2R_ERR_INVALID, "last_commitments: no completed prove call on this context");
    if (capacity_points < pts) return fail(ctx, B2R_ERR_INVALID, "last_commitments: destination too small");
    B2R_CUDA(ctx, cudaMemcpyAsync(dst, ctx->commit_log, pts * sizeof(affine_t), dst_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    if (!dst_on_device) B2R_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
} B2R_ABI_CATCH(ctx)

// RSASignatureVerifier::verify_pkcs1v15_signature from the MESSAGE on (reference src/lib.rs:183-248): SHA-256 of every
// message on the device (sha256.cu) -> the digest limbs of the sha_tail program's byte cells -> create_proof
int32_t b2r_rsa_prove_msgs_batch(b2r_ctx* ctx, const b2r_pk* pk, const uint64_t* n_limbs, const uint64_t* sig_limbs, const uint8_t* msgs,
                                 const uint64_t* msg_offsets, size_t batch, const uint8_t seed32[32], uint64_t nonce, uint32_t flags,
                                 uint8_t* proofs, uint8_t* status, uint8_t* digests) try {
    B2R_ENTER(ctx);
    if (!pk || !n_limbs || !sig_limbs || !msg_offsets || !seed32 || !proofs || !status) return fail(ctx, B2R_ERR_INVALID, "rsa_prove_msgs: null pointer");
    if (flags & ~(uint32_t)B2R_PROVE_SEED64) return fail(ctx, B2R_ERR_INVALID, "rsa_prove_msgs: unknown flag (inputs are host buffers)");
    if (pk->prog->aux_words != 4) return fail(ctx, B2R_ERR_INVALID, "rsa_prove_msgs: the key's program does not take a 4-limb digest");
    if (batch == 0) return 0;
    if (batch > 0xffffffffull) return fail(ctx, B2R_ERR_INVALID, "rsa_prove_msgs: batch too large");
    BlindKey bkey;
    if (flags & B2R_PROVE_SEED64) {
        uint64_t s64 = 0;
        for (int i = 0; i < 8; i++) s64 |= (uint64_t)seed32[i] << (8 * i);
        if (s64 == 0) return fail(ctx, B2R_ERR_INVALID, "rsa_prove_msgs: seed must be non-zero (blinding rows)");
        bkey = blind_key_from_seed64(s64, nonce);
    } else {
        bkey = blind_key_from_bytes(seed32, nonce);
    }
    B2R_TRY(sha256_check_offsets(ctx, msg_offsets, batch));
    const uint64_t base = msg_offsets[0], total = msg_offsets[batch] - base;
    if (total && !msgs) return fail(ctx, B2R_ERR_INVALID, "rsa_prove_msgs: null message buffer");
    const size_t nl = pk->prog->num_limbs;
    const size_t in_bytes = (batch * (2 * nl + 4) * 8 + 255) & ~(size_t)255, off_bytes = ((batch + 1) * 8 + 255) & ~(size_t)255,
                 dig_bytes = (batch * 32 + 255) & ~(size_t)255;
    char* d = nullptr;
    B2R_TRY(scratch_get(ctx, SC_MISC, in_bytes + off_bytes + dig_bytes + total + 256, (void**)&d));
    uint64_t *d_n = (uint64_t*)d, *d_s = d_n + batch * nl, *d_h = d_s + batch * nl;
    uint64_t* d_off = (uint64_t*)(d + in_bytes);
    uint8_t* d_dig = (uint8_t*)(d + in_bytes + off_bytes);
    uint8_t* d_msgs = (uint8_t*)(d + in_bytes + off_bytes + dig_bytes);
    std::vector<uint64_t> rel(batch + 1);
    for (size_t i = 0; i <= batch; i++) rel[i] = msg_offsets[i] - base;
    B2R_CUDA(ctx, cudaMemcpyAsync(d_n, n_limbs, batch * nl * 8, cudaMemcpyHostToDevice, ctx->stream));
    B2R_CUDA(ctx, cudaMemcpyAsync(d_s, sig_limbs, batch * nl * 8, cudaMemcpyHostToDevice, ctx->stream));
    B2R_CUDA(ctx, cudaMemcpyAsync(d_off, rel.data(), (batch + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (total) B2R_CUDA(ctx, cudaMemcpyAsync(d_msgs, msgs + base, total, cudaMemcpyHostToDevice, ctx->stream));
    B2R_TRY(sha256_msgs_dev(ctx, d_msgs, d_off, batch, d_h, 4, d_dig));
    if (digests) B2R_CUDA(ctx, cudaMemcpyAsync(digests, d_dig, batch * 32, cudaMemcpyDeviceToHost, ctx->stream));
    B2R_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // `rel` leaves scope below; the digests are complete for the caller
    return prove_all(ctx, pk, d_n, d_s, d_h, batch, bkey, proofs, status);
} B2R_ABI_CATCH(ctx)

}  // extern "C"
