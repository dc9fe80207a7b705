// RSA witness synthesis on the GPU: replay of a recorded circuit program for a batch.
//
// Replaces the witness side of the reference's Circuit::synthesize for the pkcs1v15
// circuit (benches/bench.rs:132-225 -> src/chip.rs:128-199 -> src/big_integer/chip.rs).
// circuit.hpp records the call sequence once (it is data independent); here
//   k_witness_eval  one CTA per proof instance walks the value graph level by level
//                   (nodes are sorted by dependency depth, so a level is a contiguous
//                   id range evaluated by all threads, __syncthreads between levels).
//                   Field values are kept in Montgomery form = halo2curves' memory format.
//                   The multi-limb BigUint steps (q, r = divrem(a*b, n), chip.rs:556-567;
//                   a - b, chip.rs:1297-1300) are "big ops" evaluated in shared memory.
//   k_witness_emit  gathers values into the 5 advice columns (column-major, 2^k rows,
//                   unassigned cells = 0, optional seeded blinding rows): pure HBM write
//                   stream, 5 * 2^k * 32 B per proof.
#include <cuda_runtime.h>

#include <algorithm>
#include <numeric>

#include "circuit.hpp"
#include "ctx.hpp"
#include "devutil.cuh"
#include "prog.hpp"

using namespace b2r;
using namespace b2r::circuit;

namespace b2r {

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ fe_t fe_from_u64_dev(uint64_t v) {
    fe_t c = Fr::zero();
    c.l[0] = (uint32_t)v;
    c.l[1] = (uint32_t)(v >> 32);
    return Fr::to_mont(c);
}
// canonical integer ops on 8 x u32
__device__ __forceinline__ fe_t int_shr(const fe_t& x, uint32_t b) {
    fe_t r = Fr::zero();
    uint32_t ws = b >> 5, bs = b & 31;
    for (int i = 0; i < 8; i++) {
        uint32_t src = i + ws;
        uint32_t lo = src < 8 ? x.l[src] : 0, hi = src + 1 < 8 ? x.l[src + 1] : 0;
        r.l[i] = bs ? ((lo >> bs) | (hi << (32 - bs))) : lo;
    }
    return r;
}
__device__ __forceinline__ fe_t int_lowbits(const fe_t& x, uint32_t b) {
    fe_t r;
    for (int i = 0; i < 8; i++) {
        uint32_t lo = 32 * i;
        if (b >= lo + 32) r.l[i] = x.l[i];
        else if (b <= lo) r.l[i] = 0;
        else r.l[i] = x.l[i] & ((1u << (b - lo)) - 1);
    }
    return r;
}
__device__ __forceinline__ fe_t int_clearlow(const fe_t& x, uint32_t b) {
    fe_t r;
    for (int i = 0; i < 8; i++) {
        uint32_t lo = 32 * i;
        if (b >= lo + 32) r.l[i] = 0;
        else if (b <= lo) r.l[i] = x.l[i];
        else r.l[i] = x.l[i] & ~((1u << (b - lo)) - 1);
    }
    return r;
}

// ---- big ops: BigUint arithmetic in shared memory, u32 limbs ------------------------------------
// to_big_uint (src/big_integer/mod.rs:348-359): sum_i int(limb_i) << (width * i); limbs are field
// elements and may exceed `width` bits, so this is a real multi-word accumulation.
__device__ void big_gather(const fe_t* values, const uint32_t* ids, uint32_t n, uint32_t width, uint32_t* out, uint32_t nwords) {
    for (uint32_t i = 0; i < nwords; i++) out[i] = 0;
    for (uint32_t i = 0; i < n; i++) {
        fe_t c = Fr::from_mont(ldv(values + ids[i]));
        uint32_t bit = width * i, ws = bit >> 5, bs = bit & 31;
        uint64_t carry = 0;
        for (uint32_t j = 0; j < 9; j++) {
            uint32_t lo = j < 8 ? c.l[j] : 0, prev = j > 0 ? c.l[j - 1] : 0;
            uint32_t w = bs ? ((lo << bs) | (prev >> (32 - bs))) : lo;
            if (ws + j >= nwords) break;
            carry += (uint64_t)out[ws + j] + w;
            out[ws + j] = (uint32_t)carry;
            carry >>= 32;
        }
        for (uint32_t j = ws + 9; carry && j < nwords; j++) {
            carry += out[j];
            out[j] = (uint32_t)carry;
            carry >>= 32;
        }
    }
}
__device__ __forceinline__ uint32_t big_len(const uint32_t* a, uint32_t n) {
    while (n > 0 && a[n - 1] == 0) n--;
    return n;
}
// Knuth D on normalised copies; un has m+n+1 words, vn has n words (n >= 2); q gets m+1 words
__device__ void big_divrem(uint32_t* un, uint32_t ulen, const uint32_t* v, uint32_t n, uint32_t* vn, uint32_t* q, uint32_t* rem) {
    // returns quotient in q[0..ulen-n], remainder in rem[0..n)
    uint32_t s = __clz(v[n - 1]);
    for (uint32_t i = n - 1; i > 0; i--) vn[i] = s ? ((v[i] << s) | (v[i - 1] >> (32 - s))) : v[i];
    vn[0] = v[0] << s;
    un[ulen] = s ? (un[ulen - 1] >> (32 - s)) : 0;
    for (uint32_t i = ulen - 1; i > 0; i--) un[i] = s ? ((un[i] << s) | (un[i - 1] >> (32 - s))) : un[i];
    un[0] = un[0] << s;
    uint32_t m = ulen - n;
    for (int32_t j = (int32_t)m; j >= 0; j--) {
        uint64_t num = ((uint64_t)un[j + n] << 32) | un[j + n - 1];
        uint64_t qhat = num / vn[n - 1], rhat = num % vn[n - 1];
        while (qhat >= (1ull << 32) || qhat * vn[n - 2] > ((rhat << 32) | un[j + n - 2])) {
            qhat--;
            rhat += vn[n - 1];
            if (rhat >= (1ull << 32)) break;
        }
        int64_t borrow = 0;
        uint64_t carry = 0;
        for (uint32_t i = 0; i < n; i++) {
            uint64_t p = qhat * vn[i] + carry;
            carry = p >> 32;
            int64_t t = (int64_t)un[i + j] - borrow - (int64_t)(p & 0xffffffffull);
            un[i + j] = (uint32_t)t;
            borrow = t < 0 ? 1 : 0;
        }
        int64_t t = (int64_t)un[j + n] - borrow - (int64_t)carry;
        un[j + n] = (uint32_t)t;
        if (t < 0) {
            qhat--;
            uint64_t c = 0;
            for (uint32_t i = 0; i < n; i++) {
                c += (uint64_t)un[i + j] + vn[i];
                un[i + j] = (uint32_t)c;
                c >>= 32;
            }
            un[j + n] += (uint32_t)c;
        }
        q[j] = (uint32_t)qhat;
    }
    for (uint32_t i = 0; i < n; i++) rem[i] = s ? ((un[i] >> s) | (un[i + 1] << (32 - s))) : un[i];
}

// one thread evaluates one big op; `sh` is a shared scratch of 6 * W words
__device__ bool big_eval(const BigOp& op, const uint32_t* ids, fe_t* values, uint32_t out0, uint32_t* sh, uint32_t W) {
    const uint32_t lw = op.limb_width;
    uint32_t* A = sh;           // W
    uint32_t* Bv = sh + W;      // W
    uint32_t* Nv = sh + 2 * W;  // W
    uint32_t* PR = sh + 3 * W;  // 2W + 2 (product / dividend)
    uint32_t* Q = sh + 5 * W + 2;  // W + 2
    uint32_t* VN = Q + W + 2;      // W
    bool ok = true;
    big_gather(values, ids, op.na, lw, A, W);
    big_gather(values, ids + op.na, op.nb, lw, Bv, W);
    if (op.kind == BIG_SUB) {
        // c = a - b ; a < b panics in the reference (BigUint underflow)
        int64_t br = 0;
        for (uint32_t i = 0; i < W; i++) {
            int64_t t = (int64_t)A[i] - Bv[i] - br;
            PR[i] = (uint32_t)t;
            br = t < 0 ? 1 : 0;
        }
        if (br) {
            ok = false;
            for (uint32_t i = 0; i < W; i++) PR[i] = 0;
        }
        for (uint32_t i = 0; i < op.nout; i++) {
            uint32_t bit = lw * i;  // lw == 64
            uint64_t v = (uint64_t)PR[bit >> 5] | ((uint64_t)PR[(bit >> 5) + 1] << 32);
            stv(values + out0 + i, fe_from_u64_dev(v));
        }
        return ok;
    }
    big_gather(values, ids + op.na + op.nb, op.nn, lw, Nv, W);
    uint32_t la = big_len(A, W), lb = big_len(Bv, W), ln = big_len(Nv, W);
    for (uint32_t i = 0; i < 2 * W + 2; i++) PR[i] = 0;
    for (uint32_t i = 0; i < la; i++) {
        uint64_t c = 0;
        uint32_t ai = A[i];
        for (uint32_t j = 0; j < lb; j++) {
            c += (uint64_t)ai * Bv[j] + PR[i + j];
            PR[i + j] = (uint32_t)c;
            c >>= 32;
        }
        PR[i + lb] = (uint32_t)c;
    }
    uint32_t lp = big_len(PR, 2 * W);
    for (uint32_t i = 0; i < W + 2; i++) Q[i] = 0;
    // remainder reuses A
    for (uint32_t i = 0; i < W; i++) A[i] = 0;
    if (ln == 0) {
        ok = false;  // division by zero panics in the reference
    } else if (lp < ln) {
        for (uint32_t i = 0; i < lp; i++) A[i] = PR[i];
    } else if (ln == 1) {
        uint64_t r = 0;
        for (int32_t i = (int32_t)lp - 1; i >= 0; i--) {
            uint64_t cur = (r << 32) | PR[i];
            uint32_t qd = (uint32_t)(cur / Nv[0]);
            if ((uint32_t)i < W + 2) Q[i] = qd; else if (qd) ok = false;
            r = cur % Nv[0];
        }
        A[0] = (uint32_t)r;
    } else {
        if (lp - ln + 1 > W + 2) {
            ok = false;
        } else {
            big_divrem(PR, lp, Nv, ln, VN, Q, A);
        }
    }
    // q must fit nb limbs and r must fit na limbs (asserts at chip.rs:583-584)
    uint32_t lq = big_len(Q, W + 2), lr = big_len(A, W);
    if (lq * 32 > op.nb * lw && lq > 0) {
        uint32_t bits = 32 * (lq - 1) + (32 - __clz(Q[lq - 1]));
        if (bits > op.nb * lw) ok = false;
    }
    if (lr * 32 > op.na * lw && lr > 0) {
        uint32_t bits = 32 * (lr - 1) + (32 - __clz(A[lr - 1]));
        if (bits > op.na * lw) ok = false;
    }
    for (uint32_t i = 0; i < op.nb; i++) {
        uint32_t w0 = 2 * i;
        uint64_t v = (uint64_t)(w0 < W + 2 ? Q[w0] : 0) | ((uint64_t)(w0 + 1 < W + 2 ? Q[w0 + 1] : 0) << 32);
        stv(values + out0 + i, fe_from_u64_dev(ok ? v : 0));
    }
    for (uint32_t i = 0; i < op.na; i++) {
        uint32_t w0 = 2 * i;
        uint64_t v = (uint64_t)(w0 < W ? A[w0] : 0) | ((uint64_t)(w0 + 1 < W ? A[w0 + 1] : 0) << 32);
        stv(values + out0 + op.nb + i, fe_from_u64_dev(ok ? v : 0));
    }
    return ok;
}

// ---- the same big ops evaluated by the whole CTA -----------------------------------------------------------
// The 19 modular multiplications of x^65537 mod n are a dependent chain, so each one is on the critical path of its
// instance; one thread doing schoolbook multiplication and Knuth division in shared memory took 0.38 ms per
// operation (8 ms per 64 instances: profiles/r01_ncu_full_baseline.md, k_witness_eval).  Here the CTA cooperates:
// limbs are gathered in parallel, products are formed column-wise (thread t sums the partial products of column t
// in 96 bits, one thread resolves the carries), and the division is a Barrett reduction with mu = floor(b^2k / n)
// computed once per modulus and cached in shared memory (two more column-parallel products and at most two
// correcting subtractions give the exact quotient and remainder).  Anything unusual - a limb wider than 64 bits, a
// one-word or zero modulus, a product longer than 2k words - takes the single-thread path above, which stays the
// definition of the result.
struct BigScratch {
    uint32_t *A, *B, *N, *P, *Q, *R, *MU, *NC, *T, *c0, *c1, *c2;
    uint32_t W;
    __device__ void carve(uint32_t* sh, uint32_t W_) {
        W = W_;
        uint32_t o = 7 * W + 16;  // the serial path's area comes first and is shared with it
        auto take = [&](uint32_t n) { uint32_t* r = sh + o; o += n; return r; };
        A = sh; B = sh + W; N = sh + 2 * W; P = sh + 3 * W;   // same places as in big_eval
        Q = take(W + 4); R = take(W + 4); MU = take(W + 4); NC = take(W + 4); T = take(2 * W + 8);
        c0 = take(2 * W + 8); c1 = take(2 * W + 8); c2 = take(2 * W + 8);
    }
    static __host__ __device__ uint32_t words(uint32_t W) { return 7 * W + 16 + 4 * (W + 4) + 4 * (2 * W + 8); }
};

// out[0 .. lx+ly) = X * Y, all threads; lx, ly >= 1
__device__ void cta_mul(const uint32_t* X, uint32_t lx, const uint32_t* Y, uint32_t ly, uint32_t* out, BigScratch& S) {
    const uint32_t tid = threadIdx.x, T = blockDim.x, cols = lx + ly;
    for (uint32_t t = tid; t < cols; t += T) {
        uint64_t lo = 0;
        uint32_t hi = 0;
        const uint32_t i0 = t >= ly ? t - ly + 1 : 0, i1 = t < lx ? t : lx - 1;
        for (uint32_t i = i0; i <= i1 && t < cols - 1; i++) {
            const uint64_t pr = (uint64_t)X[i] * Y[t - i];
            lo += pr;
            hi += lo < pr ? 1u : 0u;
        }
        S.c0[t] = (uint32_t)lo;
        S.c1[t] = (uint32_t)(lo >> 32);
        S.c2[t] = hi;
    }
    __syncthreads();
    if (tid == 0) {
        uint64_t carry = 0;
        for (uint32_t t = 0; t < cols; t++) {
            carry += (uint64_t)S.c0[t] + (t >= 1 ? S.c1[t - 1] : 0) + (t >= 2 ? S.c2[t - 2] : 0);
            out[t] = (uint32_t)carry;
            carry >>= 32;
        }
    }
    __syncthreads();
}

// returns false if this op must go through the serial path; *ok = false if the reference would have panicked
__device__ bool big_eval_cta(const BigOp& op, const uint32_t* ids, fe_t* values, uint32_t out0, BigScratch& S, int* sh_flag, bool* ok) {
    const uint32_t tid = threadIdx.x, T = blockDim.x, W = S.W;
    if (op.limb_width != 64) return false;
    const uint32_t nin = op.na + op.nb + op.nn;
    if (2 * op.na > W || 2 * op.nb > W || 2 * op.nn > W) return false;
    // ---- gather (limbs < 2^64 go straight to their two words)
    if (tid == 0) *sh_flag = 0;
    for (uint32_t i = tid; i < 3 * W; i += T) S.A[i] = 0;   // A, B, N are contiguous
    __syncthreads();
    for (uint32_t i = tid; i < nin; i += T) {
        const fe_t c = Fr::from_mont(ldv(values + ids[i]));
        if (c.l[2] | c.l[3] | c.l[4] | c.l[5] | c.l[6] | c.l[7]) atomicOr(sh_flag, 1);
        uint32_t* dst = i < op.na ? S.A + 2 * i : (i < op.na + op.nb ? S.B + 2 * (i - op.na) : S.N + 2 * (i - op.na - op.nb));
        dst[0] = c.l[0];
        dst[1] = c.l[1];
    }
    __syncthreads();
    const bool wide_limb = *sh_flag != 0;
    __syncthreads();  // every thread has read the flag before it is reused below
    if (wide_limb) return false;
    *ok = true;
    if (op.kind == BIG_SUB) {
        if (tid == 0) {
            int64_t br = 0;
            for (uint32_t i = 0; i < W; i++) {
                int64_t t = (int64_t)S.A[i] - S.B[i] - br;
                S.P[i] = (uint32_t)t;
                br = t < 0 ? 1 : 0;
            }
            *sh_flag = br ? 2 : 0;
        }
        __syncthreads();
        const bool under = *sh_flag == 2;
        for (uint32_t i = tid; i < op.nout; i += T) {
            const uint64_t v = under ? 0 : ((uint64_t)S.P[2 * i] | ((uint64_t)S.P[2 * i + 1] << 32));
            stv(values + out0 + i, fe_from_u64_dev(v));
        }
        *ok = !under;
        __syncthreads();
        return true;
    }
    const uint32_t la = big_len(S.A, W), lb = big_len(S.B, W), k = big_len(S.N, W);
    if (k < 2 || la == 0 || lb == 0 || la + lb > 2 * k) return false;   // uniform: every thread reads the same shared words
    // ---- mu for this modulus (cached)
    if (tid == 0) *sh_flag = 0;
    __syncthreads();
    for (uint32_t i = tid; i < W; i += T)
        if (S.NC[i] != S.N[i]) atomicOr(sh_flag, 1);
    __syncthreads();
    const bool new_modulus = *sh_flag != 0;
    __syncthreads();
    if (new_modulus) {
        if (tid == 0) {
            // mu = floor(b^(2k) / N): Knuth D on (1, 0 x 2k) by N -> k + 1 quotient words
            uint32_t* un = S.T;   // 2k + 2 words
            for (uint32_t i = 0; i < 2 * k + 2; i++) un[i] = 0;
            un[2 * k] = 1;
            for (uint32_t i = 0; i < W + 4; i++) S.MU[i] = 0;
            big_divrem(un, 2 * k + 1, S.N, k, S.c0, S.MU, S.c1);
            for (uint32_t i = 0; i < W; i++) S.NC[i] = S.N[i];
        }
        __syncthreads();
    }
    // ---- x = A * B
    for (uint32_t i = tid; i < 2 * W + 2; i += T) S.P[i] = 0;
    __syncthreads();
    cta_mul(S.A, la, S.B, lb, S.P, S);
    const uint32_t lp = big_len(S.P, 2 * W);
    // ---- Barrett: q3 = floor(floor(x / b^(k-1)) * mu / b^(k+1))
    const uint32_t lq1 = lp > k - 1 ? lp - (k - 1) : 0;
    for (uint32_t i = tid; i < W + 4; i += T) S.Q[i] = 0;
    __syncthreads();
    if (lq1) {
        const uint32_t lmu = big_len(S.MU, k + 2);
        cta_mul(S.P + (k - 1), lq1, S.MU, lmu, S.T, S);
        const uint32_t lt = lq1 + lmu;
        for (uint32_t i = tid; i + (k + 1) < lt && i < W + 4; i += T) S.Q[i] = S.T[i + k + 1];
        __syncthreads();
    }
    // ---- r = x - q3 * N  (mod b^(k+1)), then at most two corrections
    const uint32_t lq3 = big_len(S.Q, k + 2);
    for (uint32_t i = tid; i < 2 * W + 8; i += T) S.T[i] = 0;
    __syncthreads();
    if (lq3) cta_mul(S.Q, lq3, S.N, k, S.T, S);
    if (tid == 0) {
        int64_t br = 0;
        for (uint32_t i = 0; i < k + 1; i++) {
            int64_t t = (int64_t)S.P[i] - S.T[i] - br;
            S.R[i] = (uint32_t)t;
            br = t < 0 ? 1 : 0;
        }
        // r < 3N (Barrett): subtract N while r >= N
        for (int iter = 0; iter < 3; iter++) {
            bool ge = true;
            if (S.R[k] == 0) {
                for (int32_t i = (int32_t)k - 1; i >= 0; i--)
                    if (S.R[i] != S.N[i]) { ge = S.R[i] > S.N[i]; break; }
            }
            if (!ge) break;
            int64_t b2 = 0;
            for (uint32_t i = 0; i < k + 1; i++) {
                int64_t t = (int64_t)S.R[i] - (i < k ? S.N[i] : 0) - b2;
                S.R[i] = (uint32_t)t;
                b2 = t < 0 ? 1 : 0;
            }
            for (uint32_t i = 0; i < W + 4; i++)
                if (++S.Q[i]) break;
        }
        // q must fit nb limbs and r must fit na limbs (asserts at chip.rs:583-584)
        bool good = true;
        const uint32_t lq = big_len(S.Q, W + 4), lr = big_len(S.R, k + 1);
        if (lq > 2 * op.nb || lr > 2 * op.na) good = false;
        *sh_flag = good ? 0 : 2;
    }
    __syncthreads();
    const bool good = *sh_flag == 0;
    for (uint32_t i = tid; i < op.nb; i += T) {
        const uint64_t v = (uint64_t)S.Q[2 * i] | ((uint64_t)S.Q[2 * i + 1] << 32);
        stv(values + out0 + i, fe_from_u64_dev(good ? v : 0));
    }
    for (uint32_t i = tid; i < op.na; i += T) {
        const uint64_t v = (uint64_t)(2 * i < k + 1 ? S.R[2 * i] : 0) | ((uint64_t)(2 * i + 1 < k + 1 ? S.R[2 * i + 1] : 0) << 32);
        stv(values + out0 + op.nb + i, fe_from_u64_dev(good ? v : 0));
    }
    *ok = good;
    __syncthreads();
    return true;
}

struct WitnessArgs {
    const Node* nodes;
    const LevelRange* levels;
    const uint32_t* big_nodes;
    const BigOp* big_ops;
    const uint32_t* big_inputs;
    const fe_t* consts;
    const uint64_t* n_limbs;
    const uint64_t* sig_limbs;
    const uint64_t* hash_limbs;
    fe_t* values;  // [batch][num_values]
    uint8_t* is_valid;
    uint32_t num_levels, num_values, num_limbs, big_words, aux_words;
    int32_t is_valid_vid;
};

__global__ void __launch_bounds__(1024) k_witness_eval(const WitnessArgs A) {
    extern __shared__ uint32_t sh_big[];
    __shared__ int sh_err, sh_flag;
    const uint32_t p = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    fe_t* val = A.values + (size_t)p * A.num_values;
    BigScratch BS;
    BS.carve(sh_big, A.big_words);
    if (tid == 0) sh_err = 0;
    for (uint32_t i = tid; i < A.big_words + 4; i += T) BS.NC[i] = 0;   // no modulus cached yet (a zero modulus never gets here)
    __syncthreads();
    for (uint32_t lv = 0; lv < A.num_levels; lv++) {
        const LevelRange L = A.levels[lv];
        for (uint32_t id = L.start + tid; id < L.end; id += T) {
            const Node nd = A.nodes[id];
            fe_t r;
            switch (nd.op) {
                case OP_CONST: r = ldv(A.consts + nd.a); break;
                case OP_INPUT: {
                    uint32_t w = nd.a, nl = A.num_limbs;
                    uint64_t v = w < nl ? A.n_limbs[(size_t)p * nl + w]
                               : w < 2 * nl ? A.sig_limbs[(size_t)p * nl + (w - nl)]
                                            : A.hash_limbs[(size_t)p * A.aux_words + (w - 2 * nl)];
                    r = fe_from_u64_dev(v);
                    break;
                }
                case OP_ADD: r = Fr::add(ldv(val + nd.a), ldv(val + nd.b)); break;
                case OP_SUB: r = Fr::sub(ldv(val + nd.a), ldv(val + nd.b)); break;
                case OP_MUL: r = Fr::mul(ldv(val + nd.a), ldv(val + nd.b)); break;
                case OP_MULADD: r = Fr::add(Fr::mul(ldv(val + nd.a), ldv(val + nd.b)), ldv(val + nd.c)); break;
                case OP_ADDC: r = Fr::add(ldv(val + nd.a), ldv(A.consts + nd.b)); break;
                case OP_ADD2C: r = Fr::add(Fr::add(ldv(val + nd.a), ldv(val + nd.b)), ldv(A.consts + nd.c)); break;
                case OP_NOT: r = Fr::sub(Fr::one(), ldv(val + nd.a)); break;
                case OP_SELECT: {
                    fe_t c = ldv(val + nd.c);
                    r = Fr::eq(c, Fr::one()) ? ldv(val + nd.a) : ldv(val + nd.b);
                    break;
                }
                case OP_ISZERO: r = Fr::is_zero(ldv(val + nd.a)) ? Fr::one() : Fr::zero(); break;
                case OP_INVORONE: {
                    fe_t x = ldv(val + nd.a);
                    r = Fr::is_zero(x) ? Fr::one() : Fr::inv(x);
                    break;
                }
                case OP_SHR: r = Fr::to_mont(int_shr(Fr::from_mont(ldv(val + nd.a)), nd.b)); break;
                case OP_LOWBITS: r = Fr::to_mont(int_lowbits(Fr::from_mont(ldv(val + nd.a)), nd.b)); break;
                case OP_SUBLIMB: r = Fr::to_mont(int_lowbits(int_shr(Fr::from_mont(ldv(val + nd.a)), nd.b), nd.c)); break;
                case OP_CLEARLOW: r = Fr::to_mont(int_clearlow(Fr::from_mont(ldv(val + nd.a)), nd.b)); break;
                default: continue;  // OP_BIG / OP_BIGOUT: written below
            }
            stv(val + id, r);
        }
        __syncthreads();  // the level's ordinary nodes are written before the big ops read them
        for (uint32_t bi = L.bstart; bi < L.bend; bi++) {   // uniform across the CTA
            const uint32_t id = A.big_nodes[bi];
            const BigOp op = A.big_ops[A.nodes[id].a];
            bool ok = true;
            if (!big_eval_cta(op, A.big_inputs + op.in_off, val, id + 1, BS, &sh_flag, &ok)) {
                __syncthreads();
                if (tid == 0) ok = big_eval(op, A.big_inputs + op.in_off, val, id + 1, sh_big, A.big_words);
                if (tid != 0) ok = true;
            }
            if (!ok && tid == 0) sh_err = 1;
            __syncthreads();
        }
        __syncthreads();
    }
    if (tid == 0) {
        uint8_t v = 0xff;
        if (!sh_err) v = Fr::eq(ldv(val + A.is_valid_vid), Fr::one()) ? 1 : 0;
        A.is_valid[p] = v;
    }
}

__global__ void __launch_bounds__(256)
k_witness_emit(const int32_t* __restrict__ cellmap, const fe_t* __restrict__ values, uint32_t num_values, uint32_t k,
               uint32_t usable_rows, const BlindKey bkey, uint32_t p_base, fe_t* __restrict__ advice, size_t p_stride,
               size_t col_stride) {
    const uint32_t n = 1u << k;
    uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t col = blockIdx.y, p = blockIdx.z;
    if (row >= n) return;
    fe_t v;
    if (row >= usable_rows) {
        v = bkey.on ? blind_value(bkey, p_base + p, ST_ADVICE + col, row) : Fr::zero();
    } else {
        int32_t id = cellmap[(size_t)col * n + row];
        v = id < 0 ? Fr::zero() : ldv(values + (size_t)p * num_values + id);
    }
    stv(advice + (size_t)p * p_stride + (size_t)col * col_stride + row, v);
}

}  // namespace b2r

// ---- program construction (host) ------------------------------------------------------------------
static fe_t u256_to_mont(const U256& v) {
    fe_t c;
    for (int i = 0; i < 4; i++) {
        c.l[2 * i] = (uint32_t)v.l[i];
        c.l[2 * i + 1] = (uint32_t)(v.l[i] >> 32);
    }
    return Fr::to_mont(c);
}

template <class T>
static int32_t upload(b2r_ctx* ctx, const std::vector<T>& v, T** d) {
    size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    B2R_CUDA(ctx, cudaMalloc((void**)d, bytes));
    if (!v.empty()) B2R_CUDA(ctx, cudaMemcpyAsync(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

// sorts the recorded nodes by level, renumbers value ids, uploads everything
static int32_t finalize_program(b2r_ctx* ctx, RegionCtx& rc, AssignedValue is_valid, b2r_prog* prog) {
    const size_t N = rc.nodes.size();
    std::vector<uint32_t> order(N);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return rc.level[a] < rc.level[b]; });
    std::vector<uint32_t> newid(N);
    for (size_t i = 0; i < N; i++) newid[order[i]] = (uint32_t)i;
    std::vector<Node> nodes(N);
    uint32_t nlev = 0;
    for (size_t i = 0; i < N; i++) nlev = std::max(nlev, rc.level[i] + 1);
    std::vector<LevelRange> levels(nlev);
    for (auto& l : levels) l.start = l.end = l.bstart = l.bend = 0;
    std::vector<uint32_t> big_nodes;
    for (size_t i = 0; i < N; i++) {
        Node nd = rc.nodes[order[i]];
        switch (nd.op) {
            case OP_ADD: case OP_SUB: case OP_MUL: case OP_ADD2C:
                nd.a = newid[nd.a]; nd.b = newid[nd.b]; break;
            case OP_MULADD: case OP_SELECT:
                nd.a = newid[nd.a]; nd.b = newid[nd.b]; nd.c = newid[nd.c]; break;
            case OP_ADDC: case OP_NOT: case OP_ISZERO: case OP_INVORONE: case OP_SHR: case OP_LOWBITS: case OP_SUBLIMB: case OP_CLEARLOW:
                nd.a = newid[nd.a]; break;
            default: break;
        }
        nodes[i] = nd;
        uint32_t lv = rc.level[order[i]];
        if (i == 0 || rc.level[order[i - 1]] != lv) {
            levels[lv].start = (uint32_t)i;
            levels[lv].bstart = (uint32_t)big_nodes.size();
        }
        levels[lv].end = (uint32_t)i + 1;
        if (nd.op == OP_BIG) big_nodes.push_back((uint32_t)i);
        levels[lv].bend = (uint32_t)big_nodes.size();
    }
    for (auto& id : rc.big_inputs) id = newid[id];
    const uint32_t n = 1u << prog->k;
    std::vector<int32_t> cellmap((size_t)NUM_ADVICE * n, -1);
    for (int c = 0; c < NUM_ADVICE; c++)
        for (size_t r = 0; r < rc.cell[c].size(); r++)
            if (rc.cell[c][r] >= 0) cellmap[(size_t)c * n + r] = (int32_t)newid[rc.cell[c][r]];
    std::vector<fe_t> consts(rc.constants.size());
    for (size_t i = 0; i < consts.size(); i++) consts[i] = u256_to_mont(rc.constants[i]);
    uint32_t maxw = 0;
    for (const auto& op : rc.big_ops) {
        uint32_t w = std::max(std::max(op.na, op.nb), op.nn) * op.limb_width / 32 + 10;
        maxw = std::max(maxw, w);
    }
    prog->max_big_words = maxw;
    prog->rows_used = rc.offset;
    prog->num_values = (uint32_t)N;
    prog->num_levels = nlev;
    prog->is_valid_vid = (int32_t)newid[is_valid.vid];
    B2R_TRY(upload(ctx, nodes, &prog->d_nodes));
    B2R_TRY(upload(ctx, levels, &prog->d_levels));
    B2R_TRY(upload(ctx, big_nodes, &prog->d_big_nodes));
    B2R_TRY(upload(ctx, rc.big_ops, &prog->d_big_ops));
    B2R_TRY(upload(ctx, rc.big_inputs, &prog->d_big_inputs));
    B2R_TRY(upload(ctx, consts, &prog->d_consts));
    B2R_TRY(upload(ctx, cellmap, &prog->d_cellmap));
    B2R_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    prog->fixed = std::move(rc.fixed);
    prog->range_tags = std::move(rc.range_tags);
    prog->copies = std::move(rc.copies);
    prog->constants = rc.constants;
    for (int b = 1; b < 80; b++)
        if (rc.tag_of_bits[b] > 0 && rc.tag_of_bits[b] < 16) prog->tag_bits[rc.tag_of_bits[b]] = (uint8_t)b;
    return 0;
}

static void prog_release(b2r_prog* p) {
    cudaFree(p->d_nodes);
    cudaFree(p->d_levels);
    cudaFree(p->d_big_nodes);
    cudaFree(p->d_big_ops);
    cudaFree(p->d_big_inputs);
    cudaFree(p->d_consts);
    cudaFree(p->d_cellmap);
    delete p;
}


extern "C" {

}  // extern "C"

// shared tail of the program builders: `record` fills the RegionCtx and returns the cell that reports the outcome
template <class Rec>
static int32_t build_program(b2r_ctx* ctx, uint32_t bits_len, uint32_t k, uint32_t aux_words, Rec record, b2r_prog** out) {
    *out = nullptr;
    if (bits_len < 512 || bits_len > 4096 || bits_len % 64) return fail(ctx, B2R_ERR_INVALID, "program_build: bits_len must be a multiple of 64 in [512, 4096]");
    if (k < 4 || k > 24) return fail(ctx, B2R_ERR_INVALID, "program_build: k out of range");
    b2r_prog* prog = new b2r_prog();
    prog->bits_len = bits_len;
    prog->num_limbs = bits_len / 64;
    prog->k = k;
    prog->aux_words = aux_words;
    prog->num_inputs = 2 * prog->num_limbs + aux_words;
    try {
        RegionCtx rc((1u << k) - BLINDING_ROWS);
        AssignedValue result = record(rc);
        int32_t r = finalize_program(ctx, rc, result, prog);
        if (r) {
            prog_release(prog);
            return r;
        }
    } catch (const SynthError& e) {
        prog_release(prog);
        return fail(ctx, e.code == -5 ? B2R_ERR_LAYOUT : (e.code == -1 ? B2R_ERR_INVALID : B2R_ERR_SYNTH), e.what());
    } catch (const std::exception& e) {
        prog_release(prog);
        return fail(ctx, B2R_ERR_NOMEM, e.what());
    }
    *out = prog;
    return 0;
}

extern "C" {

int32_t b2r_rsa_program_build(b2r_ctx* ctx, uint32_t bits_len, const uint8_t* e_le, size_t e_len, uint32_t k, b2r_prog** out) try {
    B2R_ENTER(ctx);
    if (!out || !e_le || e_len == 0) return fail(ctx, B2R_ERR_INVALID, "rsa_program_build: null argument");
    *out = nullptr;
    if (bits_len < 512 || bits_len > 4096 || bits_len % 64) return fail(ctx, B2R_ERR_INVALID, "rsa_program_build: bits_len must be a multiple of 64 in [512, 4096]");
    if (k < 4 || k > 24) return fail(ctx, B2R_ERR_INVALID, "rsa_program_build: k out of range");
    bool e_nonzero = false;
    for (size_t i = 0; i < e_len; i++) e_nonzero |= e_le[i] != 0;
    if (!e_nonzero) return fail(ctx, B2R_ERR_INVALID, "rsa_program_build: exponent is zero");
    std::vector<uint8_t> e(e_le, e_le + e_len);
    return build_program(ctx, bits_len, k, 4, [&](RegionCtx& rc) { return record_rsa_pkcs1v15(rc, bits_len, e); }, out);
} B2R_ABI_CATCH(ctx)

int32_t b2r_rsa_program_build_sha_tail(b2r_ctx* ctx, uint32_t bits_len, const uint8_t* e_le, size_t e_len, uint32_t k, b2r_prog** out) try {
    B2R_ENTER(ctx);
    if (!out || !e_le || e_len == 0) return fail(ctx, B2R_ERR_INVALID, "rsa_program_build_sha_tail: null argument");
    *out = nullptr;
    bool e_nonzero = false;
    for (size_t i = 0; i < e_len; i++) e_nonzero |= e_le[i] != 0;
    if (!e_nonzero) return fail(ctx, B2R_ERR_INVALID, "rsa_program_build_sha_tail: exponent is zero");
    std::vector<uint8_t> e(e_le, e_le + e_len);
    return build_program(ctx, bits_len, k, 4, [&](RegionCtx& rc) { return record_rsa_verifier_from_digest(rc, bits_len, e); }, out);
} B2R_ABI_CATCH(ctx)

int32_t b2r_rsa_program_build_var(b2r_ctx* ctx, uint32_t bits_len, uint32_t exp_limb_bits, uint32_t k, b2r_prog** out) try {
    B2R_ENTER(ctx);
    if (!out) return fail(ctx, B2R_ERR_INVALID, "rsa_program_build_var: null argument");
    if (exp_limb_bits == 0 || exp_limb_bits > 64) return fail(ctx, B2R_ERR_INVALID, "rsa_program_build_var: exp_limb_bits must be in [1, 64]");
    return build_program(ctx, bits_len, k, 5, [&](RegionCtx& rc) { return record_rsa_pkcs1v15_var(rc, bits_len, exp_limb_bits); }, out);
} B2R_ABI_CATCH(ctx)

int32_t b2r_bigint_program_build(b2r_ctx* ctx, uint32_t op, uint32_t bits_len, uint32_t exp_limb_bits, uint32_t k, b2r_prog** out) try {
    B2R_ENTER(ctx);
    if (!out) return fail(ctx, B2R_ERR_INVALID, "bigint_program_build: null argument");
    if (op < BT_REFRESH || op > BT_SQUARE_MOD) return fail(ctx, B2R_ERR_INVALID, "bigint_program_build: unknown operation");
    if (exp_limb_bits == 0 || exp_limb_bits > 64) return fail(ctx, B2R_ERR_INVALID, "bigint_program_build: exp_limb_bits must be in [1, 64]");
    // inputs a | b | n | e: the third array carries n and e (num_limbs + 1 words per instance)
    return build_program(ctx, bits_len, k, bits_len / 64 + 1, [&](RegionCtx& rc) { return record_bigint_op(rc, op, bits_len, exp_limb_bits, nullptr); }, out);
} B2R_ABI_CATCH(ctx)

int32_t b2r_prog_free(b2r_ctx* ctx, b2r_prog* prog) try {
    B2R_ENTER(ctx);
    if (!ctx || !prog) return B2R_ERR_INVALID;
    cudaStreamSynchronize(ctx->stream);
    prog_release(prog);
    return 0;
} B2R_ABI_CATCH(ctx)

int32_t b2r_prog_num_limbs(const b2r_prog* prog) { return prog ? (int32_t)prog->num_limbs : B2R_ERR_INVALID; }
int32_t b2r_prog_aux_words(const b2r_prog* prog) { return prog ? (int32_t)prog->aux_words : B2R_ERR_INVALID; }

int32_t b2r_prog_info(const b2r_prog* prog, uint64_t* rows_used, uint64_t* num_values, uint64_t* num_levels) {
    if (!prog) return B2R_ERR_INVALID;
    if (rows_used) *rows_used = prog->rows_used;
    if (num_values) *num_values = prog->num_values;
    if (num_levels) *num_levels = prog->num_levels;
    return 0;
}

}  // extern "C"

namespace b2r {
// advice cell (p, col, row) is written to advice_dev[p * p_stride + col * col_stride + row]
int32_t witness_run(b2r_ctx* ctx, const b2r_prog* prog, const uint64_t* n_limbs_dev, const uint64_t* sig_limbs_dev,
                    const uint64_t* hash_limbs_dev, size_t batch, const BlindKey& bkey, b2r_fr* advice_dev,
                    uint8_t* is_valid_dev, size_t p_base, size_t p_stride, size_t col_stride) {
    if (!ctx) return B2R_ERR_INVALID;
    if (!prog || !n_limbs_dev || !sig_limbs_dev || !hash_limbs_dev || !advice_dev || !is_valid_dev)
        return fail(ctx, B2R_ERR_INVALID, "rsa_witness: null pointer");
    if (batch == 0) return 0;
    const uint32_t n = 1u << prog->k;
    // value store: process the batch in groups that fit a bounded arena
    size_t per = (size_t)prog->num_values * sizeof(fe_t);
    size_t G = std::max<size_t>(1, std::min<size_t>(batch, ((size_t)2 << 30) / per));
    fe_t* values = nullptr;
    B2R_TRY(scratch_get(ctx, SC_WIT, G * per, (void**)&values));
    size_t big_smem = (size_t)BigScratch::words(prog->max_big_words) * sizeof(uint32_t);
    if (big_smem > 48 * 1024) B2R_CUDA(ctx, cudaFuncSetAttribute(k_witness_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_smem));
    for (size_t p0 = 0; p0 < batch; p0 += G) {
        size_t g = std::min(G, batch - p0);
        WitnessArgs A;
        A.nodes = prog->d_nodes;
        A.levels = prog->d_levels;
        A.big_nodes = prog->d_big_nodes;
        A.big_ops = prog->d_big_ops;
        A.big_inputs = prog->d_big_inputs;
        A.consts = prog->d_consts;
        A.n_limbs = n_limbs_dev + p0 * prog->num_limbs;
        A.sig_limbs = sig_limbs_dev + p0 * prog->num_limbs;
        A.hash_limbs = hash_limbs_dev + p0 * prog->aux_words;
        A.aux_words = prog->aux_words;
        A.values = values;
        A.is_valid = is_valid_dev + p0;
        A.num_levels = prog->num_levels;
        A.num_values = prog->num_values;
        A.num_limbs = prog->num_limbs;
        A.big_words = prog->max_big_words;
        A.is_valid_vid = prog->is_valid_vid;
        { KTimer kt(ctx, "witness_eval", (double)g);
        k_witness_eval<<<(unsigned)g, 1024, big_smem, ctx->stream>>>(A); }
        B2R_LAUNCH_CHECK(ctx);
        dim3 grid((n + 255) / 256, NUM_ADVICE, (unsigned)g);
        KTimer kt_emit(ctx, "witness_emit", (double)g);
        k_witness_emit<<<grid, 256, 0, ctx->stream>>>(prog->d_cellmap, values, prog->num_values, prog->k, n - BLINDING_ROWS, bkey,
                                                      (uint32_t)(p_base + p0), (fe_t*)advice_dev + p0 * p_stride, p_stride, col_stride);
        B2R_LAUNCH_CHECK(ctx);
    }
    return 0;
}
}  // namespace b2r

extern "C" {

int32_t b2r_rsa_witness_batch_dev(b2r_ctx* ctx, const b2r_prog* prog, const uint64_t* n_limbs_dev, const uint64_t* sig_limbs_dev,
                                  const uint64_t* hash_limbs_dev, size_t batch, uint64_t blind_seed, b2r_fr* advice_dev,
                                  uint8_t* is_valid_dev) try {
    B2R_ENTER(ctx);
    if (!prog) return ctx ? fail(ctx, B2R_ERR_INVALID, "rsa_witness: null pointer") : B2R_ERR_INVALID;
    const size_t n = (size_t)1 << prog->k;
    return witness_run(ctx, prog, n_limbs_dev, sig_limbs_dev, hash_limbs_dev, batch, blind_key_from_seed64(blind_seed, 0), advice_dev, is_valid_dev, 0, NUM_ADVICE * n, n);
} B2R_ABI_CATCH(ctx)

int32_t b2r_rsa_witness_batch(b2r_ctx* ctx, const b2r_prog* prog, const uint64_t* n_limbs, const uint64_t* sig_limbs,
                              const uint64_t* hash_limbs, size_t batch, uint64_t blind_seed, b2r_fr* advice, uint8_t* is_valid) try {
    B2R_ENTER(ctx);
    if (!prog || !n_limbs || !sig_limbs || !hash_limbs || !advice || !is_valid) return fail(ctx, B2R_ERR_INVALID, "rsa_witness: null pointer");
    if (batch == 0) return 0;
    const size_t n = (size_t)1 << prog->k, nl = prog->num_limbs;
    const size_t aw = prog->aux_words;
    const size_t in_bytes = batch * (2 * nl + aw) * 8;
    const size_t in_al = (in_bytes + 255) & ~(size_t)255;
    // stage a bounded number of instances at a time (20 MiB of advice each at k = 17)
    const size_t per_adv = NUM_ADVICE * n * sizeof(fe_t);
    const size_t G = std::max<size_t>(1, std::min<size_t>(batch, ((size_t)4 << 30) / per_adv));
    char* d = nullptr;
    B2R_TRY(scratch_get(ctx, SC_STAGE, in_al + 256 + ((batch + 255) & ~(size_t)255) + G * per_adv, (void**)&d));
    uint64_t* d_n = (uint64_t*)d;
    uint64_t* d_s = d_n + batch * nl;
    uint64_t* d_h = d_s + batch * nl;
    uint8_t* d_valid = (uint8_t*)(d + in_al);
    fe_t* d_adv = (fe_t*)(d + in_al + 256 + ((batch + 255) & ~(size_t)255));
    B2R_CUDA(ctx, cudaMemcpyAsync(d_n, n_limbs, batch * nl * 8, cudaMemcpyHostToDevice, ctx->stream));
    B2R_CUDA(ctx, cudaMemcpyAsync(d_s, sig_limbs, batch * nl * 8, cudaMemcpyHostToDevice, ctx->stream));
    B2R_CUDA(ctx, cudaMemcpyAsync(d_h, hash_limbs, batch * aw * 8, cudaMemcpyHostToDevice, ctx->stream));
    for (size_t p0 = 0; p0 < batch; p0 += G) {
        size_t g = std::min(G, batch - p0);
        B2R_TRY(witness_run(ctx, prog, d_n + p0 * nl, d_s + p0 * nl, d_h + p0 * aw, g, blind_key_from_seed64(blind_seed, 0), (b2r_fr*)d_adv, d_valid + p0, p0, NUM_ADVICE * n, n));
        B2R_CUDA(ctx, cudaMemcpyAsync((char*)advice + p0 * per_adv, d_adv, g * per_adv, cudaMemcpyDeviceToHost, ctx->stream));
        B2R_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    B2R_CUDA(ctx, cudaMemcpyAsync(is_valid, d_valid, batch, cudaMemcpyDeviceToHost, ctx->stream));
    B2R_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
} B2R_ABI_CATCH(ctx)

}  // extern "C"
