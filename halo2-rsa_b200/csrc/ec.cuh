// BN254 G1 (y^2 = x^3 + 3 over Fq) point arithmetic in extended Jacobian ("XYZZ")
// coordinates: x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; identity has ZZ = 0.
// Mixed addition costs 8M + 2S, full addition 12M + 2S (EFD madd-2008-s / add-2008-s /
// dbl-2008-s-1 with a = 0).  All exceptional cases (identity operands, P + P, P - P) are
// handled because resident base tables do contain repeated points (2^c * P_i == P_j).
#pragma once
#include "field.cuh"

namespace b2r {

struct affine_t {
    fe_t x, y;  // identity = (0, 0), halo2curves convention
};
struct xyzz_t {
    fe_t x, y, zz, zzz;
};

B2R_HD bool affine_is_identity(const affine_t& p) { return Fq::is_zero(p.x) && Fq::is_zero(p.y); }
B2R_HD bool xyzz_is_identity(const xyzz_t& p) { return Fq::is_zero(p.zz); }
B2R_HD xyzz_t xyzz_identity() {
    xyzz_t r;
    r.x = Fq::zero();
    r.y = Fq::zero();
    r.zz = Fq::zero();
    r.zzz = Fq::zero();
    return r;
}
B2R_HD xyzz_t xyzz_from_affine(const affine_t& p) {
    if (affine_is_identity(p)) return xyzz_identity();
    xyzz_t r;
    r.x = p.x;
    r.y = p.y;
    r.zz = Fq::one();
    r.zzz = Fq::one();
    return r;
}

// 2 * (x, y) for an affine, non-identity point (mdbl-2008-s-1)
B2R_HD xyzz_t xyzz_double_affine(const affine_t& p) {
    fe_t U = Fq::dbl(p.y);
    fe_t V = Fq::sqr(U);
    fe_t W = Fq::mul(U, V);
    fe_t S = Fq::mul(p.x, V);
    fe_t xx = Fq::sqr(p.x);
    fe_t M = Fq::add(Fq::dbl(xx), xx);
    xyzz_t r;
    r.x = Fq::sub(Fq::sqr(M), Fq::dbl(S));
    r.y = Fq::mul_sub_mul(M, Fq::sub(S, r.x), W, p.y);
    r.zz = V;
    r.zzz = W;
    return r;  // y == 0 cannot happen on a prime-order curve
}

B2R_HD xyzz_t xyzz_double(const xyzz_t& p) {
    if (xyzz_is_identity(p)) return p;
    fe_t U = Fq::dbl(p.y);
    fe_t V = Fq::sqr(U);
    fe_t W = Fq::mul(U, V);
    fe_t S = Fq::mul(p.x, V);
    fe_t xx = Fq::sqr(p.x);
    fe_t M = Fq::add(Fq::dbl(xx), xx);
    xyzz_t r;
    r.x = Fq::sub(Fq::sqr(M), Fq::dbl(S));
    r.y = Fq::mul_sub_mul(M, Fq::sub(S, r.x), W, p.y);
    r.zz = Fq::mul(V, p.zz);
    r.zzz = Fq::mul(W, p.zzz);
    return r;
}

// acc += q (affine); `neg` adds -q
B2R_HD void xyzz_madd(xyzz_t& acc, const affine_t& q, bool neg) {
    if (affine_is_identity(q)) return;
    fe_t qy = neg ? Fq::neg(q.y) : q.y;
    if (xyzz_is_identity(acc)) {
        acc.x = q.x;
        acc.y = qy;
        acc.zz = Fq::one();
        acc.zzz = Fq::one();
        return;
    }
    fe_t U2 = Fq::mul(q.x, acc.zz);
    fe_t S2 = Fq::mul(qy, acc.zzz);
    fe_t P = Fq::sub(U2, acc.x);
    fe_t R = Fq::sub(S2, acc.y);
    if (Fq::is_zero(P)) {
        if (Fq::is_zero(R)) {
            affine_t t;
            t.x = q.x;
            t.y = qy;
            acc = xyzz_double_affine(t);
        } else {
            acc = xyzz_identity();
        }
        return;
    }
    fe_t PP = Fq::sqr(P);
    fe_t PPP = Fq::mul(P, PP);
    fe_t Q = Fq::mul(acc.x, PP);
    fe_t X3 = Fq::sub(Fq::sub(Fq::sqr(R), PPP), Fq::dbl(Q));
    fe_t Y3 = Fq::mul_sub_mul(R, Fq::sub(Q, X3), acc.y, PPP);   // one reduction for both products
    acc.x = X3;
    acc.y = Y3;
    acc.zz = Fq::mul(acc.zz, PP);
    acc.zzz = Fq::mul(acc.zzz, PPP);
}

// acc += q (XYZZ)
B2R_HD void xyzz_add(xyzz_t& acc, const xyzz_t& q) {
    if (xyzz_is_identity(q)) return;
    if (xyzz_is_identity(acc)) {
        acc = q;
        return;
    }
    fe_t U1 = Fq::mul(acc.x, q.zz);
    fe_t U2 = Fq::mul(q.x, acc.zz);
    fe_t S1 = Fq::mul(acc.y, q.zzz);
    fe_t S2 = Fq::mul(q.y, acc.zzz);
    fe_t P = Fq::sub(U2, U1);
    fe_t R = Fq::sub(S2, S1);
    if (Fq::is_zero(P)) {
        if (Fq::is_zero(R)) {
            acc = xyzz_double(acc);
        } else {
            acc = xyzz_identity();
        }
        return;
    }
    fe_t PP = Fq::sqr(P);
    fe_t PPP = Fq::mul(P, PP);
    fe_t Q = Fq::mul(U1, PP);
    fe_t X3 = Fq::sub(Fq::sub(Fq::sqr(R), PPP), Fq::dbl(Q));
    fe_t Y3 = Fq::mul_sub_mul(R, Fq::sub(Q, X3), S1, PPP);
    acc.x = X3;
    acc.y = Y3;
    acc.zz = Fq::mul(Fq::mul(acc.zz, q.zz), PP);
    acc.zzz = Fq::mul(Fq::mul(acc.zzz, q.zzz), PPP);
}

// The P + P case of the lock-step mixed addition is taken about once per 2^254 random operands (and for repeated table
// points), but its doubling is 7 field products of straight-line code inside the hot loop: 20 of the 65 KB of
// k_accum_entries.  Out of line on the device: the loop shrinks (accumulation 221.8 -> 219.7 ms per step), the rare call
// pays the ABI's argument copies.  (The same for the full addition made k_br_level1 slower: caller-saved spills.)
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ xyzz_t xyzz_double_affine_ool(affine_t p) { return xyzz_double_affine(p); }
#else
inline xyzz_t xyzz_double_affine_ool(affine_t p) { return xyzz_double_affine(p); }
#endif

// ---- lock-step variants -----------------------------------------------------------------------
// Same results as xyzz_madd / xyzz_add, written so that all lanes of a warp execute the same
// instruction stream: the general formulas run unconditionally and the identity cases are
// resolved with selects; only the P + P / P - P case (equal x) branches, and it is rare.
B2R_HD fe_t fe_select(bool c, const fe_t& a, const fe_t& b) {
    fe_t r;
    for (int i = 0; i < 8; i++) r.l[i] = c ? a.l[i] : b.l[i];
    return r;
}
B2R_HD xyzz_t xyzz_from_affine_signed(const affine_t& p, bool neg) {
    bool id = affine_is_identity(p);
    xyzz_t r;
    r.x = p.x;
    r.y = neg ? Fq::neg(p.y) : p.y;
    r.zz = id ? Fq::zero() : Fq::one();
    r.zzz = r.zz;
    return r;
}
B2R_HD void xyzz_madd_ls(xyzz_t& acc, const affine_t& q, bool neg) {
    // identity operands leave through rare (possibly divergent) branches: real base tables hold no identity points and
    // a running sum only becomes the identity through P - P.  The general case then updates acc in place, products
    // ordered so that every operand dies at its last use (no selects holding the old sum alive next to the new one).
    if (affine_is_identity(q)) return;
    fe_t qy = neg ? Fq::neg(q.y) : q.y;
    if (xyzz_is_identity(acc)) {
        acc.x = q.x;
        acc.y = qy;
        acc.zz = Fq::one();
        acc.zzz = Fq::one();
        return;
    }
    fe_t P = Fq::sub(Fq::mul(q.x, acc.zz), acc.x);
    fe_t R = Fq::sub(Fq::mul(qy, acc.zzz), acc.y);
    if (Fq::is_zero(P)) {
        if (Fq::is_zero(R)) {
            affine_t t;
            t.x = q.x;
            t.y = qy;
            acc = xyzz_double_affine_ool(t);
        } else {
            acc = xyzz_identity();
        }
        return;
    }
    fe_t PP = Fq::sqr(P);
    acc.zz = Fq::mul(acc.zz, PP);
    fe_t Q = Fq::mul(acc.x, PP);
    fe_t PPP = Fq::mul(P, PP);
    acc.zzz = Fq::mul(acc.zzz, PPP);
    acc.x = Fq::sub(Fq::sub(Fq::sqr(R), PPP), Fq::dbl(Q));
    // Y3 = R (Q - X3) - Y1 PPP with one reduction for both products
    acc.y = Fq::mul_sub_mul(R, Fq::sub(Q, acc.x), acc.y, PPP);
}
B2R_HD void xyzz_add_ls(xyzz_t& acc, const xyzz_t& q) {
    // same shape as xyzz_madd_ls: identity operands and equal x leave through branches (whole warps of empty buckets are
    // skipped by the callers' votes before they get here), the general case runs in place
    if (xyzz_is_identity(q)) return;
    if (xyzz_is_identity(acc)) {
        acc = q;
        return;
    }
    fe_t U1 = Fq::mul(acc.x, q.zz);
    fe_t S1 = Fq::mul(acc.y, q.zzz);
    fe_t P = Fq::sub(Fq::mul(q.x, acc.zz), U1);
    fe_t R = Fq::sub(Fq::mul(q.y, acc.zzz), S1);
    if (Fq::is_zero(P)) {
        if (Fq::is_zero(R)) acc = xyzz_double(acc);
        else acc = xyzz_identity();
        return;
    }
    acc.zz = Fq::mul(acc.zz, q.zz);
    acc.zzz = Fq::mul(acc.zzz, q.zzz);
    fe_t PP = Fq::sqr(P);
    acc.zz = Fq::mul(acc.zz, PP);
    fe_t Q = Fq::mul(U1, PP);
    fe_t PPP = Fq::mul(P, PP);
    acc.zzz = Fq::mul(acc.zzz, PPP);
    acc.x = Fq::sub(Fq::sub(Fq::sqr(R), PPP), Fq::dbl(Q));
    acc.y = Fq::mul_sub_mul(R, Fq::sub(Q, acc.x), S1, PPP);
}

// normalise; identity -> (0, 0)
B2R_HD affine_t xyzz_to_affine(const xyzz_t& p) {
    affine_t r;
    if (xyzz_is_identity(p)) {
        r.x = Fq::zero();
        r.y = Fq::zero();
        return r;
    }
    fe_t t = Fq::inv_vartime(Fq::mul(p.zz, p.zzz));   // binary Euclid: this sits at the end of every MSM's single-thread tail
    r.x = Fq::mul(p.x, Fq::mul(t, p.zzz));  // X / ZZ
    r.y = Fq::mul(p.y, Fq::mul(t, p.zz));   // Y / ZZZ
    return r;
}

}  // namespace b2r
