// BN254 G1 arithmetic in XYZZ coordinates over the 29-bit-limb lazy field (field29.cuh): the bucket
// accumulation / reduction kernels of msm.cu run on these.  Same formulas and the same exceptional-case
// handling as ec.cuh (EFD madd-2008-s / add-2008-s / dbl-2008-s-1, a = 0); what changes is that no
// coordinate is ever reduced, only carry-normalised, under these bounds (multiples of q):
//
//     stored / live XYZZ point:   X < 8q,  Y < 4q,  ZZ < 2q,  ZZZ < 2q      (identity: ZZ limbs all zero)
//     affine operand:             x, y < 2q                                   (identity: x = y = 0 limbs)
//
// Every product below has B(a) B(b) <= 128 as field29.cuh requires; the derivation is in the comments
// next to each line ("<Nq").  Coordinates are in the Montgomery-2^261 domain.
#pragma once
#include "../../halo2-rsa_b200/csrc/ec.cuh"
#include "field29.cuh"

namespace b2r {

using fq29 = Fq29::el;

struct affine29_t {
    fq29 x, y;
};
struct xyzz29_t {
    fq29 x, y, zz, zzz;
};
// HBM image of an XYZZ point: 36 words (9 x uint4)
struct alignas(16) xyzz29_mem_t {
    uint32_t w[36];
};

// R' mod q = 2^261 mod q (the Montgomery-2^261 image of 1) as limbs; computed at compile time from 2^256 mod q
B2R_HD constexpr Fq29::Tab fq29_one_tab() {
    // (R mod q) * 32 mod q by five modular doublings on 32-bit words
    uint32_t v[8] = {};
    for (int i = 0; i < 8; i++) v[i] = FqP::ONE(i);
    for (int d = 0; d < 5; d++) {
        uint32_t carry = 0;
        for (int i = 0; i < 8; i++) {
            uint32_t nv = (v[i] << 1) | carry;
            carry = v[i] >> 31;
            v[i] = nv;
        }
        // v < 2q < 2^255: conditional subtract
        bool ge = true;
        for (int i = 7; i >= 0; i--) {
            if (v[i] != FqP::MOD(i)) {
                ge = v[i] > FqP::MOD(i);
                break;
            }
        }
        if (ge) {
            uint64_t br = 0;
            for (int i = 0; i < 8; i++) {
                uint64_t t = (uint64_t)v[i] - FqP::MOD(i) - br;
                v[i] = (uint32_t)t;
                br = (t >> 63) & 1;
            }
        }
    }
    Fq29::Tab t = {};
    for (int i = 0; i < 9; i++) {
        const int bit = 29 * i, q = bit >> 5, s = bit & 31;
        uint64_t two = (uint64_t)v[q] | ((uint64_t)(q + 1 < 8 ? v[q + 1] : 0u) << 32);
        t.v[i] = (uint32_t)(two >> s) & Fq29::MASK;
    }
    return t;
}
B2R_HD fq29 fq29_one() {
    constexpr Fq29::Tab t = fq29_one_tab();
    fq29 r;
#pragma unroll
    for (int i = 0; i < 9; i++) r.v[i] = t.v[i];
    return r;
}

B2R_HD bool affine29_is_identity(const affine29_t& p) { return Fq29::is_zero_limbs(p.x) && Fq29::is_zero_limbs(p.y); }
B2R_HD bool xyzz29_is_identity(const xyzz29_t& p) { return Fq29::is_zero_limbs(p.zz); }
B2R_HD xyzz29_t xyzz29_identity() {
    xyzz29_t r;
    r.x = Fq29::zero();
    r.y = Fq29::zero();
    r.zz = Fq29::zero();
    r.zzz = Fq29::zero();
    return r;
}
// table entry (x R', y R' canonical, packed 8 x 32) -> limbs
B2R_HD affine29_t affine29_unpack(const affine_t& p) {
    affine29_t r;
    r.x = Fq29::unpack(p.x);
    r.y = Fq29::unpack(p.y);
    return r;
}
// a + K q - b - 2 c in one carry pass
template <uint32_t K>
B2R_HD fq29 fq29_sub_sub2(const fq29& a, const fq29& b, const fq29& c) {
    constexpr Fq29::Tab kp = Fq29::kmod_tab(K);
    int32_t t[9];
#pragma unroll
    for (int i = 0; i < 9; i++) t[i] = (int32_t)(a.v[i] + kp.v[i]) - (int32_t)b.v[i] - (int32_t)(c.v[i] << 1);
    return Fq29::normalise(t);
}
// 3 a
B2R_HD fq29 fq29_triple(const fq29& a) {
    int32_t t[9];
#pragma unroll
    for (int i = 0; i < 9; i++) t[i] = (int32_t)(a.v[i] * 3u);
    return Fq29::normalise(t);
}

B2R_HD xyzz29_t xyzz29_from_affine_signed(const affine29_t& p, bool neg) {
    const bool id = affine29_is_identity(p);
    xyzz29_t r;
    r.x = p.x;
    r.y = neg ? Fq29::neg<2>(p.y) : p.y;  // < 2q
    r.zz = id ? Fq29::zero() : fq29_one();
    r.zzz = r.zz;
    return r;
}

// 2 (x, y) for an affine non-identity point; x, y < 2q
B2R_HD xyzz29_t xyzz29_double_affine(const fq29& x, const fq29& y) {
    const fq29 U = Fq29::dbl(y);                   // <4q
    const fq29 V = Fq29::sqr(U);                   // 16/128: <2q
    const fq29 W = Fq29::mul(U, V);                // 8: <2q
    const fq29 S = Fq29::mul(x, V);                // 4: <2q
    const fq29 M = fq29_triple(Fq29::sqr(x));      // <6q
    xyzz29_t r;
    r.x = fq29_sub_sub2<4>(Fq29::sqr(M), Fq29::zero(), S);                              // MM + 4q - 2S: <6q
    r.y = Fq29::sub<2>(Fq29::mul(M, Fq29::sub<6>(S, r.x)), Fq29::mul(W, y));            // 6*8=48 ; 2*2 : <4q
    r.zz = V;
    r.zzz = W;
    return r;
}
B2R_HD xyzz29_t xyzz29_double(const xyzz29_t& p) {
    if (xyzz29_is_identity(p)) return p;
    const fq29 U = Fq29::dbl(p.y);                 // <8q
    const fq29 V = Fq29::sqr(U);                   // 64/128: <2q
    const fq29 W = Fq29::mul(U, V);                // 16: <2q
    const fq29 S = Fq29::mul(p.x, V);              // 16: <2q
    const fq29 M = fq29_triple(Fq29::sqr(p.x));    // 64/128 -> <2q ; x3: <6q
    xyzz29_t r;
    r.x = fq29_sub_sub2<4>(Fq29::sqr(M), Fq29::zero(), S);                              // <6q
    r.y = Fq29::sub<2>(Fq29::mul(M, Fq29::sub<6>(S, r.x)), Fq29::mul(W, p.y));          // 48 ; 2*4 : <4q
    r.zz = Fq29::mul(V, p.zz);
    r.zzz = Fq29::mul(W, p.zzz);
    return r;
}

// acc += +-q, lock-step form (see ec.cuh xyzz_madd_ls): the general formulas run unconditionally, identity
// operands are resolved with selects, only the equal-x case branches.
B2R_HD void xyzz29_madd_ls(xyzz29_t& acc, const affine29_t& q, bool neg) {
    const bool q_id = affine29_is_identity(q), a_id = xyzz29_is_identity(acc);
    const fq29 qy = neg ? Fq29::neg<2>(q.y) : q.y;                  // <2q
    const fq29 U2 = Fq29::mul(q.x, acc.zz);                         // 2*2: <2q
    const fq29 S2 = Fq29::mul(qy, acc.zzz);                         // <2q
    const fq29 P = Fq29::sub<8>(U2, acc.x);                         // <10q
    const fq29 R = Fq29::sub<4>(S2, acc.y);                         // <6q
    if (!q_id && !a_id && Fq29::is_zero_mod_p(P)) {
        if (Fq29::is_zero_mod_p(R)) acc = xyzz29_double_affine(q.x, qy);
        else acc = xyzz29_identity();
        return;
    }
    const fq29 PP = Fq29::sqr(P);                                   // 100/128: <2q
    const fq29 PPP = Fq29::mul(P, PP);                              // 20: <2q
    const fq29 Q = Fq29::mul(acc.x, PP);                            // 16: <2q
    fq29 X3 = fq29_sub_sub2<6>(Fq29::sqr(R), PPP, Q);               // RR + 6q - PPP - 2Q: <8q
    fq29 Y3 = Fq29::sub<2>(Fq29::mul(R, Fq29::sub<8>(Q, X3)), Fq29::mul(acc.y, PPP));  // 6*10=60 ; 4*2 : <4q
    fq29 ZZ3 = Fq29::mul(acc.zz, PP);                               // <2q
    fq29 ZZZ3 = Fq29::mul(acc.zzz, PPP);                            // <2q
    if (a_id) {
        X3 = q.x;
        Y3 = qy;
        ZZ3 = fq29_one();
        ZZZ3 = ZZ3;
    }
    if (!q_id) {
        acc.x = X3;
        acc.y = Y3;
        acc.zz = ZZ3;
        acc.zzz = ZZZ3;
    }
}

// acc += q (both XYZZ), lock-step form
B2R_HD void xyzz29_add_ls(xyzz29_t& acc, const xyzz29_t& q) {
    const bool q_id = xyzz29_is_identity(q), a_id = xyzz29_is_identity(acc);
    const fq29 U1 = Fq29::mul(acc.x, q.zz);                         // 8*2: <2q
    const fq29 U2 = Fq29::mul(q.x, acc.zz);
    const fq29 S1 = Fq29::mul(acc.y, q.zzz);                        // 4*2
    const fq29 S2 = Fq29::mul(q.y, acc.zzz);
    const fq29 P = Fq29::sub<2>(U2, U1);                            // <4q
    const fq29 R = Fq29::sub<2>(S2, S1);                            // <4q
    if (!q_id && !a_id && Fq29::is_zero_mod_p(P)) {
        if (Fq29::is_zero_mod_p(R)) acc = xyzz29_double(acc);
        else acc = xyzz29_identity();
        return;
    }
    const fq29 PP = Fq29::sqr(P);                                   // 16/128
    const fq29 PPP = Fq29::mul(P, PP);                              // 8
    const fq29 Q = Fq29::mul(U1, PP);                               // 4
    fq29 X3 = fq29_sub_sub2<6>(Fq29::sqr(R), PPP, Q);               // <8q
    fq29 Y3 = Fq29::sub<2>(Fq29::mul(R, Fq29::sub<8>(Q, X3)), Fq29::mul(S1, PPP));  // 4*10 ; 4 : <4q
    fq29 ZZ3 = Fq29::mul(Fq29::mul(acc.zz, q.zz), PP);
    fq29 ZZZ3 = Fq29::mul(Fq29::mul(acc.zzz, q.zzz), PPP);
    if (a_id) {
        X3 = q.x;
        Y3 = q.y;
        ZZ3 = q.zz;
        ZZZ3 = q.zzz;
    }
    if (!q_id) {
        acc.x = X3;
        acc.y = Y3;
        acc.zz = ZZ3;
        acc.zzz = ZZZ3;
    }
}

// a^(q-2) in the Montgomery-2^261 domain (a < 2q); inv(0) = 0
B2R_HD fq29 fq29_inv(const fq29& a) {
    uint32_t e[8];
    for (int i = 0; i < 8; i++) e[i] = FqP::MOD(i);
    e[0] -= 2u;
    fq29 acc = fq29_one();
    bool started = false;
    for (int w = 7; w >= 0; w--) {
        for (int bit = 31; bit >= 0; bit--) {
            if (started) acc = Fq29::sqr(acc);
            if ((e[w] >> bit) & 1u) {
                acc = started ? Fq29::mul(acc, a) : a;
                started = true;
            }
        }
    }
    return acc;
}
// normalise to the ABI's affine point: canonical Montgomery-2^256 coordinates, identity -> (0, 0)
B2R_HD affine_t xyzz29_to_affine256(const xyzz29_t& p) {
    affine_t r;
    if (xyzz29_is_identity(p)) {
        r.x = Fq::zero();
        r.y = Fq::zero();
        return r;
    }
    const fq29 t = fq29_inv(Fq29::mul(p.zz, p.zzz));
    r.x = Fq29::to_mont256(Fq29::mul(p.x, Fq29::mul(t, p.zzz)));  // X / ZZ
    r.y = Fq29::to_mont256(Fq29::mul(p.y, Fq29::mul(t, p.zz)));   // Y / ZZZ
    return r;
}

}  // namespace b2r
