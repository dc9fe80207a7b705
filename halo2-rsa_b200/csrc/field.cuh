// BN254 Fr / Fq arithmetic for sm_100a: 8 x 32-bit limbs, Montgomery form (R = 2^256).
//
// Memory format is halo2curves' bn256::{Fr,Fq}: 4 x u64 little-endian limbs in
// Montgomery form == 8 x u32 little-endian limbs, so a Rust &[Fr] can be handed to the
// kernels as a raw pointer.
//
// Montgomery multiplication is an interleaved (CIOS-style) product/reduction held in two
// 64-bit-lane accumulators ("primary" P at word offset 0 and "secondary" S at word
// offset 1, T = P + 2^32 * S).  Every 32x32->64 product is one mad.lo.cc/madc.hi.cc pair
// that ptxas fuses into a single IMAD.WIDE.U32 with carry-in/out, so a full multiply is
// ~130 wide MADs + ~50 carry/select instructions and no spills.
//
// The same source compiles for the host (plain C emulation of the row primitives): that
// is how the carry logic is unit-tested in this GPU-less container
// (tests/host/field_host_test.cpp).  The host build is test scaffolding for the device
// algorithm, not a CPU fallback: no product path calls it.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B2R_HD __host__ __device__ __forceinline__
#define B2R_D __device__ __forceinline__
#else
#define B2R_HD inline
#define B2R_D inline
#endif

// 1 = Fr's m * p rows without a multiplier for the low word (below): bit-exact, measured SLOWER on B200 (NTT passes 159 ->
// 177 ms, quotient 56 -> 66 ms per step: the passes are as short of ALU / issue slots as of multiplier cycles), default off
#ifndef B2R_SPARSE_P0
#define B2R_SPARSE_P0 0
#endif

namespace b2r {

struct alignas(16) fe_t {
    uint32_t l[8];
};

// ---- field parameter packs --------------------------------------------------------------
struct FrP {
    // r = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
    B2R_HD static constexpr uint32_t MOD(int i) {
        constexpr uint32_t v[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                        0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return v[i];
    }
    static constexpr uint32_t N0INV = 0xefffffffu;  // -r^-1 mod 2^32
    // R mod r
    B2R_HD static constexpr uint32_t ONE(int i) {
        constexpr uint32_t v[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                                        0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return v[i];
    }
    // R^2 mod r
    B2R_HD static constexpr uint32_t R2(int i) {
        constexpr uint32_t v[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                                       0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return v[i];
    }
};

struct FqP {
    // q = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
    B2R_HD static constexpr uint32_t MOD(int i) {
        constexpr uint32_t v[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                                        0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return v[i];
    }
    static constexpr uint32_t N0INV = 0xe4866389u;  // -q^-1 mod 2^32
    // R mod q
    B2R_HD static constexpr uint32_t ONE(int i) {
        constexpr uint32_t v[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                                        0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return v[i];
    }
    // R^2 mod q
    B2R_HD static constexpr uint32_t R2(int i) {
        constexpr uint32_t v[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                                       0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return v[i];
    }
};

// ---- row primitives ----------------------------------------------------------------------
// A "row" is four 64-bit lanes: lane l covers words (2l, 2l+1).  `a` is indexed with
// stride 2 (a[0], a[2], a[4], a[6]); pass &x[0] for the even limbs, &x[1] for the odd.

// X[0..7] = lanes a[2l] * w
B2R_HD void row_mul(uint32_t* X, const uint32_t* a, uint32_t w) {
#if defined(__CUDA_ARCH__)
    asm("mul.lo.u32 %0, %8, %12; mul.hi.u32 %1, %8, %12;"
        "mul.lo.u32 %2, %9, %12; mul.hi.u32 %3, %9, %12;"
        "mul.lo.u32 %4, %10, %12; mul.hi.u32 %5, %10, %12;"
        "mul.lo.u32 %6, %11, %12; mul.hi.u32 %7, %11, %12;"
        : "=&r"(X[0]), "=&r"(X[1]), "=&r"(X[2]), "=&r"(X[3]), "=&r"(X[4]), "=&r"(X[5]),
          "=&r"(X[6]), "=&r"(X[7])
        : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(w));
#else
    for (int l = 0; l < 4; l++) {
        uint64_t p = (uint64_t)a[2 * l] * w;
        X[2 * l] = (uint32_t)p;
        X[2 * l + 1] = (uint32_t)(p >> 32);
    }
#endif
}

// X[0..7] += lanes a[2l] * w ; returns carry out of word 7 (0 or 1)
B2R_HD uint32_t row_mad(uint32_t* X, const uint32_t* a, uint32_t w) {
    uint32_t c;
#if defined(__CUDA_ARCH__)
    asm("mad.lo.cc.u32 %0, %9, %13, %0; madc.hi.cc.u32 %1, %9, %13, %1;"
        "madc.lo.cc.u32 %2, %10, %13, %2; madc.hi.cc.u32 %3, %10, %13, %3;"
        "madc.lo.cc.u32 %4, %11, %13, %4; madc.hi.cc.u32 %5, %11, %13, %5;"
        "madc.lo.cc.u32 %6, %12, %13, %6; madc.hi.cc.u32 %7, %12, %13, %7;"
        "addc.u32 %8, 0, 0;"
        : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]),
          "+r"(X[7]), "=r"(c)
        : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(w));
#else
    uint64_t cy = 0;
    for (int l = 0; l < 4; l++) {
        uint64_t x = (uint64_t)X[2 * l] | ((uint64_t)X[2 * l + 1] << 32);
        unsigned __int128 t = (unsigned __int128)a[2 * l] * w + x + cy;
        X[2 * l] = (uint32_t)t;
        X[2 * l + 1] = (uint32_t)(t >> 32);
        cy = (uint64_t)(t >> 64);
    }
    c = (uint32_t)cy;
#endif
    return c;
}

// same, carry out of word 7 is known to be zero by a value bound (not materialised)
B2R_HD void row_mad_nc(uint32_t* X, const uint32_t* a, uint32_t w) {
#if defined(__CUDA_ARCH__)
    asm("mad.lo.cc.u32 %0, %8, %12, %0; madc.hi.cc.u32 %1, %8, %12, %1;"
        "madc.lo.cc.u32 %2, %9, %12, %2; madc.hi.cc.u32 %3, %9, %12, %3;"
        "madc.lo.cc.u32 %4, %10, %12, %4; madc.hi.cc.u32 %5, %10, %12, %5;"
        "madc.lo.cc.u32 %6, %11, %12, %6; madc.hi.u32 %7, %11, %12, %7;"
        : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]),
          "+r"(X[7])
        : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(w));
#else
    (void)row_mad(X, a, w);
#endif
}

// The shift step.  In:  P[0..8] (P[0] == 0 mod 2^32 already consumed), S[0..7].
// Computes  S[0] += P[1]  (carry c), then  Q = (P >> 64) + lanes a[2l]*w + c  where
// (P >> 64) = words P[2..8] followed by a zero word.  Q is written over P[2..9].
B2R_HD void row_shift_mad(uint32_t* P /*10 words*/, uint32_t* S0, const uint32_t* a, uint32_t w) {
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %8, %8, %9;"
        "madc.lo.cc.u32 %0, %10, %14, %0; madc.hi.cc.u32 %1, %10, %14, %1;"
        "madc.lo.cc.u32 %2, %11, %14, %2; madc.hi.cc.u32 %3, %11, %14, %3;"
        "madc.lo.cc.u32 %4, %12, %14, %4; madc.hi.cc.u32 %5, %12, %14, %5;"
        "madc.lo.cc.u32 %6, %13, %14, %6; madc.hi.u32 %7, %13, %14, 0;"
        : "+r"(P[2]), "+r"(P[3]), "+r"(P[4]), "+r"(P[5]), "+r"(P[6]), "+r"(P[7]), "+r"(P[8]),
          "=&r"(P[9]), "+r"(*S0)
        : "r"(P[1]), "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(w));
#else
    uint64_t s = (uint64_t)*S0 + P[1];
    *S0 = (uint32_t)s;
    uint64_t cy = s >> 32;
    P[9] = 0;
    for (int l = 0; l < 4; l++) {
        uint64_t x = (uint64_t)P[2 + 2 * l] | ((uint64_t)P[3 + 2 * l] << 32);
        unsigned __int128 t = (unsigned __int128)a[2 * l] * w + x + cy;
        P[2 + 2 * l] = (uint32_t)t;
        P[3 + 2 * l] = (uint32_t)(t >> 32);
        cy = (uint64_t)(t >> 64);
    }
#endif
}

// ---- squaring primitives -------------------------------------------------------------------
// X[0..2K-1] += K lanes a[2l] * w, carry out added into X[2K] (which only ever holds earlier carries: no ripple)
template <int K>
B2R_HD void lanes_mad_c(uint32_t* X, const uint32_t* a, uint32_t w) {
#if defined(__CUDA_ARCH__)
    if constexpr (K == 1) {
        asm("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;"
            : "+r"(X[0]), "+r"(X[1]), "+r"(X[2])
            : "r"(a[0]), "r"(w));
    } else if constexpr (K == 2) {
        asm("mad.lo.cc.u32 %0, %5, %7, %0; madc.hi.cc.u32 %1, %5, %7, %1;"
            "madc.lo.cc.u32 %2, %6, %7, %2; madc.hi.cc.u32 %3, %6, %7, %3; addc.u32 %4, %4, 0;"
            : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4])
            : "r"(a[0]), "r"(a[2]), "r"(w));
    } else if constexpr (K == 3) {
        asm("mad.lo.cc.u32 %0, %7, %10, %0; madc.hi.cc.u32 %1, %7, %10, %1;"
            "madc.lo.cc.u32 %2, %8, %10, %2; madc.hi.cc.u32 %3, %8, %10, %3;"
            "madc.lo.cc.u32 %4, %9, %10, %4; madc.hi.cc.u32 %5, %9, %10, %5; addc.u32 %6, %6, 0;"
            : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6])
            : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(w));
    } else {
        asm("mad.lo.cc.u32 %0, %9, %13, %0; madc.hi.cc.u32 %1, %9, %13, %1;"
            "madc.lo.cc.u32 %2, %10, %13, %2; madc.hi.cc.u32 %3, %10, %13, %3;"
            "madc.lo.cc.u32 %4, %11, %13, %4; madc.hi.cc.u32 %5, %11, %13, %5;"
            "madc.lo.cc.u32 %6, %12, %13, %6; madc.hi.cc.u32 %7, %12, %13, %7; addc.u32 %8, %8, 0;"
            : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]), "+r"(X[7]), "+r"(X[8])
            : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(w));
    }
#else
    uint64_t cy = 0;
    for (int l = 0; l < K; l++) {
        uint64_t x = (uint64_t)X[2 * l] | ((uint64_t)X[2 * l + 1] << 32);
        unsigned __int128 t = (unsigned __int128)a[2 * l] * w + x + cy;
        X[2 * l] = (uint32_t)t;
        X[2 * l + 1] = (uint32_t)(t >> 32);
        cy = (uint64_t)(t >> 64);
    }
    X[2 * K] += (uint32_t)cy;
#endif
}

// 8-word add / sub with carry chains
B2R_HD uint32_t add8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t c;
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %0, %9, %17; addc.cc.u32 %1, %10, %18; addc.cc.u32 %2, %11, %19;"
        "addc.cc.u32 %3, %12, %20; addc.cc.u32 %4, %13, %21; addc.cc.u32 %5, %14, %22;"
        "addc.cc.u32 %6, %15, %23; addc.cc.u32 %7, %16, %24; addc.u32 %8, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]),
          "=&r"(r[6]), "=&r"(r[7]), "=&r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]),
          "r"(a[7]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]),
          "r"(b[6]), "r"(b[7]));
#else
    uint64_t cy = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t t = (uint64_t)a[i] + b[i] + cy;
        r[i] = (uint32_t)t;
        cy = t >> 32;
    }
    c = (uint32_t)cy;
#endif
    return c;
}

// r = a - b, returns borrow (1 if a < b)
B2R_HD uint32_t sub8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t bw;
#if defined(__CUDA_ARCH__)
    asm("sub.cc.u32 %0, %9, %17; subc.cc.u32 %1, %10, %18; subc.cc.u32 %2, %11, %19;"
        "subc.cc.u32 %3, %12, %20; subc.cc.u32 %4, %13, %21; subc.cc.u32 %5, %14, %22;"
        "subc.cc.u32 %6, %15, %23; subc.cc.u32 %7, %16, %24; subc.u32 %8, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]),
          "=&r"(r[6]), "=&r"(r[7]), "=&r"(bw)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]),
          "r"(a[7]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]),
          "r"(b[6]), "r"(b[7]));
    bw &= 1u;
#else
    int64_t br = 0;
    for (int i = 0; i < 8; i++) {
        int64_t t = (int64_t)a[i] - b[i] - br;
        r[i] = (uint32_t)t;
        br = (t < 0) ? 1 : 0;
    }
    bw = (uint32_t)br;
#endif
    return bw;
}

// ---- field ops -----------------------------------------------------------------------------
template <class P>
struct Field {
    B2R_HD static fe_t zero() {
        fe_t r;
        for (int i = 0; i < 8; i++) r.l[i] = 0;
        return r;
    }
    B2R_HD static fe_t one() {
        fe_t r;
        for (int i = 0; i < 8; i++) r.l[i] = P::ONE(i);
        return r;
    }
    B2R_HD static fe_t r2() {
        fe_t r;
        for (int i = 0; i < 8; i++) r.l[i] = P::R2(i);
        return r;
    }
    B2R_HD static bool is_zero(const fe_t& a) {
        uint32_t o = 0;
        for (int i = 0; i < 8; i++) o |= a.l[i];
        return o == 0;
    }
    B2R_HD static bool eq(const fe_t& a, const fe_t& b) {
        uint32_t o = 0;
        for (int i = 0; i < 8; i++) o |= a.l[i] ^ b.l[i];
        return o == 0;
    }
    // r = (x >= p) ? x - p : x      for x < 2p
    B2R_HD static void final_sub(uint32_t* x) {
        uint32_t m[8], t[8];
        for (int i = 0; i < 8; i++) m[i] = P::MOD(i);
        uint32_t bw = sub8(t, x, m);
        for (int i = 0; i < 8; i++) x[i] = bw ? x[i] : t[i];
    }
    B2R_HD static fe_t add(const fe_t& a, const fe_t& b) {
        fe_t r;
        add8(r.l, a.l, b.l);  // < 2p < 2^255: no carry out
        final_sub(r.l);
        return r;
    }
    B2R_HD static fe_t dbl(const fe_t& a) { return add(a, a); }
    B2R_HD static fe_t sub(const fe_t& a, const fe_t& b) {
        fe_t r;
        uint32_t m[8], t[8];
        for (int i = 0; i < 8; i++) m[i] = P::MOD(i);
        uint32_t bw = sub8(r.l, a.l, b.l);
        add8(t, r.l, m);
        for (int i = 0; i < 8; i++) r.l[i] = bw ? t[i] : r.l[i];
        return r;
    }
    B2R_HD static fe_t neg(const fe_t& a) {
        fe_t r;
        uint32_t m[8];
        for (int i = 0; i < 8; i++) m[i] = P::MOD(i);
        sub8(r.l, m, a.l);
        bool z = is_zero(a);
        for (int i = 0; i < 8; i++) r.l[i] = z ? 0u : r.l[i];
        return r;
    }

    // ---- the m * p rows -------------------------------------------------------------------------------------
    // m = -t0 / p mod 2^32.  Fr's modulus ends in 0xf0000001 = 2^32 - 2^28 + 1, so n0' = 0xefffffff = -(2^28 + 1) and
    // m = -(t0 + (t0 << 28)): two ALU operations instead of a multiply on the (binding) fma pipe.
    static constexpr bool SPARSE_P0 = B2R_SPARSE_P0 && P::MOD(0) == 0xf0000001u;
    B2R_HD static uint32_t mont_m(uint32_t t0) {
        if constexpr (SPARSE_P0) return 0u - t0 - (t0 << 28);
        else return t0 * P::N0INV;
    }
    // X[0..7] += lanes (p0, p2, p4, p6) * mi, returns the carry out of word 7.  X[0] + lo(p0 * mi) is 0 mod 2^32 by the
    // choice of mi and never read again, so for Fr the first lane needs no multiplier: the carry into word 1 is
    // (X[0] != 0) and hi(mi * (2^32 - 2^28 + 1)) = mi - ceil(mi (2^28 - 1) / 2^32) comes from shifts and subtractions -
    // one wide MAD per row (8 of the 128 of a product) moved to the ALU pipe.  X[0] is left stale.
    B2R_HD static uint32_t mod_row_even(uint32_t* X, const uint32_t* m, uint32_t mi) {
        if constexpr (!SPARSE_P0) {
            return row_mad(X, m, mi);
        } else {
            uint32_t c;
#if defined(__CUDA_ARCH__)
            uint32_t ylo, yhi;
            asm("sub.cc.u32 %0, %2, %3; subc.u32 %1, %4, 0;" : "=r"(ylo), "=r"(yhi) : "r"(mi << 28), "r"(mi), "r"(mi >> 4));
            const uint32_t h = mi - yhi - (ylo != 0u ? 1u : 0u);
            asm("add.cc.u32 %7, %8, 0xffffffff;"
                "addc.cc.u32 %0, %0, %9;"
                "madc.lo.cc.u32 %1, %10, %13, %1; madc.hi.cc.u32 %2, %10, %13, %2;"
                "madc.lo.cc.u32 %3, %11, %13, %3; madc.hi.cc.u32 %4, %11, %13, %4;"
                "madc.lo.cc.u32 %5, %12, %13, %5; madc.hi.cc.u32 %6, %12, %13, %6;"
                "addc.u32 %7, 0, 0;"
                : "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]), "+r"(X[7]), "=&r"(c)
                : "r"(X[0]), "r"(h), "r"(m[2]), "r"(m[4]), "r"(m[6]), "r"(mi));
#else
            const uint64_t y = ((uint64_t)mi << 28) - mi;   // mi (2^28 - 1)
            const uint32_t h = mi - (uint32_t)(y >> 32) - ((uint32_t)y != 0u ? 1u : 0u);
            uint64_t cy = X[0] != 0u ? 1u : 0u;
            uint64_t t1 = (uint64_t)X[1] + h + cy;
            X[1] = (uint32_t)t1;
            cy = t1 >> 32;
            for (int l = 1; l < 4; l++) {
                uint64_t x = (uint64_t)X[2 * l] | ((uint64_t)X[2 * l + 1] << 32);
                unsigned __int128 t = (unsigned __int128)m[2 * l] * mi + x + cy;
                X[2 * l] = (uint32_t)t;
                X[2 * l + 1] = (uint32_t)(t >> 32);
                cy = (uint64_t)(t >> 64);
            }
            c = (uint32_t)cy;
#endif
            return c;
        }
    }

    // Montgomery product a*b*R^-1 mod p, inputs and output fully reduced (< p).
    // Invariant per iteration (T = Pw + 2^32 * Sw): T < a + p < 2^255 at iteration start,
    // so the S chains never carry out of 8 words and Pw needs one carry word (<= 2).
    B2R_HD static fe_t mul(const fe_t& a, const fe_t& b) {
        uint32_t m[8];
        for (int i = 0; i < 8; i++) m[i] = P::MOD(i);
        uint32_t Pw[10], Sw[10];
        // i = 0
        row_mul(Pw, &a.l[0], b.l[0]);
        row_mul(Sw, &a.l[1], b.l[0]);
        uint32_t mi = mont_m(Pw[0]);
        Pw[8] = mod_row_even(Pw, &m[0], mi);
        row_mad_nc(Sw, &m[1], mi);
#pragma unroll
        for (int i = 1; i < 8; i++) {
            // T /= 2^32 and T += a * b[i]:   newP = S + P[1],  newS = (P >> 64) + odd(a)*b[i] + carry
            row_shift_mad(Pw, &Sw[0], &a.l[1], b.l[i]);
            uint32_t nP[10], nS[10];
            for (int k = 0; k < 8; k++) nS[k] = Pw[k + 2];
            for (int k = 0; k < 8; k++) nP[k] = Sw[k];
            nP[8] = row_mad(nP, &a.l[0], b.l[i]);
            mi = mont_m(nP[0]);
            nP[8] += mod_row_even(nP, &m[0], mi);
            row_mad_nc(nS, &m[1], mi);
            for (int k = 0; k < 9; k++) Pw[k] = nP[k];
            for (int k = 0; k < 8; k++) Sw[k] = nS[k];
        }
        // final shift: result = S + P[1] + 2^32 * (P >> 64)  ==  S + (P >> 32)
        fe_t r;
        add8(r.l, Sw, &Pw[1]);
        final_sub(r.l);
        return r;
    }
    // a*b + c*d (Montgomery, fully reduced) with ONE reduction: the rows of both products go into the same pair of
    // accumulators before the m * p rows of the iteration - 192 wide MADs instead of the 256 of two products.
    // Inputs < p (c may equal p: see mul_sub_mul).  Bounds: at iteration start T < a + c + p <= 3p < 2^256 (p < 2^253.6),
    // inside an iteration T + a b_i + c d_i + p m < (a + c + p) 2^32 < 2^288, so the S lanes (words 1..8) never carry out
    // and the P lanes need their carry word; the result (a b + c d + M p) / 2^256 < p (2p / 2^256 + 1) < 1.38 p: one
    // conditional subtraction.
    B2R_HD static fe_t mul_add_mul(const fe_t& a, const fe_t& b, const fe_t& c, const fe_t& d) {
        uint32_t m[8];
        for (int i = 0; i < 8; i++) m[i] = P::MOD(i);
        uint32_t Pw[10], Sw[10];
        row_mul(Pw, &a.l[0], b.l[0]);
        row_mul(Sw, &a.l[1], b.l[0]);
        Pw[8] = row_mad(Pw, &c.l[0], d.l[0]);
        row_mad_nc(Sw, &c.l[1], d.l[0]);
        uint32_t mi = mont_m(Pw[0]);
        Pw[8] += mod_row_even(Pw, &m[0], mi);
        row_mad_nc(Sw, &m[1], mi);
#pragma unroll
        for (int i = 1; i < 8; i++) {
            row_shift_mad(Pw, &Sw[0], &a.l[1], b.l[i]);
            uint32_t nP[10], nS[10];
            for (int k = 0; k < 8; k++) nS[k] = Pw[k + 2];
            for (int k = 0; k < 8; k++) nP[k] = Sw[k];
            nP[8] = row_mad(nP, &a.l[0], b.l[i]);
            nP[8] += row_mad(nP, &c.l[0], d.l[i]);
            row_mad_nc(nS, &c.l[1], d.l[i]);
            mi = mont_m(nP[0]);
            nP[8] += mod_row_even(nP, &m[0], mi);
            row_mad_nc(nS, &m[1], mi);
            for (int k = 0; k < 9; k++) Pw[k] = nP[k];
            for (int k = 0; k < 8; k++) Sw[k] = nS[k];
        }
        fe_t r;
        add8(r.l, Sw, &Pw[1]);
        final_sub(r.l);
        return r;
    }
    // a*b - c*d = a*b + (p - c)*d; c = 0 gives the multiplicand p (adds p*d: still in the bounds above)
    B2R_HD static fe_t mul_sub_mul(const fe_t& a, const fe_t& b, const fe_t& c, const fe_t& d) {
        fe_t nc;
        uint32_t m[8];
        for (int i = 0; i < 8; i++) m[i] = P::MOD(i);
        sub8(nc.l, m, c.l);
        return mul_add_mul(a, b, nc, d);
    }
    // sum of four products with one reduction (320 wide MADs instead of 512).  T < a0 + a1 + a2 + a3 + p <= 5p < 2^256
    // at iteration start, < 2^288 inside; result < p (4p / 2^256 + 1) < 1.76 p.
    B2R_HD static fe_t dot4(const fe_t& a0, const fe_t& b0, const fe_t& a1, const fe_t& b1, const fe_t& a2, const fe_t& b2, const fe_t& a3,
                            const fe_t& b3) {
        uint32_t m[8];
        for (int i = 0; i < 8; i++) m[i] = P::MOD(i);
        uint32_t Pw[10], Sw[10];
        row_mul(Pw, &a0.l[0], b0.l[0]);
        row_mul(Sw, &a0.l[1], b0.l[0]);
        Pw[8] = row_mad(Pw, &a1.l[0], b1.l[0]);
        row_mad_nc(Sw, &a1.l[1], b1.l[0]);
        Pw[8] += row_mad(Pw, &a2.l[0], b2.l[0]);
        row_mad_nc(Sw, &a2.l[1], b2.l[0]);
        Pw[8] += row_mad(Pw, &a3.l[0], b3.l[0]);
        row_mad_nc(Sw, &a3.l[1], b3.l[0]);
        uint32_t mi = mont_m(Pw[0]);
        Pw[8] += mod_row_even(Pw, &m[0], mi);
        row_mad_nc(Sw, &m[1], mi);
#pragma unroll
        for (int i = 1; i < 8; i++) {
            row_shift_mad(Pw, &Sw[0], &a0.l[1], b0.l[i]);
            uint32_t nP[10], nS[10];
            for (int k = 0; k < 8; k++) nS[k] = Pw[k + 2];
            for (int k = 0; k < 8; k++) nP[k] = Sw[k];
            nP[8] = row_mad(nP, &a0.l[0], b0.l[i]);
            nP[8] += row_mad(nP, &a1.l[0], b1.l[i]);
            row_mad_nc(nS, &a1.l[1], b1.l[i]);
            nP[8] += row_mad(nP, &a2.l[0], b2.l[i]);
            row_mad_nc(nS, &a2.l[1], b2.l[i]);
            nP[8] += row_mad(nP, &a3.l[0], b3.l[i]);
            row_mad_nc(nS, &a3.l[1], b3.l[i]);
            mi = mont_m(nP[0]);
            nP[8] += mod_row_even(nP, &m[0], mi);
            row_mad_nc(nS, &m[1], mi);
            for (int k = 0; k < 9; k++) Pw[k] = nP[k];
            for (int k = 0; k < 8; k++) Sw[k] = nS[k];
        }
        fe_t r;
        add8(r.l, Sw, &Pw[1]);
        final_sub(r.l);
        return r;
    }
    // iteration I >= 1 of sqr(): shift fused with the odd lanes of m * p, even lanes of m * p, then row I of the square
    template <int I>
    B2R_HD static void sqr_step(uint32_t* Pw, uint32_t* Sw, const uint32_t* a, const uint32_t* D, const uint32_t* m) {
        const uint32_t mi = mont_m(Sw[0] + Pw[1]);
        row_shift_mad(Pw, &Sw[0], &m[1], mi);
        uint32_t nP[10], nS[10];
        for (int k = 0; k < 8; k++) nS[k] = Pw[k + 2];
        for (int k = 0; k < 8; k++) nP[k] = Sw[k];
        nS[8] = 0;
        nP[8] = mod_row_even(nP, &m[0], mi);
        uint32_t V[9];
        for (int k = 0; k < 8; k++) V[k] = D[k];
        V[8] = 0;
        V[I] = a[I];
        if constexpr (I < 7) V[I + 1] = D[I + 1] & ~1u;
        constexpr int le = (I + 1) / 2, lo = I / 2;   // first even-position lane (P) / odd-position lane (S) of the row
        if constexpr (le < 4) lanes_mad_c<4 - le>(&nP[2 * le], &V[2 * le], a[I]);
        lanes_mad_c<4 - lo>(&nS[2 * lo], &V[2 * lo + 1], a[I]);   // carry word nS[8] stays 0: the S lanes never carry out (bound)
        for (int k = 0; k < 9; k++) Pw[k] = nP[k];
        for (int k = 0; k < 8; k++) Sw[k] = nS[k];
    }
    // Montgomery square, interleaved like mul(): iteration i adds a_i * (a_i, 2a_{i+1}, .., 2a_7) at relative word
    // positions i..7 - the doubled cross terms come from the words of D = 2a (a < 2^254), with the bit that a_i shifted
    // into D_{i+1} cleared - and then the m * p row: 36 + 64 wide MADs instead of 128, in the register footprint of
    // mul().  (The first version gathered all 16 words of a^2 before reducing: two 16-word accumulators, which spilled
    // wherever a point addition wanted R^2 next to six other live values.)  Row i >= 1 does not touch relative word 0,
    // so m is known before the row is added and the shift is fused with the odd lanes of m * p.
    // Bounds as in mul(): the window stays below (2a + p) 2^32 < 2^288, the result below 1.19 p.
    B2R_HD static fe_t sqr(const fe_t& a) {
        uint32_t m[8];
        for (int i = 0; i < 8; i++) m[i] = P::MOD(i);
        uint32_t D[8];
        D[0] = a.l[0] << 1;
        for (int k = 1; k < 8; k++) D[k] = (a.l[k] << 1) | (a.l[k - 1] >> 31);
        uint32_t Pw[10], Sw[10];
        {
            uint32_t V[8];
            V[0] = a.l[0];
            V[1] = D[1] & ~1u;
            for (int k = 2; k < 8; k++) V[k] = D[k];
            row_mul(Pw, &V[0], a.l[0]);
            row_mul(Sw, &V[1], a.l[0]);
        }
        uint32_t mi = mont_m(Pw[0]);
        Pw[8] = mod_row_even(Pw, &m[0], mi);
        row_mad_nc(Sw, &m[1], mi);
        sqr_step<1>(Pw, Sw, a.l, D, m);
        sqr_step<2>(Pw, Sw, a.l, D, m);
        sqr_step<3>(Pw, Sw, a.l, D, m);
        sqr_step<4>(Pw, Sw, a.l, D, m);
        sqr_step<5>(Pw, Sw, a.l, D, m);
        sqr_step<6>(Pw, Sw, a.l, D, m);
        sqr_step<7>(Pw, Sw, a.l, D, m);
        fe_t r;
        add8(r.l, Sw, &Pw[1]);
        final_sub(r.l);
        return r;
    }

    B2R_HD static fe_t to_mont(const fe_t& a) { return mul(a, r2()); }
    B2R_HD static fe_t from_mont(const fe_t& a) {
        fe_t o = zero();
        o.l[0] = 1;
        return mul(a, o);
    }

    // a^e, e given as 8 little-endian words (not secret: variable time)
    B2R_HD static fe_t pow(const fe_t& a, const uint32_t* e) {
        fe_t acc = one();
        bool started = false;
        for (int w = 7; w >= 0; w--) {
            for (int bit = 31; bit >= 0; bit--) {
                if (started) acc = sqr(acc);
                if ((e[w] >> bit) & 1u) {
                    acc = started ? mul(acc, a) : a;
                    started = true;
                }
            }
        }
        return acc;
    }
    // a^-1 = a^(p-2) (Fermat); inv(0) = 0
    B2R_HD static fe_t inv(const fe_t& a) {
        uint32_t e[8];
        for (int i = 0; i < 8; i++) e[i] = P::MOD(i);
        e[0] -= 2u;  // both moduli end in ...01 / ...47: no borrow
        return pow(a, e);
    }
    // The same inverse by the binary extended Euclid algorithm: ~500 iterations of 256-bit shifts / subtractions on the
    // ALU instead of 254 squarings + ~127 products on the multiplier.  One thread's Fermat chain is ~380 dependent
    // Montgomery products (0.2 ms at single-warp latency: the tail of every MSM's bucket reduction); this is ~10x
    // shorter.  Variable time (public data only).  Input / output in Montgomery form: (aR)^-1 * R^3 * R^-1 = a^-1 R.
    B2R_HD static fe_t inv_vartime(const fe_t& a) {
        if (is_zero(a)) return a;
        uint32_t u[8], v[8], x1[8], x2[8], m[8], t[8];
        for (int i = 0; i < 8; i++) { u[i] = a.l[i]; v[i] = m[i] = P::MOD(i); x1[i] = x2[i] = 0; }
        x1[0] = 1;
        auto is_one = [](const uint32_t* w) {
            uint32_t o = w[0] ^ 1u;
            for (int i = 1; i < 8; i++) o |= w[i];
            return o == 0;
        };
        auto shr1 = [](uint32_t* w) {
            for (int i = 0; i < 7; i++) w[i] = (w[i] >> 1) | (w[i + 1] << 31);
            w[7] >>= 1;
        };
        auto halve_mod = [&](uint32_t* x) {   // x / 2 mod p for x < p (x + p < 2^255: no carry out)
            if (x[0] & 1u) add8(x, x, m);
            shr1(x);
        };
        while (!is_one(u) && !is_one(v)) {
            while (!(u[0] & 1u)) { shr1(u); halve_mod(x1); }
            while (!(v[0] & 1u)) { shr1(v); halve_mod(x2); }
            if (!sub8(t, u, v)) {   // u >= v
                for (int i = 0; i < 8; i++) u[i] = t[i];
                if (sub8(x1, x1, x2)) add8(x1, x1, m);
            } else {
                sub8(v, v, u);
                if (sub8(x2, x2, x1)) add8(x2, x2, m);
            }
        }
        fe_t r;
        const bool first = is_one(u);
        for (int i = 0; i < 8; i++) r.l[i] = first ? x1[i] : x2[i];
        const fe_t r3 = mul(r2(), r2());
        return mul(r, r3);
    }
};

using Fr = Field<FrP>;
using Fq = Field<FqP>;

}  // namespace b2r
