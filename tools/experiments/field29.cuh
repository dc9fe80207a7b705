// EXPERIMENT (not used by the product): BN254 Fr / Fq arithmetic in a reduced radix, 9 limbs of 29 bits,
// Montgomery form with R' = 2^261, lazily reduced.
//
// Hypothesis: field.cuh's carry-chained IMAD.WIDE.U32.X rows run at half the plain IMAD.WIDE rate, so a
// carry-free product-scanning multiplier (64-bit column accumulators, 162 plain IMAD.WIDE + shifts) would win.
// Measured on B200 (tools/pipebench.cu, profiles/r01_pipebench.md): it does not.  ncu shows IMAD.WIDE.U32 of ANY
// form occupies the fmaheavy pipe for 4 cycles per warp instruction (32 lanes/clk/SM; 32-bit IMAD is 64), so the
// cost of a Montgomery product is ~4 cycles x (number of 32x32->64 products): 128 for field.cuh (64 Gmul/s/GPU),
// 171 here (53 Gmul/s).  The 32-bit-limb CIOS multiplier is already at the integer-multiplier roofline; the only
// larger multiplier on the SM is the FP64 pipe (DFMA 64/clk/SM, separate from fmaheavy).  Kept as a record and
// as a host-tested starting point for lazy-reduction bookkeeping.
//
// Representation invariants of F29<P>::el
//   * limbs v[0..8] < 2^29 ("normalised"), value V = sum v[i] 2^(29 i) < 2^261;
//   * V is only congruent to the element: V < B * p with a bound B the caller tracks.  p < 2^254, so
//     R' / p > 128 and  mul(a, b) < p (1 + B(a) B(b) / 128):  B(a) B(b) <= 128  =>  result < 2p.
//   * add / sub never reduce: B(a + b) = B(a) + B(b);  sub<K>(a, b) = a + K p - b needs B(b) <= K.
//   * the additive identity used as a flag (point at infinity: ZZ = 0) is all limbs zero, and
//     mul(0, x) is exactly 0.
// The memory format stays halo2curves' 4 x u64 Montgomery (R = 2^256) at the ABI; pack / unpack move
// between the 8 x 32 container and the limbs, and the callers fold 2^5 = R' / R into constants.
//
// Host + device source (plain 64-bit C arithmetic): tests/host/field29_host_test.cpp checks it against
// field.cuh in this GPU-less container.
#pragma once
#include "../../halo2-rsa_b200/csrc/field.cuh"

namespace b2r {

// acc + a * b: one IMAD.WIDE.U32 accumulating in place (ptxas fuses mul.wide.u32 + add.s64 into a chain).
B2R_HD uint64_t mad_wide(uint32_t a, uint32_t b, uint64_t acc) { return acc + (uint64_t)a * b; }
// Hides the value range of a Montgomery quotient digit from NVVM.  Without this it keeps the digit as a 64-bit
// value ((acc * N0INV) & MASK in 64 bits) and every digit * modulus product becomes a 64-bit multiply: one extra
// add per product, 81 per Montgomery product (seen in the SASS as VIADD Rhi, Rhi, UR<zero>).
B2R_HD uint32_t opaque32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    asm("mov.b32 %0, %0;" : "+r"(x));
#endif
    return x;
}

template <class P>
struct F29 {
    static constexpr int NL = 9;
    static constexpr uint32_t MASK = (1u << 29) - 1;
    struct el {
        uint32_t v[NL];
    };

    struct Tab {
        uint32_t v[NL];
    };
    // limb i of p: bits [29 i, 29 i + 29)
    B2R_HD static constexpr uint32_t MOD(int i) {
        const int bit = 29 * i, q = bit >> 5, s = bit & 31;
        uint64_t two = (uint64_t)P::MOD(q) | ((uint64_t)(q + 1 < 8 ? P::MOD(q + 1) : 0u) << 32);
        return (uint32_t)(two >> s) & MASK;
    }
    static constexpr uint32_t N0INV = P::N0INV & MASK;  // -p^-1 mod 2^29
    // limbs of K * p (K <= 128 keeps K p < 2^261; the top limb is not masked)
    B2R_HD static constexpr Tab kmod_tab(uint32_t K) {
        Tab t = {};
        uint64_t carry = 0;
        for (int j = 0; j < NL; j++) {
            uint64_t x = (uint64_t)MOD(j) * K + carry;
            t.v[j] = (uint32_t)((j == NL - 1) ? x : (x & MASK));
            carry = x >> 29;
        }
        return t;
    }

    B2R_HD static el zero() {
        el r;
        for (int i = 0; i < NL; i++) r.v[i] = 0;
        return r;
    }
    B2R_HD static bool is_zero_limbs(const el& a) {
        uint32_t o = 0;
        for (int i = 0; i < NL; i++) o |= a.v[i];
        return o == 0;
    }
    B2R_HD static el select(bool c, const el& a, const el& b) {
        el r;
        for (int i = 0; i < NL; i++) r.v[i] = c ? a.v[i] : b.v[i];
        return r;
    }

    // 8 x 32 container (any value < 2^256) -> limbs
    B2R_HD static el unpack(const fe_t& a) {
        el r;
#pragma unroll
        for (int i = 0; i < NL; i++) {
            const int bit = 29 * i, q = bit >> 5, s = bit & 31;
            uint32_t lo = a.l[q] >> s;
            if (s > 3 && q + 1 < 8) lo |= a.l[q + 1] << (32 - s);
            r.v[i] = lo & MASK;
        }
        return r;
    }
    // limbs (value < 2^256) -> 8 x 32 container
    B2R_HD static fe_t pack(const el& a) {
        fe_t r;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            // word q covers bits [32 q, 32 q + 32): limbs i0 = floor(32 q / 29) and i0 + 1
            const int i0 = (32 * q) / 29, s = 32 * q - 29 * i0;
            uint32_t w = a.v[i0] >> s;
            if (i0 + 1 < NL) w |= a.v[i0 + 1] << (29 - s);
            if (29 - s + 29 < 32 && i0 + 2 < NL) w |= a.v[i0 + 2] << (58 - s);
            r.l[q] = w;
        }
        return r;
    }

    // Montgomery product a b 2^-261 (mod p), finely integrated product scanning.
    B2R_HD static el mul(const el& a, const el& b) {
        uint32_t m[NL];
        el r;
        uint64_t acc = 0;
#pragma unroll
        for (int k = 0; k < NL; k++) {
#pragma unroll
            for (int i = 0; i <= k; i++) acc = mad_wide(a.v[i], b.v[k - i], acc);
#pragma unroll
            for (int i = 0; i < k; i++) acc = mad_wide(m[i], MOD(k - i), acc);
            m[k] = opaque32(((uint32_t)acc * N0INV) & MASK);
            acc = mad_wide(m[k], MOD(0), acc);
            acc >>= 29;
        }
#pragma unroll
        for (int k = NL; k < 2 * NL - 1; k++) {
#pragma unroll
            for (int i = k - NL + 1; i < NL; i++) acc = mad_wide(a.v[i], b.v[k - i], acc);
#pragma unroll
            for (int i = k - NL + 1; i < NL; i++) acc = mad_wide(m[i], MOD(k - i), acc);
            r.v[k - NL] = (uint32_t)acc & MASK;
            acc >>= 29;
        }
        r.v[NL - 1] = (uint32_t)acc;
        return r;
    }
    // a^2 2^-261: the 36 cross products are taken once against the doubled operand
    B2R_HD static el sqr(const el& a) {
        uint32_t m[NL], d[NL];
        el r;
#pragma unroll
        for (int i = 0; i < NL; i++) d[i] = a.v[i] << 1;
        uint64_t acc = 0;
#pragma unroll
        for (int k = 0; k < NL; k++) {
#pragma unroll
            for (int i = 0; 2 * i < k; i++) acc = mad_wide(d[i], a.v[k - i], acc);
            if ((k & 1) == 0) acc = mad_wide(a.v[k / 2], a.v[k / 2], acc);
#pragma unroll
            for (int i = 0; i < k; i++) acc = mad_wide(m[i], MOD(k - i), acc);
            m[k] = opaque32(((uint32_t)acc * N0INV) & MASK);
            acc = mad_wide(m[k], MOD(0), acc);
            acc >>= 29;
        }
#pragma unroll
        for (int k = NL; k < 2 * NL - 1; k++) {
#pragma unroll
            for (int i = k - NL + 1; 2 * i < k; i++) acc = mad_wide(d[i], a.v[k - i], acc);
            if ((k & 1) == 0) acc = mad_wide(a.v[k / 2], a.v[k / 2], acc);
#pragma unroll
            for (int i = k - NL + 1; i < NL; i++) acc = mad_wide(m[i], MOD(k - i), acc);
            r.v[k - NL] = (uint32_t)acc & MASK;
            acc >>= 29;
        }
        r.v[NL - 1] = (uint32_t)acc;
        return r;
    }

    // limb-wise carry propagation of signed 32-bit limb sums; the value must be in [0, 2^261)
    B2R_HD static el normalise(const int32_t* t) {
        el r;
        int32_t carry = 0;
#pragma unroll
        for (int i = 0; i < NL - 1; i++) {
            int32_t x = t[i] + carry;
            r.v[i] = (uint32_t)x & MASK;
            carry = x >> 29;
        }
        r.v[NL - 1] = (uint32_t)(t[NL - 1] + carry);
        return r;
    }
    B2R_HD static el add(const el& a, const el& b) {
        int32_t t[NL];
#pragma unroll
        for (int i = 0; i < NL; i++) t[i] = (int32_t)(a.v[i] + b.v[i]);
        return normalise(t);
    }
    B2R_HD static el dbl(const el& a) { return add(a, a); }
    // a + K p - b;  requires value(b) <= K p
    template <uint32_t K>
    B2R_HD static el sub(const el& a, const el& b) {
        constexpr Tab kp = kmod_tab(K);
        int32_t t[NL];
#pragma unroll
        for (int i = 0; i < NL; i++) t[i] = (int32_t)(a.v[i] + kp.v[i]) - (int32_t)b.v[i];
        return normalise(t);
    }
    // K p - a
    template <uint32_t K>
    B2R_HD static el neg(const el& a) {
        constexpr Tab kp = kmod_tab(K);
        int32_t t[NL];
#pragma unroll
        for (int i = 0; i < NL; i++) t[i] = (int32_t)kp.v[i] - (int32_t)a.v[i];
        return normalise(t);
    }

    // a - p if that is non-negative, else a
    B2R_HD static el cond_sub_p(const el& a) {
        int32_t t[NL];
#pragma unroll
        for (int i = 0; i < NL; i++) t[i] = (int32_t)a.v[i] - (int32_t)MOD(i);
        el r;
        int32_t carry = 0;
#pragma unroll
        for (int i = 0; i < NL - 1; i++) {
            int32_t x = t[i] + carry;
            r.v[i] = (uint32_t)x & MASK;
            carry = x >> 29;
        }
        const int32_t top = t[NL - 1] + carry;
        r.v[NL - 1] = (uint32_t)top;
        return select(top < 0, a, r);
    }
    // canonical representative in [0, p) of any normalised value (< 2^261)
    B2R_HD static el reduce(const el& a) {
        // quotient estimate from the top limb: p / 2^232 has 22 bits; q <= floor(V / p) <= q + 1
        constexpr uint32_t PT = MOD(NL - 1) + 1;  // > p / 2^232
        constexpr uint32_t RECIP = (uint32_t)(((uint64_t)1 << 32) / PT);
        const uint32_t q = (uint32_t)(((uint64_t)a.v[NL - 1] * RECIP) >> 32);
        int64_t t;
        int32_t s[NL];
        int64_t carry = 0;
#pragma unroll
        for (int i = 0; i < NL - 1; i++) {
            t = (int64_t)a.v[i] - (int64_t)((uint64_t)q * MOD(i)) + carry;
            s[i] = (int32_t)((uint32_t)t & MASK);
            carry = t >> 29;
        }
        s[NL - 1] = (int32_t)((int64_t)a.v[NL - 1] - (int64_t)((uint64_t)q * MOD(NL - 1)) + carry);
        el r;
#pragma unroll
        for (int i = 0; i < NL; i++) r.v[i] = (uint32_t)s[i];
        r = cond_sub_p(r);
        r = cond_sub_p(r);
        r = cond_sub_p(r);
        return r;
    }
    // value == 0 (mod p)?  cheap filter on the low limb first, exact reduction only when it passes
    B2R_HD static bool is_zero_mod_p(const el& a) {
        constexpr uint32_t PT = MOD(NL - 1) + 1;
        constexpr uint32_t RECIP = (uint32_t)(((uint64_t)1 << 32) / PT);
        const uint32_t q = (uint32_t)(((uint64_t)a.v[NL - 1] * RECIP) >> 32);
        bool maybe = false;
#pragma unroll
        for (uint32_t d = 0; d < 4; d++) maybe = maybe || (((a.v[0] - (q + d) * MOD(0)) & MASK) == 0);
        if (!maybe) return false;
        return is_zero_limbs(reduce(a));
    }

    // ---- conversions against field.cuh's Montgomery-2^256 container ---------------------------------
    // from a canonical Montgomery-2^256 element to the lazy Montgomery-2^261 domain (result < 2p)
    B2R_HD static el from_mont256(const fe_t& a) {
        // 2^266 mod p = (2^10 as a field element) in Montgomery-2^256 form, computed with field.cuh on the fly is
        // wasteful in kernels: callers that convert in bulk should pre-scale instead.  This helper is for set-up code.
        fe_t c = Field<P>::zero();
        c.l[0] = 1u << 10;
        c = Field<P>::to_mont(c);  // 2^10 * 2^256 = 2^266 (mod p)
        return mul(unpack(a), unpack(c));
    }
    // back to a canonical Montgomery-2^256 element
    B2R_HD static fe_t to_mont256(const el& a) {
        fe_t c = Field<P>::one();  // 2^256 mod p
        return pack(reduce(mul(a, unpack(c))));
    }
};

using Fr29 = F29<FrP>;
using Fq29 = F29<FqP>;

}  // namespace b2r
