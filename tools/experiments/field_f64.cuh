// EXPERIMENT (not used by the product): Montgomery product with the a*b half on the FP64 pipe.
//
// Idea (profiles/r01_pipebench.md): on B200 every IMAD.WIDE costs 4 cycles of the fmaheavy pipe, so the 128 of them in
// field.cuh's product bound every kernel of this library, while the FP64 pipe (DFMA at 64 lanes/clk/SM, a separate
// pipe) sits idle.  Here the 512-bit product T = a*b is formed from 5 x 52-bit limbs held as doubles: for every limb
// pair two fused multiply-adds in round-toward-zero mode return the exact high and low 52 bits of the 104-bit product
// (the classic double-precision big-number trick), the partial products are summed as 64-bit integers straight from
// the doubles' bit patterns, and only the Montgomery reduction (64 IMAD.WIDE) stays on the integer multiplier.
// Results are bit-identical to Field<P>::mul (checked on the device: 0 mismatches in 303,104 chains of 64 products).
//
// Measured on B200: 60.1 G products/s against 64.5 for field.cuh - no gain yet.  ncu: 373 instructions per product
// (50 DFMA.RZ + 35 DADD, 56 IMAD.WIDE, 82 IMAD-class moves/adds that ptxas also places on the fma pipe, 170 alu),
// IPC 0.6 per scheduler with no pipe saturated (alu 49 %, fma 32 %, fp64 26 % = the pipe shared by two schedulers),
// stalls split between math-pipe throttle, dispatch stall and not-selected.  To pay off it needs ~340 instructions at
// IPC > 0.9: the limb split / digit layout (62 shifts and masks) and the register moves of the reduction are the
// places to cut.  Kept host-tested (field_f64_host_test.cpp emulates the two DFMA steps with 128-bit integers) as
// the starting point.
#pragma once
#include "../../halo2-rsa_b200/csrc/field.cuh"

namespace b2r {

// bit patterns of hi = RZ(a*b + 2^104) and lo = RZ(a*b + (2^104 + 2^52 - hi)) for integers a, b < 2^52 held in doubles:
//   hi = 0x467 << 52 | floor(a*b / 2^52),   lo = 0x433 << 52 | (a*b mod 2^52)
B2R_HD void dfma_split(double a, double b, uint64_t& hi_bits, uint64_t& lo_bits) {
#if defined(__CUDA_ARCH__)
    const double c104 = __longlong_as_double(0x4670000000000000ll);      // 2^104
    const double c104p52 = __longlong_as_double(0x4670000000000001ll);   // 2^104 + 2^52
    const double hi = __fma_rz(a, b, c104);
    const double sub = c104p52 - hi;  // (1 - H) 2^52, exact
    const double lo = __fma_rz(a, b, sub);
    hi_bits = (uint64_t)__double_as_longlong(hi);
    lo_bits = (uint64_t)__double_as_longlong(lo);
#else
    const unsigned __int128 p = (unsigned __int128)(uint64_t)a * (uint64_t)b;
    hi_bits = ((uint64_t)0x467 << 52) | (uint64_t)(p >> 52);
    lo_bits = ((uint64_t)0x433 << 52) | ((uint64_t)p & (((uint64_t)1 << 52) - 1));
#endif
}

// integer < 2^52 given as (hi 20 bits, lo 32 bits) -> double
B2R_HD double limb_to_double(uint32_t hi20, uint32_t lo32) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double((int)(hi20 | 0x43300000u), (int)lo32) - 4503599627370496.0;
#else
    return (double)(((uint64_t)hi20 << 32) | lo32);
#endif
}

template <class P>
struct FieldF64 {
    // 8 x 32-bit words (value < 2^256) -> 5 limbs of 52 bits as doubles
    B2R_HD static void split52(const fe_t& a, double* d) {
        const uint32_t* w = a.l;
        d[0] = limb_to_double(w[1] & 0xfffffu, w[0]);
        d[1] = limb_to_double(((w[2] >> 20) | (w[3] << 12)) & 0xfffffu, (w[1] >> 20) | (w[2] << 12));
        d[2] = limb_to_double((w[4] >> 8) & 0xfffffu, (w[3] >> 8) | (w[4] << 24));
        d[3] = limb_to_double(((w[5] >> 28) | (w[6] << 4)) & 0xfffffu, (w[4] >> 28) | (w[5] << 4));
        d[4] = limb_to_double((w[7] >> 16) & 0xfffffu, (w[6] >> 16) | (w[7] << 16));
    }

    // T[0..15] = a * b from the 5 x 5 limb products
    B2R_HD static void product(const double* A, const double* B, uint32_t* T) {
        // column k (weight 2^(52 k)) collects the low halves of the pairs with i + j = k and the high halves of the
        // pairs with i + j = k - 1; the exponent fields of the bit patterns are cancelled by the start value
        uint64_t col[10];
#pragma unroll
        for (int k = 0; k < 10; k++) {
            const int nlo = (k <= 4) ? k + 1 : (k <= 8 ? 9 - k : 0);
            const int nhi = (k == 0) ? 0 : ((k - 1 <= 4) ? k : (k - 1 <= 8 ? 10 - k : 0));
            col[k] = (uint64_t)0 - (((uint64_t)nlo * 0x433 + (uint64_t)nhi * 0x467) << 52);
        }
#pragma unroll
        for (int i = 0; i < 5; i++) {
#pragma unroll
            for (int j = 0; j < 5; j++) {
                uint64_t hb, lb;
                dfma_split(A[i], B[j], hb, lb);
                col[i + j] += lb;
                col[i + j + 1] += hb;
            }
        }
        // carry-normalise to 52-bit digits and lay them out as 32-bit words (disjoint bit ranges: OR)
#pragma unroll
        for (int k = 0; k < 16; k++) T[k] = 0;
        uint64_t carry = 0;
#pragma unroll
        for (int k = 0; k < 10; k++) {
            const uint64_t v = col[k] + carry;
            const uint64_t dgt = v & (((uint64_t)1 << 52) - 1);
            carry = v >> 52;
            const int bit = 52 * k, wd = bit >> 5, sh = bit & 31;
            if (wd < 16) T[wd] |= (uint32_t)(dgt << sh);
            if (wd + 1 < 16) T[wd + 1] |= (uint32_t)(sh ? (dgt >> (32 - sh)) : (dgt >> 32));
            if (sh > 12 && wd + 2 < 16) T[wd + 2] |= (uint32_t)(dgt >> (64 - sh));
        }
    }

    // Montgomery reduction of T (16 words, T < p * 2^256) to T / 2^256 mod p, fully reduced.  Word-serial REDC in the
    // two-accumulator form of Field<P>::mul: V = Pw + 2^32 * Sw holds the sliding 9-word window, each round adds m_i * p
    // (even limbs of p into Pw, odd limbs fused into the shift) and injects the next high word of T.
    B2R_HD static fe_t redc(const uint32_t* T) {
        uint32_t m[8];
#pragma unroll
        for (int i = 0; i < 8; i++) m[i] = P::MOD(i);
        uint32_t Pw[10], Sw[10];
#pragma unroll
        for (int k = 0; k < 8; k++) Pw[k] = T[k];
        uint32_t mi = Pw[0] * P::N0INV;
        Pw[8] = row_mad(Pw, &m[0], mi);
        row_mul(Sw, &m[1], mi);
#pragma unroll
        for (int i = 1; i < 8; i++) {
            mi = (Sw[0] + Pw[1]) * P::N0INV;
            row_shift_mad(Pw, &Sw[0], &m[1], mi);
            uint32_t nP[10], nS[10];
#pragma unroll
            for (int k = 0; k < 8; k++) nS[k] = Pw[k + 2];
#pragma unroll
            for (int k = 0; k < 8; k++) nP[k] = Sw[k];
            nP[8] = row_mad(nP, &m[0], mi);
            // next high word of T at relative word 7
            const uint64_t t7 = (uint64_t)nP[7] + T[7 + i];
            nP[7] = (uint32_t)t7;
            nP[8] += (uint32_t)(t7 >> 32);
#pragma unroll
            for (int k = 0; k < 9; k++) Pw[k] = nP[k];
#pragma unroll
            for (int k = 0; k < 8; k++) Sw[k] = nS[k];
        }
        fe_t r;
        add8(r.l, Sw, &Pw[1]);
        r.l[7] += T[15];  // < 2p < 2^255: no carry out
        Field<P>::final_sub(r.l);
        return r;
    }

    B2R_HD static fe_t mul(const fe_t& a, const fe_t& b) {
        double A[5], B[5];
        split52(a, A);
        split52(b, B);
        uint32_t T[16];
        product(A, B, T);
        return redc(T);
    }
    B2R_HD static fe_t sqr(const fe_t& a) {
        double A[5];
        split52(a, A);
        uint32_t T[16];
        product(A, A, T);
        return redc(T);
    }
};

using FrF64 = FieldF64<FrP>;
using FqF64 = FieldF64<FqP>;

}  // namespace b2r
