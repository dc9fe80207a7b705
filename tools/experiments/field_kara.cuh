// EXPERIMENT (not used by the product): Montgomery product whose a*b half is ONE level of Karatsuba.
//
// Idea ("lower the floor instead of approaching it"): every kernel of the library is bound by the 128
// IMAD.WIDE of field.cuh's product (4 fmaheavy cycles each).  Splitting a = a1 2^128 + a0, b = b1 2^128 + b0,
//     a b = z2 2^256 + (zm - z0 - z2) 2^128 + z0,   z0 = a0 b0, z2 = a1 b1, zm = (a0 + a1)(b0 + b1),
// needs 3 x 16 = 48 wide products instead of 64 for the 512-bit product; the Montgomery reduction (64 wide products) is
// field.cuh's own.  112 instead of 128 wide MADs (-12.5 %), paid for with ~90 ALU-pipe additions / selects, which run
// on the other pipe.  Bit-identical to Field<P>::mul (host test: kara_host_test.cpp; device: tools/pipebench).
#pragma once
#include "../../halo2-rsa_b200/csrc/field.cuh"

namespace b2r {

// Z[0..7] = x[0..3] * y[0..3] (4 x 4 limbs): rows land on an even- and an odd-aligned lane accumulator, summed at the end
B2R_HD void mul4x4(uint32_t* Z, const uint32_t* x, const uint32_t* y) {
    uint32_t E[10], O[10];
    for (int i = 0; i < 10; i++) E[i] = O[i] = 0;
    // j = 0: (x0, x2) y0 at words (0,1), (2,3) -> E ; (x1, x3) y0 at words (1,2), (3,4) -> O
    lanes_mad_c<2>(&E[0], &x[0], y[0]);
    lanes_mad_c<2>(&O[1], &x[1], y[0]);
    // j = 1: (x0, x2) y1 at (1,2), (3,4) -> O ; (x1, x3) y1 at (2,3), (4,5) -> E
    lanes_mad_c<2>(&O[1], &x[0], y[1]);
    lanes_mad_c<2>(&E[2], &x[1], y[1]);
    // j = 2
    lanes_mad_c<2>(&E[2], &x[0], y[2]);
    lanes_mad_c<2>(&O[3], &x[1], y[2]);
    // j = 3
    lanes_mad_c<2>(&O[3], &x[0], y[3]);
    lanes_mad_c<2>(&E[4], &x[1], y[3]);
    add8(Z, E, O);   // the product fits 8 words: no carry out
}

template <class P>
struct FieldKara {
    using F = Field<P>;
    // Montgomery reduction of a 16-word value T < p * 2^256 (field.cuh: the tail of Field::sqr)
    B2R_HD static fe_t reduce16(const uint32_t* T) {
        uint32_t m[8];
        for (int i = 0; i < 8; i++) m[i] = P::MOD(i);
        uint32_t Pw[10], Sw[10];
        for (int k = 0; k < 8; k++) Pw[k] = T[k];
        uint32_t mi = Pw[0] * P::N0INV;
        Pw[8] = row_mad(Pw, &m[0], mi);
        row_mul(Sw, &m[1], mi);
#pragma unroll
        for (int i = 1; i < 8; i++) {
            mi = (Sw[0] + Pw[1]) * P::N0INV;
            row_shift_mad(Pw, &Sw[0], &m[1], mi);
            uint32_t nP[10], nS[10];
            for (int k = 0; k < 8; k++) nS[k] = Pw[k + 2];
            for (int k = 0; k < 8; k++) nP[k] = Sw[k];
            nP[8] = row_mad(nP, &m[0], mi);
            for (int k = 0; k < 9; k++) Pw[k] = nP[k];
            for (int k = 0; k < 8; k++) Sw[k] = nS[k];
        }
        fe_t r;
        uint32_t u[8];
        add8(u, Sw, &Pw[1]);
        add8(r.l, u, &T[8]);
        F::final_sub(r.l);
        return r;
    }

    B2R_HD static fe_t mul(const fe_t& a, const fe_t& b) {
        uint32_t T[16];
        mul4x4(&T[0], &a.l[0], &b.l[0]);   // z0
        mul4x4(&T[8], &a.l[4], &b.l[4]);   // z2
        // sa = a0 + a1, sb = b0 + b1 (4 words + carry bit)
        uint32_t sa[4], sb[4], ca, cb;
        {
            uint64_t c = 0;
            for (int i = 0; i < 4; i++) { c += (uint64_t)a.l[i] + a.l[4 + i]; sa[i] = (uint32_t)c; c >>= 32; }
            ca = (uint32_t)c;
            c = 0;
            for (int i = 0; i < 4; i++) { c += (uint64_t)b.l[i] + b.l[4 + i]; sb[i] = (uint32_t)c; c >>= 32; }
            cb = (uint32_t)c;
        }
        // zm = (sa + ca 2^128)(sb + cb 2^128), 9 words (+ the bit ca cb at word 8)
        uint32_t zm[10];
        mul4x4(zm, sa, sb);
        zm[8] = ca & cb;
        zm[9] = 0;
        {
            const uint32_t ma = 0u - ca, mb = 0u - cb;
            uint64_t c = 0;
            for (int i = 0; i < 4; i++) { c += (uint64_t)zm[4 + i] + (sb[i] & ma) + (sa[i] & mb); zm[4 + i] = (uint32_t)c; c >>= 32; }
            zm[8] += (uint32_t)c;
        }
        // mid = zm - z0 - z2 (>= 0, fits 9 words)
        {
            int64_t br = 0;
            for (int i = 0; i < 9; i++) {
                int64_t t = (int64_t)zm[i] - (i < 8 ? (int64_t)T[i] : 0) - (i < 8 ? (int64_t)T[8 + i] : 0) + br;
                zm[i] = (uint32_t)t;
                br = t >> 32;   // arithmetic shift: -2 .. 0
            }
        }
        // T += mid * 2^128
        {
            uint64_t c = 0;
            for (int i = 0; i < 12; i++) { c += (uint64_t)T[4 + i] + (i < 9 ? zm[i] : 0u); T[4 + i] = (uint32_t)c; c >>= 32; }
        }
        return reduce16(T);
    }
};

using FqKara = FieldKara<FqP>;
using FrKara = FieldKara<FrP>;

}  // namespace b2r
