// Host-side Fiat-Shamir transcript of the prover: halo2_proofs' Blake2bWrite<_, G1Affine, Challenge255<_>>
// (third-party crate; reference call site benches/bench.rs:320).  Blake2b-512 (RFC 7693) personalised
// "Halo2-Transcript"; prefix byte 0 before squeezing a challenge (64-byte digest of a CLONE of the
// state, reduced mod r), 1 before a point (x, y canonical little-endian), 2 before a scalar.  The
// proof stream receives compressed points (x little-endian, sign of y in bit 255; identity = zeros)
// and canonical scalars.  The hashing is a few KB per proof, so it stays on the host between the
// GPU phases; oracle/plonk.py restates the same transcript independently (hashlib).
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

#include "field.cuh"

namespace b2r {

struct Blake2b {
    uint64_t h[8];
    uint64_t t[2];
    uint8_t buf[128];
    size_t buflen;

    static uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
    static uint64_t load64(const uint8_t* p) {
        uint64_t v = 0;
        for (int i = 7; i >= 0; i--) v = (v << 8) | p[i];
        return v;
    }
    static const uint64_t* iv() {
        static const uint64_t IV[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                                       0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
        return IV;
    }
    void init(const char personal[16]) {
        uint8_t param[64];
        memset(param, 0, sizeof param);
        param[0] = 64;  // digest length
        param[2] = 1;   // fanout
        param[3] = 1;   // depth
        memcpy(param + 48, personal, 16);
        for (int i = 0; i < 8; i++) h[i] = iv()[i] ^ load64(param + 8 * i);
        t[0] = t[1] = 0;
        buflen = 0;
    }
    void compress(const uint8_t* block, bool last) {
        static const uint8_t S[12][16] = {
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
            {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
            {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
            {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
            {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
        uint64_t m[16], v[16];
        for (int i = 0; i < 16; i++) m[i] = load64(block + 8 * i);
        for (int i = 0; i < 8; i++) v[i] = h[i], v[i + 8] = iv()[i];
        v[12] ^= t[0];
        v[13] ^= t[1];
        if (last) v[14] = ~v[14];
#define B2R_G(r, i, a, b, c, d)                    \
    do {                                           \
        a = a + b + m[S[r][2 * i]];                \
        d = rotr(d ^ a, 32);                       \
        c = c + d;                                 \
        b = rotr(b ^ c, 24);                       \
        a = a + b + m[S[r][2 * i + 1]];            \
        d = rotr(d ^ a, 16);                       \
        c = c + d;                                 \
        b = rotr(b ^ c, 63);                       \
    } while (0)
        for (int r = 0; r < 12; r++) {
            B2R_G(r, 0, v[0], v[4], v[8], v[12]);
            B2R_G(r, 1, v[1], v[5], v[9], v[13]);
            B2R_G(r, 2, v[2], v[6], v[10], v[14]);
            B2R_G(r, 3, v[3], v[7], v[11], v[15]);
            B2R_G(r, 4, v[0], v[5], v[10], v[15]);
            B2R_G(r, 5, v[1], v[6], v[11], v[12]);
            B2R_G(r, 6, v[2], v[7], v[8], v[13]);
            B2R_G(r, 7, v[3], v[4], v[9], v[14]);
        }
#undef B2R_G
        for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
    }
    void update(const uint8_t* in, size_t len) {
        while (len > 0) {
            if (buflen == 128) {  // only compress a full buffer when more input follows (the last block is special)
                t[0] += 128;
                if (t[0] < 128) t[1]++;
                compress(buf, false);
                buflen = 0;
            }
            size_t take = 128 - buflen;
            if (take > len) take = len;
            memcpy(buf + buflen, in, take);
            buflen += take;
            in += take;
            len -= take;
        }
    }
    // digest of the data so far; the state itself is left untouched (callers squeeze from a clone)
    void digest(uint8_t out[64]) const {
        Blake2b c = *this;
        c.t[0] += c.buflen;
        if (c.t[0] < c.buflen) c.t[1]++;
        memset(c.buf + c.buflen, 0, 128 - c.buflen);
        c.compress(c.buf, true);
        for (int i = 0; i < 8; i++)
            for (int j = 0; j < 8; j++) out[8 * i + j] = (uint8_t)(c.h[i] >> (8 * j));
    }
};

// ---- host field helpers (the host build of field.cuh; a few operations per proof) ------------------
inline void fe_to_bytes(const fe_t& canon, uint8_t out[32]) {
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 4; j++) out[4 * i + j] = (uint8_t)(canon.l[i] >> (8 * j));
}
inline fe_t fe_from_bytes(const uint8_t in[32]) {
    fe_t r;
    for (int i = 0; i < 8; i++) r.l[i] = (uint32_t)in[4 * i] | ((uint32_t)in[4 * i + 1] << 8) | ((uint32_t)in[4 * i + 2] << 16) | ((uint32_t)in[4 * i + 3] << 24);
    return r;
}
// 512-bit little-endian integer mod r, returned in Montgomery form (Challenge255 / from_bytes_wide)
inline fe_t fr_from_wide(const uint8_t in[64]) {
    fe_t lo = fe_from_bytes(in), hi = fe_from_bytes(in + 32);
    // bring both halves below r first (2^256 < 6 r): the Montgomery product expects reduced inputs
    uint32_t m[8], t[8];
    for (int i = 0; i < 8; i++) m[i] = FrP::MOD(i);
    for (int it = 0; it < 6; it++) {
        if (!sub8(t, lo.l, m)) for (int i = 0; i < 8; i++) lo.l[i] = t[i];
        if (!sub8(t, hi.l, m)) for (int i = 0; i < 8; i++) hi.l[i] = t[i];
    }
    // value = lo + hi * 2^256.  mul(x, R2) = x*R = Montgomery form of x; mul(mul(hi, R2), R2) = hi * R^2 =
    // Montgomery form of hi * 2^256.
    fe_t r2 = Fr::r2();
    fe_t lo_m = Fr::mul(lo, r2);
    fe_t hi_m = Fr::mul(Fr::mul(hi, r2), r2);
    return Fr::add(lo_m, hi_m);
}

struct Transcript {
    Blake2b st;
    std::vector<uint8_t> out;
    Transcript() { st.init("Halo2-Transcript"); }
    void common_scalar(const fe_t& mont) {
        uint8_t b[33];
        b[0] = 2;
        fe_to_bytes(Fr::from_mont(mont), b + 1);
        st.update(b, 33);
    }
    void write_scalar(const fe_t& mont) {
        common_scalar(mont);
        uint8_t b[32];
        fe_to_bytes(Fr::from_mont(mont), b);
        out.insert(out.end(), b, b + 32);
    }
    // point given as Montgomery affine coordinates (identity = (0, 0))
    void write_point(const fe_t& x_mont, const fe_t& y_mont) {
        uint8_t b[65];
        b[0] = 1;
        fe_t x = Fq::from_mont(x_mont), y = Fq::from_mont(y_mont);
        fe_to_bytes(x, b + 1);
        fe_to_bytes(y, b + 33);
        st.update(b, 65);
        uint8_t c[32];
        fe_to_bytes(x, c);
        c[31] |= (uint8_t)((y.l[0] & 1u) << 7);
        out.insert(out.end(), c, c + 32);
    }
    fe_t squeeze() {
        uint8_t p = 0;
        st.update(&p, 1);
        uint8_t d[64];
        st.digest(d);
        return fr_from_wide(d);
    }
};

}  // namespace b2r
