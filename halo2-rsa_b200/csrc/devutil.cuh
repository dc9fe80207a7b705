// Small device helpers shared by the translation units: 128-bit field-element loads/stores and
// the keyed blinding stream that stands in for halo2's OsRng on both the product and the oracle
// (oracle/plonk.py: blind_fe restates it).
#pragma once
#include "field.cuh"

namespace b2r {

#if defined(__CUDACC__)
__device__ __forceinline__ fe_t ldv(const fe_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    fe_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ fe_t ldv_nc(const fe_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    fe_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void stv(fe_t* p, const fe_t& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

#endif  // __CUDACC__ (the rest also compiles for the host: tests/host/blind_host_test.cpp)

// blinding streams: one per polynomial the prover blinds (ids shared with oracle/plonk.py)
enum BlindStream : uint32_t {
    ST_ADVICE = 0,       // + advice column
    ST_LOOKUP_A = 8,     // + lookup index
    ST_LOOKUP_S = 16,
    ST_LOOKUP_Z = 24,
    ST_PERM_Z = 32,      // + permutation set
    ST_RANDOM_POLY = 40
};

// ---- blinding stream v2: ChaCha20 (RFC 7539 block function, 20 rounds) keyed by a 256-bit seed ------------------
// Every blinded cell (the rows halo2's create_proof fills from its RNG, the lookup / grand-product blinds and the
// random polynomial of the vanishing argument) is one 64-byte ChaCha20 block reduced mod r as a 512-bit
// little-endian integer - what halo2curves' Fr::random / from_bytes_wide does with 64 RNG bytes, so the field element
// is uniform up to 2^-258.  Block position: word 12 = row | stream << 24, word 13 = proof index inside the call,
// words 14-15 = the 64-bit call nonce, so one (key, nonce) pair never produces the same block for two cells.
// oracle/plonk.py (blind_fe) and oracle/plonk_prover.c restate the stream independently.
struct BlindKey {
    uint32_t key[8];
    uint32_t nonce[2];
    uint32_t on;  // 0 = leave blinding rows zero (witness-only entry points)
};

B2R_HD uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
#define B2R_CHACHA_QR(a, b, c, d)      \
    do {                               \
        a += b; d ^= a; d = rotl32(d, 16); \
        c += d; b ^= c; b = rotl32(b, 12); \
        a += b; d ^= a; d = rotl32(d, 8);  \
        c += d; b ^= c; b = rotl32(b, 7);  \
    } while (0)
B2R_HD void chacha20_block(const BlindKey& K, uint32_t w12, uint32_t w13, uint32_t out[16]) {
    uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, K.key[0], K.key[1], K.key[2], K.key[3],
                      K.key[4], K.key[5], K.key[6], K.key[7], w12, w13, K.nonce[0], K.nonce[1]};
    uint32_t x[16];
    for (int i = 0; i < 16; i++) x[i] = s[i];
    for (int r = 0; r < 10; r++) {
        B2R_CHACHA_QR(x[0], x[4], x[8], x[12]);
        B2R_CHACHA_QR(x[1], x[5], x[9], x[13]);
        B2R_CHACHA_QR(x[2], x[6], x[10], x[14]);
        B2R_CHACHA_QR(x[3], x[7], x[11], x[15]);
        B2R_CHACHA_QR(x[0], x[5], x[10], x[15]);
        B2R_CHACHA_QR(x[1], x[6], x[11], x[12]);
        B2R_CHACHA_QR(x[2], x[7], x[8], x[13]);
        B2R_CHACHA_QR(x[3], x[4], x[9], x[14]);
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
}
#undef B2R_CHACHA_QR

// (lo + hi * 2^256) mod r in Montgomery form; lo, hi arbitrary 256-bit values
B2R_HD fe_t fr_from_u512(fe_t lo, fe_t hi) {
    uint32_t m[8], t[8];
    for (int i = 0; i < 8; i++) m[i] = FrP::MOD(i);
    for (int it = 0; it < 5; it++) {  // 2^256 < 6 r
        uint32_t bl = sub8(t, lo.l, m);
        for (int i = 0; i < 8; i++) lo.l[i] = bl ? lo.l[i] : t[i];
        uint32_t bh = sub8(t, hi.l, m);
        for (int i = 0; i < 8; i++) hi.l[i] = bh ? hi.l[i] : t[i];
    }
    const fe_t r2 = Fr::r2();
    return Fr::add(Fr::mul(lo, r2), Fr::mul(Fr::mul(hi, r2), r2));
}

B2R_HD fe_t blind_value(const BlindKey& K, uint32_t proof, uint32_t stream, uint32_t row) {
    uint32_t o[16];
    chacha20_block(K, row | (stream << 24), proof, o);
    fe_t lo, hi;
    for (int i = 0; i < 8; i++) lo.l[i] = o[i], hi.l[i] = o[8 + i];
    return fr_from_u512(lo, hi);
}

// 64-bit seeds of the older entry points: key = le64(seed) || "b2rsa-blind-seed64-v2" padded with zeros to 24 bytes.
// 64 bits of entropy: for tests and benchmarks; production callers pass 32 random bytes (b2r_rsa_prove_batch_ex).
inline BlindKey blind_key_from_seed64(uint64_t seed, uint64_t nonce) {
    BlindKey K;
    uint8_t kb[32] = {0};
    for (int i = 0; i < 8; i++) kb[i] = (uint8_t)(seed >> (8 * i));
    const char* pad = "b2rsa-blind-seed64-v2";
    for (int i = 0; pad[i]; i++) kb[8 + i] = (uint8_t)pad[i];
    for (int i = 0; i < 8; i++) K.key[i] = (uint32_t)kb[4 * i] | ((uint32_t)kb[4 * i + 1] << 8) | ((uint32_t)kb[4 * i + 2] << 16) | ((uint32_t)kb[4 * i + 3] << 24);
    K.nonce[0] = (uint32_t)nonce;
    K.nonce[1] = (uint32_t)(nonce >> 32);
    K.on = seed != 0;
    return K;
}
inline BlindKey blind_key_from_bytes(const uint8_t kb[32], uint64_t nonce) {
    BlindKey K;
    for (int i = 0; i < 8; i++) K.key[i] = (uint32_t)kb[4 * i] | ((uint32_t)kb[4 * i + 1] << 8) | ((uint32_t)kb[4 * i + 2] << 16) | ((uint32_t)kb[4 * i + 3] << 24);
    K.nonce[0] = (uint32_t)nonce;
    K.nonce[1] = (uint32_t)(nonce >> 32);
    K.on = 1;
    return K;
}

}  // namespace b2r
