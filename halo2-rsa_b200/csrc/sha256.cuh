// SHA-256 (FIPS 180-4) of one message, written once for the device kernel (sha256.cu) and for the host build that unit
// tests it in this GPU-less container (tests/host/sha256_host_test.cpp - test scaffolding like the host build of
// field.cuh, not a CPU fallback: no product path calls it).
#pragma once
#include <stdint.h>

#include "field.cuh"   // B2R_HD

namespace b2r {

B2R_HD uint32_t sha_k(int t) {
    constexpr uint32_t K[64] = {
        0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u, 0xd807aa98u, 0x12835b01u,
        0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u, 0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu,
        0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau, 0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u,
        0x06ca6351u, 0x14292967u, 0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
        0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u, 0x19a4c116u, 0x1e376c08u,
        0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u, 0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u,
        0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};
    return K[t];
}
B2R_HD uint32_t sha_rotr(uint32_t x, int r) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(x, x, r);
#else
    return (x >> r) | (x << (32 - r));
#endif
}

// one 64-byte block; w[16] holds the big-endian message words and is used as the rolling schedule
B2R_HD void sha256_block(uint32_t st[8], uint32_t w[16]) {
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll 1
    for (int t0 = 0; t0 < 64; t0 += 16) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (t0) {
                const uint32_t w15 = w[(j + 1) & 15], w2 = w[(j + 14) & 15];
                const uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
                const uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
                w[j] = w[j] + s0 + w[(j + 9) & 15] + s1;
            }
            const uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
            const uint32_t ch = (e & f) ^ (~e & g);
            const uint32_t t1 = h + S1 + ch + sha_k(t0 + j) + w[j];
            const uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
            const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            const uint32_t t2 = S0 + mj;
            h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

// st = SHA-256 state after the whole padded message: message, 0x80, zeros, 64-bit big-endian bit count
B2R_HD void sha256_message(const uint8_t* m, uint64_t len, uint32_t st[8]) {
    const uint32_t iv[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
    for (int i = 0; i < 8; i++) st[i] = iv[i];
    uint32_t w[16];
    const uint64_t nblocks = (len + 9 + 63) / 64;
    for (uint64_t blk = 0; blk < nblocks; blk++) {
        const uint64_t p0 = blk * 64;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            uint32_t word = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const uint64_t p = p0 + 4 * j + b;
                uint32_t byte = 0;
                if (p < len) byte = m[p];
                else if (p == len) byte = 0x80;
                word = (word << 8) | byte;
            }
            w[j] = word;
        }
        if (blk == nblocks - 1) {
            const uint64_t bits = len * 8;
            w[14] = (uint32_t)(bits >> 32);
            w[15] = (uint32_t)bits;
        }
        sha256_block(st, w);
    }
}
// limb j = bits [64j, 64j + 64) of the digest read as a big-endian integer (what the reference composes from the reversed
// digest bytes, src/lib.rs:211-236)
B2R_HD void sha256_state_to_limbs(const uint32_t st[8], uint64_t limbs[4]) {
    for (int j = 0; j < 4; j++) limbs[j] = ((uint64_t)st[6 - 2 * j] << 32) | st[7 - 2 * j];
}
B2R_HD void sha256_state_to_digest(const uint32_t st[8], uint8_t digest[32]) {
    for (int j = 0; j < 8; j++) {
        digest[4 * j] = (uint8_t)(st[j] >> 24);
        digest[4 * j + 1] = (uint8_t)(st[j] >> 16);
        digest[4 * j + 2] = (uint8_t)(st[j] >> 8);
        digest[4 * j + 3] = (uint8_t)st[j];
    }
}

}  // namespace b2r
