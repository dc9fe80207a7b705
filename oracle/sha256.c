/* TEST INFRASTRUCTURE (oracle): SHA-256 restated from FIPS 180-4 - the digest the reference takes from the `sha2` crate
 * (benches/bench.rs:255-268, src/lib.rs test module) and that RSASignatureVerifier::verify_pkcs1v15_signature
 * (src/lib.rs:204-211) obtains from halo2-dynamic-sha256's chip.  Pinned by the FIPS 180-4 / NIST example vectors and
 * Python's hashlib in tests/test_oracle_sha256.py.  Written independently of the device kernel (csrc/sha256.cu):
 * full 64-word schedule, byte-buffer padding. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const uint32_t K256[64] = {
    0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u, 0xd807aa98u, 0x12835b01u,
    0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u, 0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu,
    0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau, 0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u,
    0x06ca6351u, 0x14292967u, 0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
    0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u, 0x19a4c116u, 0x1e376c08u,
    0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u, 0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u,
    0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};

#define ROR(x, n) (((x) >> (n)) | ((x) << (32 - (n))))

static void compress(uint32_t H[8], const uint8_t* blk) {
    uint32_t W[64];
    for (int t = 0; t < 16; t++) W[t] = ((uint32_t)blk[4 * t] << 24) | ((uint32_t)blk[4 * t + 1] << 16) | ((uint32_t)blk[4 * t + 2] << 8) | blk[4 * t + 3];
    for (int t = 16; t < 64; t++) {
        uint32_t s0 = ROR(W[t - 15], 7) ^ ROR(W[t - 15], 18) ^ (W[t - 15] >> 3);
        uint32_t s1 = ROR(W[t - 2], 17) ^ ROR(W[t - 2], 19) ^ (W[t - 2] >> 10);
        W[t] = s1 + W[t - 7] + s0 + W[t - 16];
    }
    uint32_t v[8];
    memcpy(v, H, sizeof v);
    for (int t = 0; t < 64; t++) {
        uint32_t T1 = v[7] + (ROR(v[4], 6) ^ ROR(v[4], 11) ^ ROR(v[4], 25)) + ((v[4] & v[5]) ^ (~v[4] & v[6])) + K256[t] + W[t];
        uint32_t T2 = (ROR(v[0], 2) ^ ROR(v[0], 13) ^ ROR(v[0], 22)) + ((v[0] & v[1]) ^ (v[0] & v[2]) ^ (v[1] & v[2]));
        memmove(v + 1, v, 7 * sizeof(uint32_t));
        v[4] += T1;
        v[0] = T1 + T2;
    }
    for (int i = 0; i < 8; i++) H[i] += v[i];
}

/* digest: 32 bytes in SHA order */
void orc_sha256(const uint8_t* msg, uint64_t len, uint8_t* digest) {
    uint32_t H[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
    uint64_t padded = ((len + 8) / 64 + 1) * 64;
    uint8_t* buf = (uint8_t*)calloc(padded, 1);
    if (len) memcpy(buf, msg, len);
    buf[len] = 0x80;
    for (int i = 0; i < 8; i++) buf[padded - 1 - i] = (uint8_t)((len * 8) >> (8 * i));
    for (uint64_t off = 0; off < padded; off += 64) compress(H, buf + off);
    free(buf);
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 4; j++) digest[4 * i + j] = (uint8_t)(H[i] >> (24 - 8 * j));
}

/* src/lib.rs:210-211 + :222-236: the digest bytes reversed (least significant first) are the byte cells; limb i is
 * composed from cells 8i .. 8i+7 with coefficients 2^(8j).  digest_le: 32 bytes, limbs: 4 words. */
void orc_sha256_hashed_limbs(const uint8_t* msg, uint64_t len, uint8_t* digest_le, uint64_t* limbs) {
    uint8_t d[32];
    orc_sha256(msg, len, d);
    for (int i = 0; i < 32; i++) digest_le[i] = d[31 - i];
    for (int i = 0; i < 4; i++) {
        uint64_t v = 0;
        for (int j = 0; j < 8; j++) v += (uint64_t)digest_le[8 * i + j] << (8 * j);
        limbs[i] = v;
    }
}
