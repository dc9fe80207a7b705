// SHA-256 of a batch of messages on the device: the value-level front end of
// RSASignatureVerifier::verify_pkcs1v15_signature (reference src/lib.rs:183-248, step 1: `sha256.finalize(msg)`,
// `decompose_digest_to_bytes`, `hashed_bytes.reverse()`), i.e. message bytes -> the 32 digest bytes -> the four 64-bit
// limbs that the recorded digest-tail program (b2r_rsa_program_build_sha_tail) composes from its byte cells.
// The reference's benchmark and tests obtain the same digest from the `sha2` crate (benches/bench.rs:255-268,
// src/lib.rs test module): FIPS 180-4, restated here from the standard.
//
// The constraint layout of the compression function lives in halo2-dynamic-sha256 (Cargo.toml:15, no rev, not
// vendored) and is NOT reproduced: this kernel produces the values of the digest-byte cells.
//
// One thread per message: messages of this circuit are tens to hundreds of bytes (the bench uses 32 ... 128+64), a batch
// is independent instances, and the cost is nanoseconds next to the 19-modmul witness chain behind it.
#include <cuda_runtime.h>

#include <cstring>

#include "ctx.hpp"

namespace b2r {

__constant__ uint32_t SHA_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

__device__ __forceinline__ uint32_t rotr(uint32_t x, int r) { return __funnelshift_r(x, x, r); }

// one 64-byte block; w[16] holds the big-endian message words and is used as the rolling schedule
__device__ void sha256_block(uint32_t st[8], uint32_t w[16]) {
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll 1
    for (int t0 = 0; t0 < 64; t0 += 16) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (t0) {
                const uint32_t w15 = w[(j + 1) & 15], w2 = w[(j + 14) & 15];
                const uint32_t s0 = rotr(w15, 7) ^ rotr(w15, 18) ^ (w15 >> 3);
                const uint32_t s1 = rotr(w2, 17) ^ rotr(w2, 19) ^ (w2 >> 10);
                w[j] = w[j] + s0 + w[(j + 9) & 15] + s1;
            }
            const uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
            const uint32_t ch = (e & f) ^ (~e & g);
            const uint32_t t1 = h + S1 + ch + SHA_K[t0 + j] + w[j];
            const uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
            const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            const uint32_t t2 = S0 + mj;
            h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

// msgs: concatenated message bytes, message i = [offsets[i], offsets[i + 1]).
// hash_limbs (optional): 4 words per instance at stride `limb_stride`, limb j = bits [64j, 64j + 64) of the digest read as
//   a big-endian integer (what the reference composes from the reversed digest bytes, src/lib.rs:211-236).
// digests (optional): 32 bytes per instance in SHA order (the `hashed_bytes` the reference returns, src/lib.rs:246-247).
__global__ void k_sha256_msgs(const uint8_t* __restrict__ msgs, const uint64_t* __restrict__ offsets, uint32_t batch, uint64_t* hash_limbs,
                              uint32_t limb_stride, uint8_t* digests) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    const uint64_t beg = offsets[i], end = offsets[i + 1];
    const uint64_t len = end > beg ? end - beg : 0;
    const uint8_t* m = msgs + beg;
    uint32_t st[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    uint32_t w[16];
    // padded length: message, 0x80, zeros, 64-bit big-endian bit count, a multiple of 64 bytes
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
    if (hash_limbs) {
        uint64_t* o = hash_limbs + (uint64_t)i * limb_stride;
        for (int j = 0; j < 4; j++) o[j] = ((uint64_t)st[6 - 2 * j] << 32) | st[7 - 2 * j];
    }
    if (digests) {
        uint8_t* o = digests + (uint64_t)i * 32;
        for (int j = 0; j < 8; j++) {
            o[4 * j] = (uint8_t)(st[j] >> 24);
            o[4 * j + 1] = (uint8_t)(st[j] >> 16);
            o[4 * j + 2] = (uint8_t)(st[j] >> 8);
            o[4 * j + 3] = (uint8_t)st[j];
        }
    }
}

// offsets must be non-decreasing and end inside the message buffer: checked on the host copy by the callers that have one
int32_t sha256_msgs_dev(b2r_ctx* ctx, const uint8_t* d_msgs, const uint64_t* d_offsets, size_t batch, uint64_t* d_hash_limbs, uint32_t limb_stride,
                        uint8_t* d_digests) {
    if (batch == 0) return 0;
    if (batch > 0x7fffffffull) return fail(ctx, B2R_ERR_INVALID, "sha256: batch too large");
    {
        KTimer kt(ctx, "sha256", (double)batch);
        k_sha256_msgs<<<(unsigned)((batch + 63) / 64), 64, 0, ctx->stream>>>(d_msgs, d_offsets, (uint32_t)batch, d_hash_limbs, limb_stride, d_digests);
    }
    B2R_LAUNCH_CHECK(ctx);
    return 0;
}

int32_t sha256_check_offsets(b2r_ctx* ctx, const uint64_t* offsets, size_t batch) {
    for (size_t i = 0; i < batch; i++)
        if (offsets[i + 1] < offsets[i]) return fail(ctx, B2R_ERR_INVALID, "sha256: message offsets must be non-decreasing");
    return 0;
}

}  // namespace b2r

using namespace b2r;

extern "C" {

int32_t b2r_sha256_batch_dev(b2r_ctx* ctx, const uint8_t* msgs_dev, const uint64_t* offsets_dev, size_t batch, uint64_t* hash_limbs_dev,
                             uint8_t* digests_dev) try {
    B2R_ENTER(ctx);
    if (!offsets_dev || (!hash_limbs_dev && !digests_dev)) return fail(ctx, B2R_ERR_INVALID, "sha256: null pointer");
    return sha256_msgs_dev(ctx, msgs_dev, offsets_dev, batch, hash_limbs_dev, 4, digests_dev);
} B2R_ABI_CATCH(ctx)

int32_t b2r_sha256_batch(b2r_ctx* ctx, const uint8_t* msgs, const uint64_t* offsets, size_t batch, uint64_t* hash_limbs, uint8_t* digests) try {
    B2R_ENTER(ctx);
    if (!offsets || (!hash_limbs && !digests)) return fail(ctx, B2R_ERR_INVALID, "sha256: null pointer");
    if (batch == 0) return 0;
    B2R_TRY(sha256_check_offsets(ctx, offsets, batch));
    const uint64_t base = offsets[0], total = offsets[batch] - base;
    if (total && !msgs) return fail(ctx, B2R_ERR_INVALID, "sha256: null message buffer");
    // staging: offsets (rebased to the staged copy) | limbs | digests | message bytes
    const size_t off_bytes = ((batch + 1) * 8 + 255) & ~(size_t)255, limb_bytes = (batch * 32 + 255) & ~(size_t)255;
    char* d = nullptr;
    B2R_TRY(scratch_get(ctx, SC_MISC, off_bytes + 2 * limb_bytes + total + 256, (void**)&d));
    uint64_t* d_off = (uint64_t*)d;
    uint64_t* d_limbs = (uint64_t*)(d + off_bytes);
    uint8_t* d_dig = (uint8_t*)(d + off_bytes + limb_bytes);
    uint8_t* d_msgs = (uint8_t*)(d + off_bytes + 2 * limb_bytes);
    std::vector<uint64_t> rel(batch + 1);
    for (size_t i = 0; i <= batch; i++) rel[i] = offsets[i] - base;
    B2R_CUDA(ctx, cudaMemcpyAsync(d_off, rel.data(), (batch + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (total) B2R_CUDA(ctx, cudaMemcpyAsync(d_msgs, msgs + base, total, cudaMemcpyHostToDevice, ctx->stream));
    B2R_TRY(sha256_msgs_dev(ctx, d_msgs, d_off, batch, d_limbs, 4, d_dig));
    if (hash_limbs) B2R_CUDA(ctx, cudaMemcpyAsync(hash_limbs, d_limbs, batch * 32, cudaMemcpyDeviceToHost, ctx->stream));
    if (digests) B2R_CUDA(ctx, cudaMemcpyAsync(digests, d_dig, batch * 32, cudaMemcpyDeviceToHost, ctx->stream));
    B2R_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // `rel` is pageable host memory read by the copy above
    return 0;
} B2R_ABI_CATCH(ctx)

}  // extern "C"
