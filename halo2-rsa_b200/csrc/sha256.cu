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
#include "sha256.cuh"

namespace b2r {

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
    uint32_t st[8];
    sha256_message(msgs + beg, len, st);
    if (hash_limbs) sha256_state_to_limbs(st, hash_limbs + (uint64_t)i * limb_stride);
    if (digests) sha256_state_to_digest(st, digests + (uint64_t)i * 32);
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
