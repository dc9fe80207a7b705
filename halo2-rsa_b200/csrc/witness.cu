// placeholder until the witness path lands (replaced in the next milestone)
#include "ctx.hpp"
using namespace b2r;
extern "C" {
int32_t b2r_rsa_program_build(b2r_ctx* ctx, uint32_t, const uint8_t*, size_t, uint32_t, b2r_prog**) { return fail(ctx, B2R_ERR_INVALID, "witness path not built"); }
int32_t b2r_prog_free(b2r_ctx* ctx, b2r_prog*) { return fail(ctx, B2R_ERR_INVALID, "witness path not built"); }
int32_t b2r_prog_info(const b2r_prog*, uint64_t*, uint64_t*, uint64_t*) { return B2R_ERR_INVALID; }
int32_t b2r_rsa_witness_batch(b2r_ctx* ctx, const b2r_prog*, const uint64_t*, const uint64_t*, const uint64_t*, size_t, uint64_t, b2r_fr*, uint8_t*) { return fail(ctx, B2R_ERR_INVALID, "witness path not built"); }
int32_t b2r_rsa_witness_batch_dev(b2r_ctx* ctx, const b2r_prog*, const uint64_t*, const uint64_t*, const uint64_t*, size_t, uint64_t, b2r_fr*, uint8_t*) { return fail(ctx, B2R_ERR_INVALID, "witness path not built"); }
}
