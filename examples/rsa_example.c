/* rsa_example.c - the reference's examples/rsa_example.rs flow driven through the C ABI from plain C:
 * message bytes + pkcs1v15 signature + RSA public key  ->  SHA-256 on the device  ->  witness  ->  proof.
 * (examples/rsa_example.rs:150-212 samples a key, signs SHA-256(msg), builds the circuit and runs MockProver at k = 18;
 * here the inputs come from a text file, the circuit is the recorded RSASignatureVerifier digest tail, and the output is
 * real proofs plus the verifying key, which tests/test_gpu_example.py checks with the oracle verifier.)
 *
 * build:  gcc -O2 -std=c11 -Iinclude examples/rsa_example.c -Lhalo2-rsa_b200/lib -lb2rsa -Wl,-rpath,$PWD/halo2-rsa_b200/lib -o rsa_example
 * run:    ./rsa_example <bits> <k> <inputs.txt> <out.bin>
 *   inputs.txt: one instance per line, three hex fields: n  signature  message-bytes
 *   out.bin:    u32 batch | u32 proof_bytes | u32 num_fixed | u32 num_sigma | status[batch] | digests[batch][32] |
 *               proofs[batch][proof_bytes] | fixed commitments | sigma commitments | transcript_repr
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b2rsa.h"

#define CHECK(call)                                                                          \
    do {                                                                                     \
        int32_t rc_ = (call);                                                                \
        if (rc_ != B2R_OK) {                                                                 \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, b2r_last_error(ctx));              \
            return 1;                                                                        \
        }                                                                                    \
    } while (0)

static int hexval(int c) {
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    return -1;
}
/* big-endian hex integer -> little-endian 64-bit limbs (decompose_big, benches/bench.rs:280,286) */
static int hex_to_limbs(const char* s, uint64_t* limbs, size_t nl) {
    size_t len = strlen(s);
    memset(limbs, 0, nl * 8);
    for (size_t i = 0; i < len; i++) {
        int v = hexval(s[len - 1 - i]);
        if (v < 0 || i / 16 >= nl) return -1;
        limbs[i / 16] |= (uint64_t)v << (4 * (i % 16));
    }
    return 0;
}

int main(int argc, char** argv) {
    if (argc != 5) {
        fprintf(stderr, "usage: %s <bits> <k> <inputs.txt> <out.bin>\n", argv[0]);
        return 2;
    }
    const uint32_t bits = (uint32_t)atoi(argv[1]), k = (uint32_t)atoi(argv[2]);
    const size_t nl = bits / 64;
    b2r_ctx* ctx = NULL;
    if (b2r_ctx_create(0, &ctx) != B2R_OK) {
        fprintf(stderr, "b2r_ctx_create: %s\n", b2r_last_error(NULL));   /* no sm_100 device: there is no CPU fallback */
        return 1;
    }
    /* ---- read the instances */
    FILE* f = fopen(argv[3], "r");
    if (!f) { perror(argv[3]); return 2; }
    size_t cap = 16, batch = 0, msg_cap = 1 << 16, msg_len = 0;
    uint64_t* n_limbs = malloc(cap * nl * 8);
    uint64_t* s_limbs = malloc(cap * nl * 8);
    uint64_t* offs = malloc((cap + 1) * 8);
    uint8_t* msgs = malloc(msg_cap);
    static char line[1 << 16], a[1 << 14], b[1 << 14], m[1 << 15];
    offs[0] = 0;
    while (fgets(line, sizeof line, f)) {
        m[0] = 0;
        int got = sscanf(line, "%16383s %16383s %32767s", a, b, m);
        if (got < 2) continue;
        if (batch == cap) {
            cap *= 2;
            n_limbs = realloc(n_limbs, cap * nl * 8);
            s_limbs = realloc(s_limbs, cap * nl * 8);
            offs = realloc(offs, (cap + 1) * 8);
        }
        if (hex_to_limbs(a, n_limbs + batch * nl, nl) || hex_to_limbs(b, s_limbs + batch * nl, nl)) {
            fprintf(stderr, "line %zu: bad hex or more than %u bits\n", batch + 1, bits);
            return 2;
        }
        const size_t ml = got == 3 ? strlen(m) / 2 : 0;   /* a missing third field is the empty message */
        if (msg_len + ml > msg_cap) { msg_cap = 2 * (msg_len + ml); msgs = realloc(msgs, msg_cap); }
        for (size_t i = 0; i < ml; i++) msgs[msg_len + i] = (uint8_t)(hexval(m[2 * i]) << 4 | hexval(m[2 * i + 1]));
        msg_len += ml;
        offs[++batch] = msg_len;
    }
    fclose(f);
    if (!batch) { fprintf(stderr, "no instances\n"); return 2; }

    /* ---- setup: SRS (ParamsKZG::setup with a fixed secret: a demo, not a ceremony), the recorded circuit, keygen */
    /* Montgomery form of the secret 0xB200: (0xB200 * 2^256) mod r, little-endian limbs */
    const b2r_fr secret = {{0xababee764ffc525bull, 0x49b8e1dae8b47bd9ull, 0x2403375615d39f2eull, 0x1afd2a31714a6a9dull}};
    b2r_bases *g = NULL, *gl = NULL;
    CHECK(b2r_srs_setup(ctx, k, &secret, &g, &gl));
    const uint8_t e_le[3] = {0x01, 0x00, 0x01};   /* 65537 */
    b2r_prog* prog = NULL;
    CHECK(b2r_rsa_program_build_sha_tail(ctx, bits, e_le, sizeof e_le, k, &prog));
    uint64_t rows = 0, nvals = 0, levels = 0;
    CHECK(b2r_prog_info(prog, &rows, &nvals, &levels));
    b2r_pk* pk = NULL;
    CHECK(b2r_rsa_keygen(ctx, prog, g, gl, &pk));
    uint32_t kk = 0, ext_k = 0, nfixed = 0, nsigma = 0;
    uint64_t proof_bytes = 0;
    CHECK(b2r_pk_info(pk, &kk, &ext_k, &nfixed, &nsigma, &proof_bytes));

    /* ---- message bytes -> proofs (create_proof per instance, benches/bench.rs:319-331) */
    uint8_t* proofs = malloc(batch * proof_bytes);
    uint8_t* status = malloc(batch);
    uint8_t* digests = malloc(batch * 32);
    uint8_t seed32[32];
    FILE* ur = fopen("/dev/urandom", "rb");   /* the blinding key: what OsRng is to the reference */
    if (!ur || fread(seed32, 1, 32, ur) != 32) { fprintf(stderr, "no /dev/urandom\n"); return 1; }
    fclose(ur);
    CHECK(b2r_rsa_prove_msgs_batch(ctx, pk, n_limbs, s_limbs, msgs, offs, batch, seed32, /*nonce=*/1, 0, proofs, status, digests));

    b2r_g1_affine* fixed = malloc(nfixed * sizeof(b2r_g1_affine));
    b2r_g1_affine* sigma = malloc(nsigma * sizeof(b2r_g1_affine));
    b2r_fr repr;
    CHECK(b2r_pk_export_vk(pk, fixed, sigma, &repr));

    size_t valid = 0;
    for (size_t i = 0; i < batch; i++) valid += status[i] == 1;
    printf("rsa_example: RSA-%u k=%u rows=%llu batch=%zu valid=%zu proof_bytes=%llu launches=%llu\n", bits, k, (unsigned long long)rows, batch, valid,
           (unsigned long long)proof_bytes, (unsigned long long)b2r_launch_count(ctx));
    for (size_t i = 0; i < batch; i++) {
        printf("  [%zu] status=%u sha256=", i, status[i]);
        for (int j = 0; j < 32; j++) printf("%02x", digests[i * 32 + j]);
        printf("\n");
    }
    FILE* o = fopen(argv[4], "wb");
    if (!o) { perror(argv[4]); return 2; }
    const uint32_t hdr[4] = {(uint32_t)batch, (uint32_t)proof_bytes, nfixed, nsigma};
    fwrite(hdr, 4, 4, o);
    fwrite(status, 1, batch, o);
    fwrite(digests, 32, batch, o);
    fwrite(proofs, proof_bytes, batch, o);
    fwrite(fixed, sizeof(b2r_g1_affine), nfixed, o);
    fwrite(sigma, sizeof(b2r_g1_affine), nsigma, o);
    fwrite(&repr, sizeof repr, 1, o);
    fclose(o);

    CHECK(b2r_pk_free(ctx, pk));
    CHECK(b2r_prog_free(ctx, prog));
    CHECK(b2r_bases_free(ctx, g));
    CHECK(b2r_bases_free(ctx, gl));
    b2r_ctx_destroy(ctx);
    return 0;
}
