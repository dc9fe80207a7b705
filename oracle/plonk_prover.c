/* TEST INFRASTRUCTURE ONLY - CPU oracle: create_proof for the RSA circuit's constraint system, in C.
 *
 * The same restatement of halo2_proofs' create_proof (privacy-scaling-explorations/halo2, 2022-10 era;
 * third-party, not vendored under /root/reference; reference call site benches/bench.rs:319-331) as
 * oracle/plonk.py::create_proof, statement by statement, but on the C oracle's field arithmetic and
 * threaded loops so that it finishes in seconds at the BASELINE sizes (k = 17 / 18) where the Python
 * loops need many minutes.  tests/test_oracle_plonk_c.py pins it to plonk.py byte for byte at k = 14;
 * tests/test_gpu_prover.py then compares the product's proofs with it at k = 17 and k = 18.
 * It is also the complete CPU prover that bench.py times as the reference arm ("port").
 *
 * Deliberately written the way halo2 evaluates things - Horner in y over the constraints in their
 * declaration order, one permutation set after the other, per-query Horner in v - and NOT the way
 * the product's kernels regroup them (csrc/prover.cu: l-polynomial grouping, precomputed y powers,
 * linear-combination tables), so agreement of the proof bytes is agreement of two formulations.
 * "parity unpinned" against the reference itself: see oracle/plonk.py and oracle/EXT_ASSUMPTIONS.md.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>

#include "fields.h"

/* oracle/poly.c */
void orc_best_fft(fe* a, const fe* omega, unsigned log_n, int nthreads);
fe orc_omega(unsigned k);
void orc_lagrange_to_coeff(fe* a, unsigned k, int nthreads);
void orc_coeff_to_extended(const fe* coeffs, unsigned k, unsigned ext_k, fe* out, int nthreads);
void orc_extended_to_coeff(fe* a, unsigned ext_k, int nthreads);
void orc_best_multiexp(const fe* coeffs, const g1a* bases, size_t n, int nthreads, g1a* out);

/* ---- constraint system (oracle/plonk.py constants) ---------------------------------------------- */
enum { NADV = 5, NFIX = 15, NPERM = 6, NLOOK = 5, CHUNK = 3, NSETS = 2, BF = 5, QD = 4 };
enum { F_SA = 0, F_SB, F_SC, F_SD, F_SE, F_MUL_AB, F_MUL_CD, F_SE_NEXT, F_CONST, F_TAG_COMP, F_TAG_OVER, F_T_TAG, F_T_VALUE, F_S_COMP, F_S_OVER };
enum { ST_ADVICE = 0, ST_LOOKUP_A = 8, ST_LOOKUP_S = 16, ST_LOOKUP_Z = 24, ST_PERM_Z = 32, ST_RANDOM_POLY = 40 };
static const int LK_ACOL[NLOOK] = {0, 1, 2, 3, 0};
static const int LK_FTAG[NLOOK] = {F_TAG_COMP, F_TAG_COMP, F_TAG_COMP, F_TAG_COMP, F_TAG_OVER};
static const int LK_FSEL[NLOOK] = {F_S_COMP, F_S_COMP, F_S_COMP, F_S_COMP, F_S_OVER};
#define PROOF_BYTES (32 * (NADV + 2 * NLOOK + NSETS + NLOOK + 1 + QD + 4 + 6 + NFIX + 1 + NPERM + (3 * NSETS - 1) + 5 * NLOOK))

typedef struct {
    uint32_t k;
    uint32_t table_len;
    const fe* fixed_values; /* [NFIX][n]   */
    const fe* fixed_polys;  /* [NFIX][n]   */
    const fe* fixed_cosets; /* [NFIX][4n]  */
    const fe* sigma_values; /* [NPERM][n]  */
    const fe* sigma_polys;
    const fe* sigma_cosets;
    const fe* l0;           /* [4n] */
    const fe* l_last;
    const fe* l_active;
    const g1a* g;           /* [n] */
    const g1a* g_lagrange;  /* [n] */
    fe transcript_repr;
} orc_plonk_key;

#define FRM(a, b) fe_mul(&FR, &(a), &(b))
#define FRA(a, b) fe_add(&FR, &(a), &(b))
#define FRS(a, b) fe_sub(&FR, &(a), &(b))

/* ---- fork-join ------------------------------------------------------------------------------------ */
typedef void (*pjob_fn)(void* arg, size_t lo, size_t hi);
typedef struct { pjob_fn fn; void* arg; size_t lo, hi; } pjob;
static void* pjob_tramp(void* p) { pjob* j = (pjob*)p; j->fn(j->arg, j->lo, j->hi); return 0; }
static void par_range(pjob_fn fn, void* arg, size_t n, int nthreads) {
    if (nthreads <= 1 || n < 1024) { fn(arg, 0, n); return; }
    pthread_t th[256];
    pjob jb[256];
    if (nthreads > 256) nthreads = 256;
    for (int i = 0; i < nthreads; i++) {
        jb[i].fn = fn; jb[i].arg = arg; jb[i].lo = n * (size_t)i / nthreads; jb[i].hi = n * (size_t)(i + 1) / nthreads;
        pthread_create(&th[i], 0, pjob_tramp, &jb[i]);
    }
    for (int i = 0; i < nthreads; i++) pthread_join(th[i], 0);
}

/* ---- Blake2b-512 (RFC 7693) with a 16-byte personalisation ------------------------------------- */
typedef struct { uint64_t h[8], t[2]; uint8_t buf[128]; size_t len; } b2b;
static const uint64_t B2B_IV[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                                   0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
static const uint8_t B2B_SIGMA[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
static uint64_t ror64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
static void b2b_compress(b2b* S, const uint8_t* blk, int last) {
    uint64_t m[16], v[16];
    memcpy(m, blk, 128); /* little-endian host */
    for (int i = 0; i < 8; i++) { v[i] = S->h[i]; v[i + 8] = B2B_IV[i]; }
    v[12] ^= S->t[0]; v[13] ^= S->t[1];
    if (last) v[14] = ~v[14];
#define G(r, i, a, b, c, d) do { a = a + b + m[B2B_SIGMA[r][2 * i]]; d = ror64(d ^ a, 32); c = c + d; b = ror64(b ^ c, 24); \
                                 a = a + b + m[B2B_SIGMA[r][2 * i + 1]]; d = ror64(d ^ a, 16); c = c + d; b = ror64(b ^ c, 63); } while (0)
    for (int r = 0; r < 12; r++) {
        G(r, 0, v[0], v[4], v[8], v[12]); G(r, 1, v[1], v[5], v[9], v[13]); G(r, 2, v[2], v[6], v[10], v[14]); G(r, 3, v[3], v[7], v[11], v[15]);
        G(r, 4, v[0], v[5], v[10], v[15]); G(r, 5, v[1], v[6], v[11], v[12]); G(r, 6, v[2], v[7], v[8], v[13]); G(r, 7, v[3], v[4], v[9], v[14]);
    }
#undef G
    for (int i = 0; i < 8; i++) S->h[i] ^= v[i] ^ v[i + 8];
}
static void b2b_init(b2b* S, const char person[16]) {
    uint8_t p[64];
    memset(p, 0, 64);
    p[0] = 64; p[2] = 1; p[3] = 1;
    memcpy(p + 48, person, 16);
    uint64_t pw[8];
    memcpy(pw, p, 64);
    for (int i = 0; i < 8; i++) S->h[i] = B2B_IV[i] ^ pw[i];
    S->t[0] = S->t[1] = 0; S->len = 0;
}
static void b2b_update(b2b* S, const uint8_t* in, size_t n) {
    while (n) {
        if (S->len == 128) {
            S->t[0] += 128; if (S->t[0] < 128) S->t[1]++;
            b2b_compress(S, S->buf, 0);
            S->len = 0;
        }
        size_t take = 128 - S->len;
        if (take > n) take = n;
        memcpy(S->buf + S->len, in, take);
        S->len += take; in += take; n -= take;
    }
}
static void b2b_final_copy(const b2b* S0, uint8_t out[64]) { /* finalises a CLONE: the transcript keeps absorbing */
    b2b S = *S0;
    S.t[0] += S.len; if (S.t[0] < S.len) S.t[1]++;
    memset(S.buf + S.len, 0, 128 - S.len);
    b2b_compress(&S, S.buf, 1);
    memcpy(out, S.h, 64);
}

/* 512-bit little-endian integer mod r -> Montgomery (halo2curves from_bytes_wide / from_u512: d0 * R2 + d1 * R3) */
static fe fr_from_wide(const uint8_t in[64]) {
    fe d0, d1;
    memcpy(d0.l, in, 32);
    memcpy(d1.l, in + 32, 32);
    for (int it = 0; it < 5; it++) {
        if (geq_mod(d0.l, FR.mod)) sub_mod_raw(d0.l, FR.mod);
        if (geq_mod(d1.l, FR.mod)) sub_mod_raw(d1.l, FR.mod);
    }
    fe r3 = fe_mul(&FR, &FR.r2, &FR.r2);
    fe a = fe_mul(&FR, &d0, &FR.r2), b = fe_mul(&FR, &d1, &r3);
    return fe_add(&FR, &a, &b);
}

/* ---- transcript: Blake2bWrite<_, G1Affine, Challenge255<_>> ------------------------------------ */
typedef struct { b2b st; uint8_t* out; size_t pos; } transcript;
static void tr_init(transcript* T, uint8_t* out) { b2b_init(&T->st, "Halo2-Transcript"); T->out = out; T->pos = 0; }
static void tr_common_scalar(transcript* T, const fe* s) {
    uint8_t b[33];
    uint64_t c[4];
    b[0] = 2;
    fe_from_mont(&FR, s, c);
    memcpy(b + 1, c, 32);
    b2b_update(&T->st, b, 33);
}
static void tr_write_scalar(transcript* T, const fe* s) {
    uint64_t c[4];
    tr_common_scalar(T, s);
    fe_from_mont(&FR, s, c);
    memcpy(T->out + T->pos, c, 32);
    T->pos += 32;
}
static void tr_write_point(transcript* T, const g1a* P) { /* identity = (0, 0) -> 32 zero bytes */
    uint8_t b[65];
    uint64_t x[4], y[4];
    b[0] = 1;
    fe_from_mont(&FQ, &P->x, x);
    fe_from_mont(&FQ, &P->y, y);
    memcpy(b + 1, x, 32);
    memcpy(b + 33, y, 32);
    b2b_update(&T->st, b, 65);
    uint8_t c[32];
    memcpy(c, x, 32);
    c[31] |= (uint8_t)((y[0] & 1) << 7); /* sign of y in the top bit of the last byte */
    memcpy(T->out + T->pos, c, 32);
    T->pos += 32;
}
static fe tr_squeeze(transcript* T) {
    uint8_t z = 0, d[64];
    b2b_update(&T->st, &z, 1);
    b2b_final_copy(&T->st, d);
    return fr_from_wide(d);
}

/* ---- blinding stream v2: ChaCha20 block per cell (product: csrc/devutil.cuh; oracle/plonk.py blind_fes) ------- */
typedef struct { uint32_t key[8]; uint32_t nonce[2]; } blind_key;
static uint32_t rol32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
#define QR(a, b, c, d) do { a += b; d ^= a; d = rol32(d, 16); c += d; b ^= c; b = rol32(b, 12); a += b; d ^= a; d = rol32(d, 8); c += d; b ^= c; b = rol32(b, 7); } while (0)
static fe blind_fe(const blind_key* K, uint32_t proof, uint32_t stream, uint32_t row) {
    uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, K->key[0], K->key[1], K->key[2], K->key[3], K->key[4], K->key[5],
                      K->key[6], K->key[7], row | (stream << 24), proof, K->nonce[0], K->nonce[1]};
    uint32_t x[16];
    memcpy(x, s, 64);
    for (int r = 0; r < 10; r++) {
        QR(x[0], x[4], x[8], x[12]); QR(x[1], x[5], x[9], x[13]); QR(x[2], x[6], x[10], x[14]); QR(x[3], x[7], x[11], x[15]);
        QR(x[0], x[5], x[10], x[15]); QR(x[1], x[6], x[11], x[12]); QR(x[2], x[7], x[8], x[13]); QR(x[3], x[4], x[9], x[14]);
    }
    for (int i = 0; i < 16; i++) x[i] += s[i];
    uint8_t b[64];
    memcpy(b, x, 64);
    return fr_from_wide(b);
}
#undef QR

/* ---- small vector helpers ---------------------------------------------------------------------------- */
static fe* fe_alloc(size_t n) {
    fe* p = (fe*)malloc(n * sizeof(fe));
    if (!p) { fprintf(stderr, "plonk_prover: out of memory\n"); abort(); }
    return p;
}
static int canon_cmp(const void* a, const void* b) { /* canonical 256-bit values, most significant limb first */
    const uint64_t* x = (const uint64_t*)a; const uint64_t* y = (const uint64_t*)b;
    for (int i = 3; i >= 0; i--) { if (x[i] < y[i]) return -1; if (x[i] > y[i]) return 1; }
    return 0;
}
/* Montgomery batch inversion as halo2's BatchInvert: zeros stay zero */
static void batch_invert(fe* v, size_t n) {
    fe* pref = fe_alloc(n);
    fe acc = FR.one;
    for (size_t i = 0; i < n; i++) { pref[i] = acc; if (!fe_is_zero(&v[i])) acc = FRM(acc, v[i]); }
    fe inv = fe_inv(&FR, &acc);
    for (size_t i = n; i-- > 0;) {
        if (fe_is_zero(&v[i])) continue;
        fe o = FRM(inv, pref[i]);
        inv = FRM(inv, v[i]);
        v[i] = o;
    }
    free(pref);
}
static fe eval_poly(const fe* c, size_t n, const fe* x) {
    fe acc = fe_zero();
    for (size_t i = n; i-- > 0;) { acc = FRM(acc, *x); acc = FRA(acc, c[i]); }
    return acc;
}
typedef struct { const fe* c; size_t n; const fe* x; fe* out; size_t chunk; } evalp_job;
static void evalp_fn(void* p, size_t lo, size_t hi) {
    evalp_job* j = (evalp_job*)p;
    for (size_t t = lo; t < hi; t++) {
        size_t a = t * j->chunk, b = a + j->chunk; if (b > j->n) b = j->n;
        j->out[t] = eval_poly(j->c + a, b - a, j->x);
    }
}
/* evaluate in chunks over threads, then combine with x^chunk (plain Horner is one dependent chain of n products) */
static fe eval_poly_par(const fe* c, size_t n, const fe* x, int nthreads) {
    const size_t chunk = 4096;
    if (n <= chunk) return eval_poly(c, n, x);
    size_t nch = (n + chunk - 1) / chunk;
    fe* part = fe_alloc(nch);
    evalp_job j = {c, n, x, part, chunk};
    if (nthreads > 1) {
        /* par_range refuses tiny ranges; call the job over chunks through it with a scaled count */
        pthread_t th[256]; pjob jb[256]; int nt = nthreads > 256 ? 256 : nthreads; if ((size_t)nt > nch) nt = (int)nch;
        for (int i = 0; i < nt; i++) { jb[i].fn = evalp_fn; jb[i].arg = &j; jb[i].lo = nch * (size_t)i / nt; jb[i].hi = nch * (size_t)(i + 1) / nt; pthread_create(&th[i], 0, pjob_tramp, &jb[i]); }
        for (int i = 0; i < nt; i++) pthread_join(th[i], 0);
    } else evalp_fn(&j, 0, nch);
    fe xc = FR.one;
    for (size_t i = 0; i < chunk; i++) xc = FRM(xc, *x);
    fe acc = fe_zero();
    for (size_t t = nch; t-- > 0;) { acc = FRM(acc, xc); acc = FRA(acc, part[t]); }
    free(part);
    return acc;
}
static fe fe_pow_u64(fe a, uint64_t e) {
    fe acc = FR.one;
    while (e) { if (e & 1) acc = FRM(acc, a); a = FRM(a, a); e >>= 1; }
    return acc;
}

/* halo2 lookup::prover::permute_expression_pair on the first u rows; returns 0 on success, -1 when an input value is
 * not in the table (ConstraintSystemFailure).  A, S: Montgomery; outputs a_p, s_p (u values each, Montgomery). */
static int permute_expression_pair(const fe* A, const fe* S, size_t u, fe* a_p, fe* s_p) {
    uint64_t(*ac)[4] = malloc(u * 32), (*sc)[4] = malloc(u * 32);
    size_t* repeated = malloc(u * sizeof(size_t));
    uint64_t(*left)[4] = malloc(u * 32);
    size_t nrep = 0, nleft = 0;
    for (size_t i = 0; i < u; i++) { fe_from_mont(&FR, &A[i], ac[i]); fe_from_mont(&FR, &S[i], sc[i]); }
    qsort(ac, u, 32, canon_cmp);
    qsort(sc, u, 32, canon_cmp);
    size_t sp = 0;
    int rc = 0;
    uint64_t(*spc)[4] = calloc(u, 32);
    for (size_t row = 0; row < u; row++) {
        if (row == 0 || canon_cmp(ac[row], ac[row - 1]) != 0) {
            while (sp < u && canon_cmp(sc[sp], ac[row]) < 0) { memcpy(left[nleft++], sc[sp], 32); sp++; }
            if (sp >= u || canon_cmp(sc[sp], ac[row]) != 0) { rc = -1; break; }
            sp++;
            memcpy(spc[row], ac[row], 32);
        } else {
            repeated[nrep++] = row;
        }
    }
    if (rc == 0) {
        while (sp < u) { memcpy(left[nleft++], sc[sp], 32); sp++; }
        if (nleft != nrep) rc = -2;
        for (size_t i = 0; rc == 0 && i < nleft; i++) memcpy(spc[repeated[--nrep]], left[i], 32); /* ascending leftovers, rows popped from the end */
        for (size_t i = 0; i < u; i++) { a_p[i] = fe_to_mont(&FR, ac[i]); s_p[i] = fe_to_mont(&FR, spc[i]); }
    }
    free(ac); free(sc); free(repeated); free(left); free(spc);
    return rc;
}

/* ---- quotient on the extended coset, Horner in y over halo2's constraint order ---------------------------------- */
typedef struct {
    const orc_plonk_key* key;
    size_t n, ext_n, step;
    const fe *ae[NADV], *pz[NSETS], *lz[NLOOK], *la[NLOOK], *ls[NLOOK];
    fe theta, beta, gamma, y;
    fe delta_pows[NPERM];
    fe t_inv[4];
    fe ext_omega, zeta;
    fe* h;
} quot_job;
static void quot_fn(void* p, size_t lo, size_t hi) {
    quot_job* J = (quot_job*)p;
    const orc_plonk_key* K = J->key;
    const size_t en = J->ext_n, step = J->step;
    const fe one = FR.one, y = J->y, beta = J->beta, gamma = J->gamma, theta = J->theta;
#define FX(c) (K->fixed_cosets[(size_t)(c) * en + i])
    fe xpt = FRM(J->zeta, *(fe[]){fe_pow_u64(J->ext_omega, lo)});
    for (size_t i = lo; i < hi; i++) {
        const size_t nx = (i + step) % en, pv = (i + en - step) % en, lastr = (i + en - (BF + 1) * step) % en;
        fe colv[NPERM];
        for (int c = 0; c < NADV; c++) colv[c] = J->ae[c][i];
        colv[NADV] = fe_zero(); /* instance column: no public inputs */
        fe acc, t, u;
        /* main gate */
        acc = FRM(colv[0], FX(F_SA));
        t = FRM(colv[1], FX(F_SB)); acc = FRA(acc, t);
        t = FRM(colv[2], FX(F_SC)); acc = FRA(acc, t);
        t = FRM(colv[3], FX(F_SD)); acc = FRA(acc, t);
        t = FRM(colv[4], FX(F_SE)); acc = FRA(acc, t);
        t = FRM(colv[0], colv[1]); t = FRM(t, FX(F_MUL_AB)); acc = FRA(acc, t);
        t = FRM(colv[2], colv[3]); t = FRM(t, FX(F_MUL_CD)); acc = FRA(acc, t);
        t = FRM(J->ae[4][nx], FX(F_SE_NEXT)); acc = FRA(acc, t);
        acc = FRA(acc, FX(F_CONST));
        const fe l0 = K->l0[i], ll = K->l_last[i], la_ = K->l_active[i];
        /* permutation argument */
        t = FRS(one, J->pz[0][i]); t = FRM(l0, t); acc = FRM(acc, y); acc = FRA(acc, t);
        {
            fe zl = J->pz[NSETS - 1][i];
            t = FRM(zl, zl); t = FRS(t, zl); t = FRM(ll, t); acc = FRM(acc, y); acc = FRA(acc, t);
        }
        for (int s = 1; s < NSETS; s++) {
            t = FRS(J->pz[s][i], J->pz[s - 1][lastr]); t = FRM(l0, t); acc = FRM(acc, y); acc = FRA(acc, t);
        }
        for (int s = 0; s < NSETS; s++) {
            fe left = J->pz[s][nx], right = J->pz[s][i];
            for (int c = s * CHUNK; c < (s + 1) * CHUNK && c < NPERM; c++) {
                t = FRM(beta, K->sigma_cosets[(size_t)c * en + i]); t = FRA(t, colv[c]); t = FRA(t, gamma); left = FRM(left, t);
                u = FRM(beta, xpt); u = FRM(u, J->delta_pows[c]); u = FRA(u, colv[c]); u = FRA(u, gamma); right = FRM(right, u);
            }
            t = FRS(left, right); t = FRM(la_, t); acc = FRM(acc, y); acc = FRA(acc, t);
        }
        /* lookups */
        fe tbl = FRM(FX(F_T_TAG), theta); tbl = FRA(tbl, FX(F_T_VALUE));
        for (int l = 0; l < NLOOK; l++) {
            const fe z = J->lz[l][i], zn = J->lz[l][nx], ap = J->la[l][i], sp = J->ls[l][i], apv = J->la[l][pv];
            fe inp = FRM(FX(LK_FTAG[l]), theta);
            t = FRM(FX(LK_FSEL[l]), colv[LK_ACOL[l]]); inp = FRA(inp, t);
            t = FRS(one, z); t = FRM(l0, t); acc = FRM(acc, y); acc = FRA(acc, t);
            t = FRM(z, z); t = FRS(t, z); t = FRM(ll, t); acc = FRM(acc, y); acc = FRA(acc, t);
            fe left = FRA(ap, beta); left = FRM(zn, left); u = FRA(sp, gamma); left = FRM(left, u);
            fe right = FRA(inp, beta); right = FRM(z, right); u = FRA(tbl, gamma); right = FRM(right, u);
            t = FRS(left, right); t = FRM(la_, t); acc = FRM(acc, y); acc = FRA(acc, t);
            fe d = FRS(ap, sp);
            t = FRM(l0, d); acc = FRM(acc, y); acc = FRA(acc, t);
            t = FRS(ap, apv); t = FRM(d, t); t = FRM(la_, t); acc = FRM(acc, y); acc = FRA(acc, t);
        }
        J->h[i] = FRM(acc, J->t_inv[i % step]);
        xpt = FRM(xpt, J->ext_omega);
    }
#undef FX
}

/* batch = batch * v + poly over rows [lo, hi) */
typedef struct { fe* batch; const fe* poly; fe v; } horner_job;
static void horner_fn(void* p, size_t lo, size_t hi) {
    horner_job* j = (horner_job*)p;
    for (size_t i = lo; i < hi; i++) { fe t = FRM(j->batch[i], j->v); j->batch[i] = FRA(t, j->poly[i]); }
}

/* create_proof for ONE instance.  advice: [NADV][n] Montgomery, rows >= u = n - BF - 1 are replaced by blinds.
 * seed32 / nonce / proof_index: the blinding stream.  proof: PROOF_BYTES.  challenges_out (optional): theta, beta,
 * gamma, y, x, v (Montgomery) for stage-by-stage debugging.  Returns 0, or -1 if a lookup input is not in the table. */
int orc_plonk_create_proof(const orc_plonk_key* K, const fe* advice, const uint8_t seed32[32], uint64_t nonce, uint32_t proof_index,
                           int nthreads, uint8_t* proof, fe* challenges_out) {
    const uint32_t k = K->k, ext_k = k + 2;
    if (k < 4 || k > 26) return -4;
    const size_t n = (size_t)1 << k, en = n * 4, u = n - (BF + 1), step = 4;
    blind_key BK;
    for (int i = 0; i < 8; i++) BK.key[i] = (uint32_t)seed32[4 * i] | ((uint32_t)seed32[4 * i + 1] << 8) | ((uint32_t)seed32[4 * i + 2] << 16) | ((uint32_t)seed32[4 * i + 3] << 24);
    BK.nonce[0] = (uint32_t)nonce; BK.nonce[1] = (uint32_t)(nonce >> 32);
    const fe omega = orc_omega(k), omega_inv = fe_inv(&FR, &omega), ext_omega = orc_omega(ext_k);
    static const uint64_t DELTA[4] = {0x870e56bbe533e9a2ull, 0x5b5f898e5e963f25ull, 0x64ec26aad4c86e71ull, 0x09226b6e22c6f0caull};
    static const uint64_t ZETA[4] = {0xb8ca0b2d36636f23ull, 0xcc37a73fec2bc5e9ull, 0x048b6e193fd84104ull, 0x30644e72e131a029ull};
    const fe delta = fe_to_mont(&FR, DELTA), zeta = fe_to_mont(&FR, ZETA);
    fe delta_pows[NPERM];
    delta_pows[0] = FR.one;
    for (int c = 1; c < NPERM; c++) delta_pows[c] = FRM(delta_pows[c - 1], delta);
    const fe* fx = K->fixed_values;
    transcript T;
    tr_init(&T, proof);
    tr_common_scalar(&T, &K->transcript_repr);
    g1a cm;

    /* advice columns: witness rows, blinding rows, commitments */
    fe* adv = fe_alloc(NADV * n);
    fe* adv_poly = fe_alloc(NADV * n);
    for (int c = 0; c < NADV; c++) {
        memcpy(adv + c * n, advice + c * n, u * sizeof(fe));
        for (size_t r = u; r < n; r++) adv[c * n + r] = blind_fe(&BK, proof_index, ST_ADVICE + c, (uint32_t)r);
        memcpy(adv_poly + c * n, adv + c * n, n * sizeof(fe));
        orc_lagrange_to_coeff(adv_poly + c * n, k, nthreads);
    }
    for (int c = 0; c < NADV; c++) { orc_best_multiexp(adv + c * n, K->g_lagrange, n, nthreads, &cm); tr_write_point(&T, &cm); }
    const fe theta = tr_squeeze(&T);

    /* lookups: compressed input / table, permuted pair, blinds, commitments */
    fe* table = fe_alloc(n);
    for (size_t i = 0; i < n; i++) { fe t = FRM(fx[F_T_TAG * n + i], theta); table[i] = FRA(t, fx[F_T_VALUE * n + i]); }
    fe* lkA = fe_alloc(NLOOK * n);   /* compressed inputs */
    fe* lk_ap = fe_alloc(NLOOK * n); /* A' */
    fe* lk_sp = fe_alloc(NLOOK * n); /* S' */
    int rc = 0;
    for (int l = 0; l < NLOOK && rc == 0; l++) {
        fe* A = lkA + l * n;
        for (size_t i = 0; i < n; i++) {
            fe t = FRM(fx[LK_FTAG[l] * n + i], theta), s = FRM(fx[LK_FSEL[l] * n + i], adv[LK_ACOL[l] * n + i]);
            A[i] = FRA(t, s);
        }
        if (permute_expression_pair(A, table, u, lk_ap + l * n, lk_sp + l * n)) { rc = -1; break; }
        for (size_t r = u; r < n; r++) {
            lk_ap[l * n + r] = blind_fe(&BK, proof_index, ST_LOOKUP_A + l, (uint32_t)r);
            lk_sp[l * n + r] = blind_fe(&BK, proof_index, ST_LOOKUP_S + l, (uint32_t)r);
        }
        orc_best_multiexp(lk_ap + l * n, K->g_lagrange, n, nthreads, &cm); tr_write_point(&T, &cm);
        orc_best_multiexp(lk_sp + l * n, K->g_lagrange, n, nthreads, &cm); tr_write_point(&T, &cm);
    }
    if (rc) { free(adv); free(adv_poly); free(table); free(lkA); free(lk_ap); free(lk_sp); return rc; }
    const fe beta = tr_squeeze(&T);
    const fe gamma = tr_squeeze(&T);

    /* permutation grand products, one per chunk of CHUNK columns; set s starts where set s-1 ended */
    fe* omega_pows = fe_alloc(n);
    { fe a = FR.one; for (size_t i = 0; i < n; i++) { omega_pows[i] = a; a = FRM(a, omega); } }
    fe* pz = fe_alloc(NSETS * n);
    {
        fe* den = fe_alloc(n);
        fe* num = fe_alloc(n);
        fe last_z = FR.one;
        for (int s = 0; s < NSETS; s++) {
            for (size_t i = 0; i < n; i++) den[i] = num[i] = FR.one;
            for (int c = s * CHUNK; c < (s + 1) * CHUNK && c < NPERM; c++) {
                for (size_t i = 0; i < n; i++) {
                    fe v = c < NADV ? adv[c * n + i] : fe_zero();
                    fe t = FRM(beta, K->sigma_values[(size_t)c * n + i]); t = FRA(t, gamma); t = FRA(t, v); den[i] = FRM(den[i], t);
                    fe w = FRM(delta_pows[c], omega_pows[i]); w = FRM(w, beta); w = FRA(w, gamma); w = FRA(w, v); num[i] = FRM(num[i], w);
                }
            }
            batch_invert(den, n);
            fe* z = pz + s * n;
            z[0] = last_z;
            for (size_t row = 1; row < n; row++) { fe t = FRM(z[row - 1], num[row - 1]); z[row] = FRM(t, den[row - 1]); }
            for (size_t r = n - BF; r < n; r++) z[r] = blind_fe(&BK, proof_index, ST_PERM_Z + s, (uint32_t)r);
            last_z = z[n - BF - 1];
        }
        free(den); free(num);
    }
    for (int s = 0; s < NSETS; s++) { orc_best_multiexp(pz + s * n, K->g_lagrange, n, nthreads, &cm); tr_write_point(&T, &cm); }
    /* lookup grand products */
    fe* lz = fe_alloc(NLOOK * n);
    {
        fe* den = fe_alloc(n);
        for (int l = 0; l < NLOOK; l++) {
            for (size_t i = 0; i < n; i++) { fe a = FRA(lk_ap[l * n + i], beta), b = FRA(lk_sp[l * n + i], gamma); den[i] = FRM(a, b); }
            batch_invert(den, n);
            fe* z = lz + l * n;
            z[0] = FR.one;
            for (size_t i = 0; i < u; i++) {
                fe a = FRA(lkA[l * n + i], beta), b = FRA(table[i], gamma);
                fe t = FRM(z[i], a); t = FRM(t, b); z[i + 1] = FRM(t, den[i]);
            }
            for (size_t r = n - BF; r < n; r++) z[r] = blind_fe(&BK, proof_index, ST_LOOKUP_Z + l, (uint32_t)r);
            orc_best_multiexp(z, K->g_lagrange, n, nthreads, &cm); tr_write_point(&T, &cm);
        }
        free(den);
    }
    /* vanishing argument: random polynomial (coefficient form, committed on g) */
    fe* random_poly = fe_alloc(n);
    for (size_t r = 0; r < n; r++) random_poly[r] = blind_fe(&BK, proof_index, ST_RANDOM_POLY, (uint32_t)r);
    orc_best_multiexp(random_poly, K->g, n, nthreads, &cm); tr_write_point(&T, &cm);
    const fe y = tr_squeeze(&T);

    /* coefficient forms and extended cosets */
    fe* pz_poly = fe_alloc(NSETS * n);
    fe* lz_poly = fe_alloc(NLOOK * n);
    fe* la_poly = fe_alloc(NLOOK * n);
    fe* ls_poly = fe_alloc(NLOOK * n);
    memcpy(pz_poly, pz, NSETS * n * sizeof(fe));
    memcpy(lz_poly, lz, NLOOK * n * sizeof(fe));
    memcpy(la_poly, lk_ap, NLOOK * n * sizeof(fe));
    memcpy(ls_poly, lk_sp, NLOOK * n * sizeof(fe));
    for (int s = 0; s < NSETS; s++) orc_lagrange_to_coeff(pz_poly + s * n, k, nthreads);
    for (int l = 0; l < NLOOK; l++) {
        orc_lagrange_to_coeff(lz_poly + l * n, k, nthreads);
        orc_lagrange_to_coeff(la_poly + l * n, k, nthreads);
        orc_lagrange_to_coeff(ls_poly + l * n, k, nthreads);
    }
    const int NEXT = NADV + NSETS + 3 * NLOOK;
    fe* ext = fe_alloc((size_t)NEXT * en);
    quot_job QJ;
    {
        int e = 0;
        for (int c = 0; c < NADV; c++, e++) { orc_coeff_to_extended(adv_poly + c * n, k, ext_k, ext + (size_t)e * en, nthreads); QJ.ae[c] = ext + (size_t)e * en; }
        for (int s = 0; s < NSETS; s++, e++) { orc_coeff_to_extended(pz_poly + s * n, k, ext_k, ext + (size_t)e * en, nthreads); QJ.pz[s] = ext + (size_t)e * en; }
        for (int l = 0; l < NLOOK; l++, e++) { orc_coeff_to_extended(lz_poly + l * n, k, ext_k, ext + (size_t)e * en, nthreads); QJ.lz[l] = ext + (size_t)e * en; }
        for (int l = 0; l < NLOOK; l++, e++) { orc_coeff_to_extended(la_poly + l * n, k, ext_k, ext + (size_t)e * en, nthreads); QJ.la[l] = ext + (size_t)e * en; }
        for (int l = 0; l < NLOOK; l++, e++) { orc_coeff_to_extended(ls_poly + l * n, k, ext_k, ext + (size_t)e * en, nthreads); QJ.ls[l] = ext + (size_t)e * en; }
    }
    fe* h = fe_alloc(en);
    QJ.key = K; QJ.n = n; QJ.ext_n = en; QJ.step = step;
    QJ.theta = theta; QJ.beta = beta; QJ.gamma = gamma; QJ.y = y;
    memcpy(QJ.delta_pows, delta_pows, sizeof delta_pows);
    {
        fe zn = fe_pow_u64(zeta, n), wn = fe_pow_u64(ext_omega, n), cur = FR.one;
        for (int i = 0; i < 4; i++) { fe t = FRM(zn, cur); t = FRS(t, FR.one); QJ.t_inv[i] = fe_inv(&FR, &t); cur = FRM(cur, wn); }
    }
    QJ.ext_omega = ext_omega; QJ.zeta = zeta; QJ.h = h;
    par_range(quot_fn, &QJ, en, nthreads);
    free(ext);
    orc_extended_to_coeff(h, ext_k, nthreads);
    for (int j = 0; j < QD; j++) { orc_best_multiexp(h + (size_t)j * n, K->g, n, nthreads, &cm); tr_write_point(&T, &cm); }
    const fe x = tr_squeeze(&T);
    const fe xn = fe_pow_u64(x, n);

    /* evaluations, in halo2's order */
    const fe x_next = FRM(x, omega), x_inv = FRM(x, omega_inv);
    fe x_last = x;
    for (int i = 0; i < BF + 1; i++) x_last = FRM(x_last, omega_inv);
    fe ev;
#define EVAL(poly, pt) do { ev = eval_poly_par((poly), n, &(pt), nthreads); tr_write_scalar(&T, &ev); } while (0)
    for (int c = 0; c < NADV; c++) EVAL(adv_poly + c * n, x);
    EVAL(adv_poly + 4 * n, x_next);
    for (int c = 0; c < NFIX; c++) EVAL(K->fixed_polys + (size_t)c * n, x);
    fe* h_poly = fe_alloc(n);
    for (size_t i = 0; i < n; i++) {
        fe a = fe_zero();
        for (int j = QD - 1; j >= 0; j--) { a = FRM(a, xn); a = FRA(a, h[(size_t)j * n + i]); }
        h_poly[i] = a;
    }
    EVAL(random_poly, x);
    for (int c = 0; c < NPERM; c++) EVAL(K->sigma_polys + (size_t)c * n, x);
    for (int s = 0; s < NSETS; s++) {
        EVAL(pz_poly + s * n, x);
        EVAL(pz_poly + s * n, x_next);
        if (s != NSETS - 1) EVAL(pz_poly + s * n, x_last);
    }
    for (int l = 0; l < NLOOK; l++) {
        EVAL(lz_poly + l * n, x);
        EVAL(lz_poly + l * n, x_next);
        EVAL(la_poly + l * n, x);
        EVAL(la_poly + l * n, x_inv);
        EVAL(ls_poly + l * n, x);
    }
#undef EVAL
    /* multiopen (GWC): queries in halo2's chain order, grouped by point in order of first appearance */
    enum { PX = 0, PNEXT = 1, PLAST = 2, PINV = 3, MAXQ = 80 };
    const fe pts[4] = {x, x_next, x_last, x_inv};
    int qpt[MAXQ];
    const fe* qpoly[MAXQ];
    int nq = 0;
#define ADDQ(pt, poly) do { qpt[nq] = (pt); qpoly[nq] = (poly); nq++; } while (0)
    for (int c = 0; c < NADV; c++) ADDQ(PX, adv_poly + c * n);
    ADDQ(PNEXT, adv_poly + 4 * n);
    for (int s = 0; s < NSETS; s++) { ADDQ(PX, pz_poly + s * n); ADDQ(PNEXT, pz_poly + s * n); }
    for (int s = NSETS - 2; s >= 0; s--) ADDQ(PLAST, pz_poly + s * n);
    for (int l = 0; l < NLOOK; l++) {
        ADDQ(PX, lz_poly + l * n); ADDQ(PX, la_poly + l * n); ADDQ(PX, ls_poly + l * n); ADDQ(PINV, la_poly + l * n); ADDQ(PNEXT, lz_poly + l * n);
    }
    for (int c = 0; c < NFIX; c++) ADDQ(PX, K->fixed_polys + (size_t)c * n);
    for (int c = 0; c < NPERM; c++) ADDQ(PX, K->sigma_polys + (size_t)c * n);
    ADDQ(PX, h_poly);
    ADDQ(PX, random_poly);
#undef ADDQ
    const fe v = tr_squeeze(&T);
    int order[4], norder = 0;
    for (int q = 0; q < nq; q++) {
        int seen = 0;
        for (int o = 0; o < norder; o++) seen |= order[o] == qpt[q];
        if (!seen) order[norder++] = qpt[q];
    }
    fe* batch = fe_alloc(n);
    fe* wq = fe_alloc(n);
    for (int o = 0; o < norder; o++) {
        memset(batch, 0, n * sizeof(fe));
        for (int q = 0; q < nq; q++) {
            if (qpt[q] != order[o]) continue;
            horner_job hj = {batch, qpoly[q], v};
            par_range(horner_fn, &hj, n, nthreads);
        }
        /* kate division by (X - z): q_{i-1} = c_i + z q_i */
        const fe z = pts[order[o]];
        fe acc = fe_zero();
        for (size_t i = n - 1; i >= 1; i--) { acc = FRM(acc, z); acc = FRA(acc, batch[i]); wq[i - 1] = acc; }
        wq[n - 1] = fe_zero();
        orc_best_multiexp(wq, K->g, n, nthreads, &cm); tr_write_point(&T, &cm);
    }
    if (challenges_out) { challenges_out[0] = theta; challenges_out[1] = beta; challenges_out[2] = gamma; challenges_out[3] = y; challenges_out[4] = x; challenges_out[5] = v; }
    free(batch); free(wq); free(h_poly); free(h); free(pz_poly); free(lz_poly); free(la_poly); free(ls_poly);
    free(random_poly); free(lz); free(pz); free(omega_pows); free(lkA); free(lk_ap); free(lk_sp); free(table); free(adv); free(adv_poly);
    return T.pos == PROOF_BYTES ? 0 : -3;
}

uint32_t orc_plonk_proof_bytes(void) { return PROOF_BYTES; }
