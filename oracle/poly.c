/* TEST INFRASTRUCTURE ONLY - CPU oracle and CPU baseline for hot path (b).
 *
 * Restates, in plain C with pthreads, the algorithms of the reference's third-party
 * prover crate (halo2_proofs, privacy-scaling-explorations/halo2, 2022-10 era; not
 * vendored under /root/reference, reached from reference benches/bench.rs:235-237 and
 * :321-329):
 *   arithmetic::best_fft        -> orc_best_fft      (bit-reverse, serial twiddle table,
 *                                                     radix-2 DIT, sub-FFTs over threads)
 *   arithmetic::best_multiexp   -> orc_best_multiexp (chunks over threads, each chunk
 *                                                     multiexp_serial: unsigned windows of
 *                                                     c = ceil(ln n) bits, 256/c + 1
 *                                                     segments, running-sum bucket fold)
 *   EvaluationDomain::{lagrange_to_coeff, coeff_to_extended, extended_to_coeff}
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this.  parity unpinned at this boundary (no reference test calls
 * create_proof): anchored on oracle/bn254.py's definitions instead.
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>

#include "fields.h"

/* ---- tiny fork-join helper -------------------------------------------------------------- */
typedef void (*job_fn)(void* arg, int tid, int nthreads);
typedef struct { job_fn fn; void* arg; int tid, n; } job_t;
static void* job_tramp(void* p) {
    job_t* j = (job_t*)p;
    j->fn(j->arg, j->tid, j->n);
    return 0;
}
static void run_parallel(job_fn fn, void* arg, int nthreads) {
    if (nthreads <= 1) { fn(arg, 0, 1); return; }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    job_t* jb = (job_t*)malloc(sizeof(job_t) * nthreads);
    for (int i = 0; i < nthreads; i++) {
        jb[i].fn = fn; jb[i].arg = arg; jb[i].tid = i; jb[i].n = nthreads;
        pthread_create(&th[i], 0, job_tramp, &jb[i]);
    }
    for (int i = 0; i < nthreads; i++) pthread_join(th[i], 0);
    free(th); free(jb);
}

/* ---- best_fft ---------------------------------------------------------------------------- */
static size_t bitreverse(size_t n, unsigned l) {
    size_t r = 0;
    for (unsigned i = 0; i < l; i++) { r = (r << 1) | (n & 1); n >>= 1; }
    return r;
}
typedef struct { fe* a; const fe* tw; size_t n; unsigned log_n; unsigned log_split; unsigned stage; } fft_job;

/* serial DIT stages [0, upto) on a contiguous block of `len` elements (len = 2^upto) */
static void fft_block(fe* a, size_t len, size_t n, const fe* tw) {
    size_t chunk = 2, twiddle_chunk = n / 2;
    while (chunk <= len) {
        for (size_t base = 0; base < len; base += chunk) {
            fe* left = a + base; fe* right = a + base + chunk / 2;
            for (size_t i = 0; i < chunk / 2; i++) {
                fe t = right[i];
                if (i) t = fe_mul(&FR, &t, &tw[i * twiddle_chunk]);
                right[i] = fe_sub(&FR, &left[i], &t);
                left[i] = fe_add(&FR, &left[i], &t);
            }
        }
        chunk *= 2; twiddle_chunk /= 2;
    }
}
static void fft_low_job(void* p, int tid, int nt) {
    fft_job* j = (fft_job*)p;
    size_t blocks = (size_t)1 << j->log_split, len = j->n >> j->log_split;
    for (size_t b = tid; b < blocks; b += nt) fft_block(j->a + b * len, len, j->n, j->tw);
}
static void fft_high_job(void* p, int tid, int nt) {
    fft_job* j = (fft_job*)p;
    size_t chunk = (size_t)1 << (j->stage + 1), half = chunk / 2, twc = j->n / chunk;
    size_t nb = j->n / 2;
    size_t lo = nb * tid / nt, hi = nb * (tid + 1) / nt;
    for (size_t bf = lo; bf < hi; bf++) {
        size_t blk = bf / half, i = bf % half;
        fe* l = j->a + blk * chunk + i; fe* r = l + half;
        fe t = *r;
        if (i) t = fe_mul(&FR, &t, &j->tw[i * twc]);
        *r = fe_sub(&FR, l, &t);
        *l = fe_add(&FR, l, &t);
    }
}
void orc_best_fft(fe* a, const fe* omega, unsigned log_n, int nthreads) {
    size_t n = (size_t)1 << log_n;
    for (size_t k = 0; k < n; k++) {
        size_t rk = bitreverse(k, log_n);
        if (k < rk) { fe t = a[rk]; a[rk] = a[k]; a[k] = t; }
    }
    if (log_n == 0) return;
    fe* tw = (fe*)malloc(sizeof(fe) * (n / 2 ? n / 2 : 1));
    fe w = FR.one;
    for (size_t i = 0; i < n / 2; i++) { tw[i] = w; w = fe_mul(&FR, &w, omega); }
    unsigned log_t = 0;
    while ((1 << (log_t + 1)) <= nthreads) log_t++;
    if (log_t > log_n) log_t = log_n;
    fft_job j = {a, tw, n, log_n, log_t, 0};
    run_parallel(fft_low_job, &j, 1 << log_t);            /* independent sub-FFTs */
    for (unsigned s = log_n - log_t; s < log_n; s++) {    /* joining stages */
        j.stage = s;
        run_parallel(fft_high_job, &j, nthreads);
    }
    free(tw);
}

/* omega_k = ROOT_OF_UNITY^(2^(28-k)) */
fe orc_omega(unsigned k) {
    static const uint64_t root[4] = {0xd34f1ed960c37c9cull, 0x3215cf6dd39329c8ull, 0x98865ea93dd31f74ull, 0x03ddb9f5166d18b7ull};
    fe w = fe_to_mont(&FR, root);
    for (unsigned i = k; i < 28; i++) w = fe_sqr(&FR, &w);
    return w;
}
static fe zeta(void) {
    static const uint64_t z[4] = {0xb8ca0b2d36636f23ull, 0xcc37a73fec2bc5e9ull, 0x048b6e193fd84104ull, 0x30644e72e131a029ull};
    return fe_to_mont(&FR, z);
}
typedef struct { fe* a; size_t n; fe m[3]; int use3; } scale_job;
static void scale_fn(void* p, int tid, int nt) {
    scale_job* j = (scale_job*)p;
    size_t lo = j->n * tid / nt, hi = j->n * (tid + 1) / nt;
    for (size_t i = lo; i < hi; i++) {
        int s = j->use3 ? (int)(i % 3) : 0;
        if (j->use3 && s == 0) continue;
        j->a[i] = fe_mul(&FR, &j->a[i], &j->m[s]);
    }
}
void orc_lagrange_to_coeff(fe* a, unsigned k, int nthreads) {
    fe w = orc_omega(k), wi = fe_inv(&FR, &w);
    orc_best_fft(a, &wi, k, nthreads);
    fe nn = fe_from_u64(&FR, (uint64_t)1 << k);
    scale_job j; j.a = a; j.n = (size_t)1 << k; j.use3 = 0; j.m[0] = fe_inv(&FR, &nn);
    run_parallel(scale_fn, &j, nthreads);
}
/* out has 2^ext_k elements; coeffs 2^k */
void orc_coeff_to_extended(const fe* coeffs, unsigned k, unsigned ext_k, fe* out, int nthreads) {
    size_t n = (size_t)1 << k, ne = (size_t)1 << ext_k;
    memcpy(out, coeffs, n * sizeof(fe));
    memset(out + n, 0, (ne - n) * sizeof(fe));
    scale_job j; j.a = out; j.n = n; j.use3 = 1; j.m[1] = zeta(); j.m[2] = fe_sqr(&FR, &j.m[1]); j.m[0] = FR.one;
    run_parallel(scale_fn, &j, nthreads);
    fe w = orc_omega(ext_k);
    orc_best_fft(out, &w, ext_k, nthreads);
}
void orc_extended_to_coeff(fe* a, unsigned ext_k, int nthreads) {
    fe w = orc_omega(ext_k), wi = fe_inv(&FR, &w);
    orc_best_fft(a, &wi, ext_k, nthreads);
    fe nn = fe_from_u64(&FR, (uint64_t)1 << ext_k), ni = fe_inv(&FR, &nn);
    scale_job j; j.a = a; j.n = (size_t)1 << ext_k; j.use3 = 0; j.m[0] = ni;
    run_parallel(scale_fn, &j, nthreads);
    fe z = zeta();
    j.use3 = 1; j.m[0] = FR.one; j.m[1] = fe_sqr(&FR, &z); j.m[2] = z;
    run_parallel(scale_fn, &j, nthreads);
}

/* ---- best_multiexp ------------------------------------------------------------------------ */
static size_t get_at(size_t segment, size_t c, const uint8_t* bytes) {
    size_t skip_bits = segment * c, skip_bytes = skip_bits / 8;
    if (skip_bytes >= 32) return 0;
    uint8_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (size_t i = 0; i < 8 && skip_bytes + i < 32; i++) v[i] = bytes[skip_bytes + i];
    uint64_t tmp;
    memcpy(&tmp, v, 8);
    tmp >>= skip_bits - skip_bytes * 8;
    tmp %= ((uint64_t)1 << c);
    return (size_t)tmp;
}
typedef struct { int kind; g1a a; g1j p; } bucket_t; /* 0 none, 1 affine, 2 projective */
static void multiexp_serial(const fe* coeffs, const g1a* bases, size_t len, g1j* acc) {
    uint8_t* repr = (uint8_t*)malloc(len * 32);
    for (size_t i = 0; i < len; i++) fe_from_mont(&FR, &coeffs[i], (uint64_t*)(repr + 32 * i));
    size_t c;
    if (len < 4) c = 1; else if (len < 32) c = 3; else c = (size_t)ceil(log((double)len));
    size_t segments = 256 / c + 1, nb = ((size_t)1 << c) - 1;
    bucket_t* buckets = (bucket_t*)malloc(sizeof(bucket_t) * nb);
    for (size_t seg = segments; seg-- > 0;) {
        for (size_t d = 0; d < c; d++) *acc = g1j_double(acc);
        for (size_t b = 0; b < nb; b++) buckets[b].kind = 0;
        for (size_t i = 0; i < len; i++) {
            size_t co = get_at(seg, c, repr + 32 * i);
            if (!co) continue;
            bucket_t* bk = &buckets[co - 1];
            if (bk->kind == 0) { bk->kind = 1; bk->a = bases[i]; }
            else if (bk->kind == 1) { g1j t; t.x = bk->a.x; t.y = bk->a.y; t.z = g1a_is_identity(&bk->a) ? fe_zero() : FQ.one;
                                      if (g1a_is_identity(&bk->a)) t = g1j_identity();
                                      bk->p = g1j_add_affine(&t, &bases[i]); bk->kind = 2; }
            else bk->p = g1j_add_affine(&bk->p, &bases[i]);
        }
        g1j running = g1j_identity();
        for (size_t b = nb; b-- > 0;) {
            if (buckets[b].kind == 1) running = g1j_add_affine(&running, &buckets[b].a);
            else if (buckets[b].kind == 2) running = g1j_add(&running, &buckets[b].p);
            *acc = g1j_add(acc, &running);
        }
    }
    free(buckets); free(repr);
}
typedef struct { const fe* coeffs; const g1a* bases; size_t n, chunk, nchunks; g1j* results; } msm_job;
static void msm_fn(void* p, int tid, int nt) {
    msm_job* j = (msm_job*)p;
    for (size_t ch = tid; ch < j->nchunks; ch += nt) {
        size_t lo = ch * j->chunk, hi = lo + j->chunk; if (hi > j->n) hi = j->n;
        j->results[ch] = g1j_identity();
        multiexp_serial(j->coeffs + lo, j->bases + lo, hi - lo, &j->results[ch]);
    }
}
/* result written normalised to affine (identity = (0,0)) */
void orc_best_multiexp(const fe* coeffs, const g1a* bases, size_t n, int nthreads, g1a* out) {
    g1j acc = g1j_identity();
    if (n > (size_t)nthreads && nthreads > 1) {
        size_t chunk = n / nthreads;
        size_t nchunks = (n + chunk - 1) / chunk;
        g1j* results = (g1j*)malloc(sizeof(g1j) * nchunks);
        msm_job j = {coeffs, bases, n, chunk, nchunks, results};
        run_parallel(msm_fn, &j, nthreads);
        for (size_t i = 0; i < nchunks; i++) acc = g1j_add(&acc, &results[i]);
        free(results);
    } else if (n > 0) {
        multiexp_serial(coeffs, bases, n, &acc);
    }
    *out = g1j_to_affine(&acc);
}

/* ---- helpers for tests / bench inputs -------------------------------------------------------- */
/* out[i] = (start + i + 1) * G, affine (bench config 5 bases) */
typedef struct { g1a* out; size_t n; } mult_job;
static g1j g1_mul_u64(const g1a* p, uint64_t k) {
    g1j acc = g1j_identity();
    for (int b = 63; b >= 0; b--) { acc = g1j_double(&acc); if ((k >> b) & 1) acc = g1j_add_affine(&acc, p); }
    return acc;
}
static void batch_normalize(const g1j* in, g1a* out, size_t n) {
    fe* pref = (fe*)malloc(sizeof(fe) * (n + 1));
    pref[0] = FQ.one;
    for (size_t i = 0; i < n; i++) pref[i + 1] = g1j_is_identity(&in[i]) ? pref[i] : fe_mul(&FQ, &pref[i], &in[i].z);
    fe inv = fe_inv(&FQ, &pref[n]);
    for (size_t i = n; i-- > 0;) {
        if (g1j_is_identity(&in[i])) { out[i].x = fe_zero(); out[i].y = fe_zero(); continue; }
        fe zi = fe_mul(&FQ, &inv, &pref[i]);
        inv = fe_mul(&FQ, &inv, &in[i].z);
        fe zi2 = fe_sqr(&FQ, &zi), zi3 = fe_mul(&FQ, &zi2, &zi);
        out[i].x = fe_mul(&FQ, &in[i].x, &zi2);
        out[i].y = fe_mul(&FQ, &in[i].y, &zi3);
    }
    free(pref);
}
static void mult_fn(void* p, int tid, int nt) {
    mult_job* j = (mult_job*)p;
    size_t lo = j->n * tid / nt, hi = j->n * (tid + 1) / nt;
    if (lo >= hi) return;
    g1a g; g.x = fe_from_u64(&FQ, 1); g.y = fe_from_u64(&FQ, 2);
    g1j* tmp = (g1j*)malloc(sizeof(g1j) * (hi - lo));
    g1j acc = g1_mul_u64(&g, lo + 1);
    for (size_t i = lo; i < hi; i++) { tmp[i - lo] = acc; acc = g1j_add_affine(&acc, &g); }
    batch_normalize(tmp, j->out + lo, hi - lo);
    free(tmp);
}
void orc_g1_multiples(g1a* out, size_t n, int nthreads) {
    mult_job j = {out, n};
    run_parallel(mult_fn, &j, nthreads);
}
/* out[i] = scalars[i] * G (KZG SRS from a seeded secret: scalars = powers of s or L_i(s)) */
typedef struct { const fe* sc; g1a* out; size_t n; } smul_job;
static void smul_fn(void* p, int tid, int nt) {
    smul_job* j = (smul_job*)p;
    size_t lo = j->n * tid / nt, hi = j->n * (tid + 1) / nt;
    if (lo >= hi) return;
    g1a g; g.x = fe_from_u64(&FQ, 1); g.y = fe_from_u64(&FQ, 2);
    /* fixed-base: table of 2^(8w) * d * G would be faster; plain double-and-add is fine for tests */
    g1j* tmp = (g1j*)malloc(sizeof(g1j) * (hi - lo));
    for (size_t i = lo; i < hi; i++) {
        uint64_t k[4]; fe_from_mont(&FR, &j->sc[i], k);
        g1j acc = g1j_identity();
        for (int w = 3; w >= 0; w--) for (int b = 63; b >= 0; b--) { acc = g1j_double(&acc); if ((k[w] >> b) & 1) acc = g1j_add_affine(&acc, &g); }
        tmp[i - lo] = acc;
    }
    batch_normalize(tmp, j->out + lo, hi - lo);
    free(tmp);
}
void orc_g1_scalar_muls(const fe* scalars, g1a* out, size_t n, int nthreads) {
    smul_job j = {scalars, out, n};
    run_parallel(smul_fn, &j, nthreads);
}
void orc_fr_sub_array(fe* a, const fe* b, size_t n) { for (size_t i = 0; i < n; i++) a[i] = fe_sub(&FR, &a[i], &b[i]); }
void orc_fr_mul_array(fe* a, const fe* b, size_t n) { for (size_t i = 0; i < n; i++) a[i] = fe_mul(&FR, &a[i], &b[i]); }

/* bulk conversions for the Python oracle prover: canonical <-> Montgomery, in place */
void orc_fr_to_mont_array(fe* a, size_t n) { for (size_t i = 0; i < n; i++) { uint64_t c[4]; memcpy(c, a[i].l, 32); a[i] = fe_to_mont(&FR, c); } }
void orc_fr_from_mont_array(fe* a, size_t n) { for (size_t i = 0; i < n; i++) { uint64_t c[4]; fe_from_mont(&FR, &a[i], c); memcpy(a[i].l, c, 32); } }
void orc_fq_to_mont_array(fe* a, size_t n) { for (size_t i = 0; i < n; i++) { uint64_t c[4]; memcpy(c, a[i].l, 32); a[i] = fe_to_mont(&FQ, c); } }
void orc_fq_from_mont_array(fe* a, size_t n) { for (size_t i = 0; i < n; i++) { uint64_t c[4]; fe_from_mont(&FQ, &a[i], c); memcpy(a[i].l, c, 32); } }
