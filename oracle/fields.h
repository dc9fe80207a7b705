/* TEST INFRASTRUCTURE ONLY - CPU oracle / CPU baseline, never linked into the product.
 *
 * BN254 Fr, Fq and G1 in the style of halo2curves' portable (non-asm) backend: 4 x u64
 * little-endian limbs, Montgomery form R = 2^256, 64x64->128 multiply-accumulate ("mac")
 * with u128.  halo2curves is a third-party dependency of the reference and is not
 * vendored under /root/reference: this restates its published algorithm
 * (CIOS Montgomery multiplication; Jacobian coordinates for G1, a = 0, b = 3).
 * Deliberately a different formulation from the product's 8 x u32 even/odd-lane kernels.
 */
#ifndef ORACLE_FIELDS_H
#define ORACLE_FIELDS_H
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;               /* Montgomery form */
typedef struct { fe x, y; } g1a;                    /* affine, identity = (0,0) */
typedef struct { fe x, y, z; } g1j;                 /* Jacobian, identity z = 0 */

typedef struct {
    uint64_t mod[4];
    uint64_t inv;      /* -mod^-1 mod 2^64 */
    fe one;            /* R mod p */
    fe r2;             /* R^2 mod p */
} field_params;

static const field_params FR = {
    {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
    0xc2e1f593efffffffull,
    {{0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full}},
    {{0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull}},
};
static const field_params FQ = {
    {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
    0x87d20782e4866389ull,
    {{0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full}},
    {{0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full}},
};

static inline int fe_is_zero(const fe* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe* a, const fe* b) { return memcmp(a, b, sizeof(fe)) == 0; }
static inline fe fe_zero(void) { fe z = {{0, 0, 0, 0}}; return z; }

static inline int geq_mod(const uint64_t* a, const uint64_t* m) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > m[i]) return 1;
        if (a[i] < m[i]) return 0;
    }
    return 1;
}
static inline void sub_mod_raw(uint64_t* a, const uint64_t* m) {
    u128 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 t = (u128)a[i] - m[i] - (uint64_t)br;
        a[i] = (uint64_t)t;
        br = (t >> 64) & 1;
    }
}
static inline fe fe_add(const field_params* P, const fe* a, const fe* b) {
    fe r;
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)a->l[i] + b->l[i];
        r.l[i] = (uint64_t)c;
        c >>= 64;
    }
    if (geq_mod(r.l, P->mod)) sub_mod_raw(r.l, P->mod);
    return r;
}
static inline fe fe_sub(const field_params* P, const fe* a, const fe* b) {
    fe r;
    u128 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 t = (u128)a->l[i] - b->l[i] - (uint64_t)br;
        r.l[i] = (uint64_t)t;
        br = (t >> 64) & 1;
    }
    if (br) {
        u128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (u128)r.l[i] + P->mod[i];
            r.l[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    return r;
}
static inline fe fe_neg(const field_params* P, const fe* a) {
    fe z = fe_zero();
    return fe_sub(P, &z, a);
}
static inline fe fe_dbl(const field_params* P, const fe* a) { return fe_add(P, a, a); }

/* CIOS Montgomery multiplication */
static inline fe fe_mul(const field_params* P, const fe* a, const fe* b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a->l[j] * b->l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * P->inv;
        c = (u128)m * P->mod[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * P->mod[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    fe r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || geq_mod(r.l, P->mod)) sub_mod_raw(r.l, P->mod);
    return r;
}
static inline fe fe_sqr(const field_params* P, const fe* a) { return fe_mul(P, a, a); }
static inline fe fe_to_mont(const field_params* P, const uint64_t* canon) {
    fe c = {{canon[0], canon[1], canon[2], canon[3]}};
    return fe_mul(P, &c, &P->r2);
}
static inline void fe_from_mont(const field_params* P, const fe* a, uint64_t* canon) {
    fe one = {{1, 0, 0, 0}};
    fe r = fe_mul(P, a, &one);
    memcpy(canon, r.l, 32);
}
static inline fe fe_from_u64(const field_params* P, uint64_t v) {
    uint64_t c[4] = {v, 0, 0, 0};
    return fe_to_mont(P, c);
}
static inline fe fe_pow(const field_params* P, const fe* a, const uint64_t* e) {
    fe acc = P->one;
    for (int w = 3; w >= 0; w--)
        for (int b = 63; b >= 0; b--) {
            acc = fe_sqr(P, &acc);
            if ((e[w] >> b) & 1) acc = fe_mul(P, &acc, a);
        }
    return acc;
}
static inline fe fe_inv(const field_params* P, const fe* a) { /* a^(p-2); inv(0) = 0 */
    uint64_t e[4] = {P->mod[0] - 2, P->mod[1], P->mod[2], P->mod[3]};
    return fe_pow(P, a, e);
}

/* ---- G1, Jacobian (halo2curves' formulas are the standard a = 0 ones) ---------------- */
static inline g1j g1j_identity(void) {
    g1j r;
    r.x = fe_zero();
    r.y = FQ.one;
    r.z = fe_zero();
    return r;
}
static inline int g1j_is_identity(const g1j* p) { return fe_is_zero(&p->z); }
static inline int g1a_is_identity(const g1a* p) { return fe_is_zero(&p->x) && fe_is_zero(&p->y); }

static inline g1j g1j_double(const g1j* p) { /* dbl-2009-l */
    if (g1j_is_identity(p)) return *p;
    const field_params* F = &FQ;
    fe a = fe_sqr(F, &p->x), b = fe_sqr(F, &p->y), c = fe_sqr(F, &b);
    fe xb = fe_add(F, &p->x, &b);
    fe d = fe_sqr(F, &xb);
    d = fe_sub(F, &d, &a);
    d = fe_sub(F, &d, &c);
    d = fe_dbl(F, &d);
    fe e = fe_add(F, &a, &a);
    e = fe_add(F, &e, &a);
    fe f = fe_sqr(F, &e);
    g1j r;
    fe z3 = fe_mul(F, &p->z, &p->y);
    r.z = fe_dbl(F, &z3);
    fe d2 = fe_dbl(F, &d);
    r.x = fe_sub(F, &f, &d2);
    fe c8 = fe_dbl(F, &c);
    c8 = fe_dbl(F, &c8);
    c8 = fe_dbl(F, &c8);
    fe t = fe_sub(F, &d, &r.x);
    t = fe_mul(F, &e, &t);
    r.y = fe_sub(F, &t, &c8);
    return r;
}
static inline g1j g1j_add_affine(const g1j* p, const g1a* q) { /* madd-2007-bl */
    const field_params* F = &FQ;
    if (g1a_is_identity(q)) return *p;
    if (g1j_is_identity(p)) {
        g1j r;
        r.x = q->x;
        r.y = q->y;
        r.z = F->one;
        return r;
    }
    fe z1z1 = fe_sqr(F, &p->z);
    fe u2 = fe_mul(F, &q->x, &z1z1);
    fe s2 = fe_mul(F, &q->y, &p->z);
    s2 = fe_mul(F, &s2, &z1z1);
    if (fe_eq(&u2, &p->x)) {
        if (fe_eq(&s2, &p->y)) return g1j_double(p);
        return g1j_identity();
    }
    fe h = fe_sub(F, &u2, &p->x);
    fe hh = fe_sqr(F, &h);
    fe i = fe_dbl(F, &hh);
    i = fe_dbl(F, &i);
    fe j = fe_mul(F, &h, &i);
    fe rr = fe_sub(F, &s2, &p->y);
    rr = fe_dbl(F, &rr);
    fe v = fe_mul(F, &p->x, &i);
    g1j r;
    fe r2 = fe_sqr(F, &rr);
    r.x = fe_sub(F, &r2, &j);
    fe v2 = fe_dbl(F, &v);
    r.x = fe_sub(F, &r.x, &v2);
    fe t = fe_sub(F, &v, &r.x);
    t = fe_mul(F, &rr, &t);
    fe yj = fe_mul(F, &p->y, &j);
    yj = fe_dbl(F, &yj);
    r.y = fe_sub(F, &t, &yj);
    fe zh = fe_add(F, &p->z, &h);
    zh = fe_sqr(F, &zh);
    zh = fe_sub(F, &zh, &z1z1);
    r.z = fe_sub(F, &zh, &hh);
    return r;
}
static inline g1j g1j_add(const g1j* p, const g1j* q) { /* add-2007-bl */
    const field_params* F = &FQ;
    if (g1j_is_identity(p)) return *q;
    if (g1j_is_identity(q)) return *p;
    fe z1z1 = fe_sqr(F, &p->z), z2z2 = fe_sqr(F, &q->z);
    fe u1 = fe_mul(F, &p->x, &z2z2), u2 = fe_mul(F, &q->x, &z1z1);
    fe s1 = fe_mul(F, &p->y, &q->z);
    s1 = fe_mul(F, &s1, &z2z2);
    fe s2 = fe_mul(F, &q->y, &p->z);
    s2 = fe_mul(F, &s2, &z1z1);
    if (fe_eq(&u1, &u2)) {
        if (fe_eq(&s1, &s2)) return g1j_double(p);
        return g1j_identity();
    }
    fe h = fe_sub(F, &u2, &u1);
    fe i = fe_dbl(F, &h);
    i = fe_sqr(F, &i);
    fe j = fe_mul(F, &h, &i);
    fe rr = fe_sub(F, &s2, &s1);
    rr = fe_dbl(F, &rr);
    fe v = fe_mul(F, &u1, &i);
    g1j r;
    fe r2 = fe_sqr(F, &rr);
    r.x = fe_sub(F, &r2, &j);
    fe v2 = fe_dbl(F, &v);
    r.x = fe_sub(F, &r.x, &v2);
    fe t = fe_sub(F, &v, &r.x);
    t = fe_mul(F, &rr, &t);
    fe sj = fe_mul(F, &s1, &j);
    sj = fe_dbl(F, &sj);
    r.y = fe_sub(F, &t, &sj);
    fe zz = fe_add(F, &p->z, &q->z);
    zz = fe_sqr(F, &zz);
    zz = fe_sub(F, &zz, &z1z1);
    zz = fe_sub(F, &zz, &z2z2);
    r.z = fe_mul(F, &zz, &h);
    return r;
}
static inline g1a g1j_to_affine(const g1j* p) {
    const field_params* F = &FQ;
    g1a r;
    if (g1j_is_identity(p)) {
        r.x = fe_zero();
        r.y = fe_zero();
        return r;
    }
    fe zi = fe_inv(F, &p->z);
    fe zi2 = fe_sqr(F, &zi);
    fe zi3 = fe_mul(F, &zi2, &zi);
    r.x = fe_mul(F, &p->x, &zi2);
    r.y = fe_mul(F, &p->y, &zi3);
    return r;
}
#endif
