/* TEST INFRASTRUCTURE ONLY - CPU oracle and CPU baseline for hot path (a): witness synthesis.
 *
 * A direct, sequential restatement of what the reference executes inside
 * Circuit::synthesize, one row at a time on one thread, exactly like the Rust code:
 *   - BigIntChip            /root/reference/src/big_integer/chip.rs (each function cites its lines)
 *   - RSAChip               /root/reference/src/chip.rs:58-199
 *   - bench circuit driver  /root/reference/benches/bench.rs:132-225 (SHA-disabled branch)
 *   - MainGate / RangeChip  third-party `maingate` crate (halo2wrong rev 63bde545, reference
 *                           Cargo.toml:13; NOT vendored under /root/reference): restated from
 *                           its published semantics - one 5-column row per arithmetic op,
 *                           gate  a*sa + b*sb + c*sc + d*sd + e*se + a*b*s_mul_ab + c*d*s_mul_cd
 *                                 + e_next*se_next + s_constant = 0,
 *                           RangeChip::assign = decomposition into sublimbs, 4 per row in a..d
 *                           with the running remainder in e, lookups on a..d.
 * Value-level parity is pinned by the reference's known-answer tests (src/chip.rs:703-803,
 * src/big_integer/chip.rs:2797-3264); cell placement follows maingate and is "layout
 * unpinned" (SURVEY.md R1).  The table this file builds is checked by orc_check(), a
 * MockProver-style verifier (gate on every row, range lookups, copy constraints).
 */
#include <stdio.h>
#include <stdlib.h>

#include "fields.h"

#define NADV 5
enum { F_SA, F_SB, F_SC, F_SD, F_SE, F_MUL_AB, F_MUL_CD, F_SE_NEXT, F_CONST, NFIX };

typedef struct { int col; uint32_t row; fe v; } aval; /* AssignedValue: cell + witness value */

typedef struct {
    unsigned k;
    size_t nrows;
    fe* adv[NADV];
    fe* fix[NFIX];
    uint8_t *s_comp, *tag_comp, *s_over, *tag_over; /* RangeChip selectors / tags per row */
    uint32_t (*copies)[4];
    size_t ncopies, capcopies;
    size_t offset;      /* RegionCtx offset (absolute row: regions are stacked by SimpleFloorPlanner) */
    int failed;         /* the reference would panic / return Err here */
    int overflow_rows;  /* ran past 2^k rows */
    int tag_of_bits[80];
    /* big-integer scratch */
} rctx;

/* ---- small unsigned big integers (u32 limbs) for the BigUint parts ---------------------- */
#define BIG_MAX 520 /* u32 limbs: enough for 2 * 4096 bits + slack */
typedef struct { uint32_t w[BIG_MAX]; int n; } big;
static void big_norm(big* a) { while (a->n > 0 && a->w[a->n - 1] == 0) a->n--; }
static void big_zero(big* a) { a->n = 0; }
static int big_cmp(const big* a, const big* b) {
    if (a->n != b->n) return a->n < b->n ? -1 : 1;
    for (int i = a->n - 1; i >= 0; i--) if (a->w[i] != b->w[i]) return a->w[i] < b->w[i] ? -1 : 1;
    return 0;
}
static void big_add_shifted_words(big* acc, const uint32_t* v, int nv, int shift_words) {
    uint64_t c = 0;
    int i = 0;
    while (acc->n < shift_words + nv + 1) acc->w[acc->n++] = 0;
    for (; i < nv; i++) { c += (uint64_t)acc->w[shift_words + i] + v[i]; acc->w[shift_words + i] = (uint32_t)c; c >>= 32; }
    for (int j = shift_words + nv; c && j < acc->n; j++) { c += acc->w[j]; acc->w[j] = (uint32_t)c; c >>= 32; }
    big_norm(acc);
}
static void big_mul(big* r, const big* a, const big* b) {
    for (int i = 0; i < a->n + b->n; i++) r->w[i] = 0;
    r->n = a->n + b->n;
    for (int i = 0; i < a->n; i++) {
        uint64_t c = 0;
        for (int j = 0; j < b->n; j++) { c += (uint64_t)a->w[i] * b->w[j] + r->w[i + j]; r->w[i + j] = (uint32_t)c; c >>= 32; }
        r->w[i + b->n] = (uint32_t)c;
    }
    big_norm(r);
}
/* r = a - b, returns 1 on underflow */
static int big_sub(big* r, const big* a, const big* b) {
    if (big_cmp(a, b) < 0) return 1;
    int64_t br = 0;
    for (int i = 0; i < a->n; i++) {
        int64_t t = (int64_t)a->w[i] - (i < b->n ? b->w[i] : 0) - br;
        r->w[i] = (uint32_t)t; br = t < 0;
    }
    r->n = a->n; big_norm(r);
    return 0;
}
/* Knuth algorithm D.  q = a / b, rem = a % b.  returns 1 if b == 0 */
static int big_divrem(big* q, big* rem, const big* a, const big* b) {
    if (b->n == 0) return 1;
    if (big_cmp(a, b) < 0) { big_zero(q); *rem = *a; return 0; }
    if (b->n == 1) {
        uint64_t r = 0; q->n = a->n;
        for (int i = a->n - 1; i >= 0; i--) { uint64_t cur = (r << 32) | a->w[i]; q->w[i] = (uint32_t)(cur / b->w[0]); r = cur % b->w[0]; }
        big_norm(q); rem->n = 1; rem->w[0] = (uint32_t)r; big_norm(rem); return 0;
    }
    int s = __builtin_clz(b->w[b->n - 1]);
    static __thread uint32_t un[BIG_MAX + 2], vn[BIG_MAX];
    int n = b->n, m = a->n - b->n;
    for (int i = n - 1; i > 0; i--) vn[i] = (b->w[i] << s) | (s ? b->w[i - 1] >> (32 - s) : 0);
    vn[0] = b->w[0] << s;
    un[a->n] = s ? a->w[a->n - 1] >> (32 - s) : 0;
    for (int i = a->n - 1; i > 0; i--) un[i] = (a->w[i] << s) | (s ? a->w[i - 1] >> (32 - s) : 0);
    un[0] = a->w[0] << s;
    q->n = m + 1;
    for (int j = m; j >= 0; j--) {
        uint64_t num = ((uint64_t)un[j + n] << 32) | un[j + n - 1];
        uint64_t qhat = num / vn[n - 1], rhat = num % vn[n - 1];
        while (qhat >= ((uint64_t)1 << 32) || qhat * vn[n - 2] > ((rhat << 32) | un[j + n - 2])) {
            qhat--; rhat += vn[n - 1];
            if (rhat >= ((uint64_t)1 << 32)) break;
        }
        int64_t borrow = 0; uint64_t carry = 0;
        for (int i = 0; i < n; i++) {
            uint64_t p = qhat * vn[i] + carry; carry = p >> 32;
            int64_t t = (int64_t)un[i + j] - borrow - (int64_t)(p & 0xffffffffu);
            un[i + j] = (uint32_t)t; borrow = t < 0;
        }
        int64_t t = (int64_t)un[j + n] - borrow - (int64_t)carry;
        un[j + n] = (uint32_t)t;
        if (t < 0) {
            qhat--; uint64_t c = 0;
            for (int i = 0; i < n; i++) { c += (uint64_t)un[i + j] + vn[i]; un[i + j] = (uint32_t)c; c >>= 32; }
            un[j + n] += (uint32_t)c;
        }
        q->w[j] = (uint32_t)qhat;
    }
    big_norm(q);
    rem->n = n;
    for (int i = 0; i < n; i++) rem->w[i] = (un[i] >> s) | (s && i + 1 <= n ? (uint32_t)(((uint64_t)un[i + 1] << (32 - s))) : 0);
    big_norm(rem);
    return 0;
}
static uint64_t big_limb64(const big* a, int i) {
    uint64_t lo = 2 * i < a->n ? a->w[2 * i] : 0, hi = 2 * i + 1 < a->n ? a->w[2 * i + 1] : 0;
    return lo | (hi << 32);
}
static int big_bits(const big* a) { return a->n == 0 ? 0 : 32 * (a->n - 1) + (32 - __builtin_clz(a->w[a->n - 1])); }

/* fe_to_big / big_to_fe (maingate helpers) */
static void fe_canon(const fe* a, uint64_t* c) { fe_from_mont(&FR, a, c); }
static fe fe_of_u64(uint64_t v) { return fe_from_u64(&FR, v); }
static fe fe_of_canon(const uint64_t* c) { return fe_to_mont(&FR, c); }

/* ---- region / main gate ------------------------------------------------------------------- */
static void add_copy(rctx* c, int ca, uint32_t ra, int cb, uint32_t rb) {
    if (c->ncopies == c->capcopies) {
        c->capcopies = c->capcopies ? 2 * c->capcopies : 1 << 16;
        c->copies = realloc(c->copies, c->capcopies * sizeof(*c->copies));
    }
    uint32_t* e = c->copies[c->ncopies++];
    e[0] = ca; e[1] = ra; e[2] = cb; e[3] = rb;
}
typedef struct { int kind; aval src; fe val; fe base; } term; /* kind: 0 Zero, 1 Assigned, 2 Unassigned */
static term t_zero(void) { term t; memset(&t, 0, sizeof t); return t; }
static fe FE_ONE(void) { return FR.one; }
static fe FE_MINUS_ONE(void) { return fe_neg(&FR, &FR.one); }
static term t_assigned(const aval* a, fe base) { term t; t.kind = 1; t.src = *a; t.val = a->v; t.base = base; return t; }
static term t_unassigned(fe v, fe base) { term t; memset(&t, 0, sizeof t); t.kind = 2; t.val = v; t.base = base; return t; }
#define T_MUL(a) t_assigned(a, fe_zero())      /* Term::assigned_to_mul */
#define T_ADD(a) t_assigned(a, FE_ONE())       /* Term::assigned_to_add */
#define T_SUB(a) t_assigned(a, FE_MINUS_ONE()) /* Term::assigned_to_sub */
#define U_MUL(v) t_unassigned(v, fe_zero())
#define U_ADD(v) t_unassigned(v, FE_ONE())
#define U_SUB(v) t_unassigned(v, FE_MINUS_ONE())

/* MainGate::apply: one row; terms go to columns a..e in order */
static void mg_apply(rctx* c, const term* terms, int nterms, fe constant, fe s_mul_ab, fe s_mul_cd, fe se_next, aval* out) {
    size_t row = c->offset;
    if (row >= c->nrows) { c->overflow_rows = 1; c->failed = 1; row = c->nrows - 1; }
    for (int i = 0; i < NADV; i++) {
        term t = i < nterms ? terms[i] : t_zero();
        c->adv[i][row] = t.val;
        c->fix[F_SA + i][row] = t.base;
        if (out) { out[i].col = i; out[i].row = (uint32_t)row; out[i].v = t.val; }
        if (t.kind == 1) add_copy(c, t.src.col, t.src.row, i, (uint32_t)row);
    }
    c->fix[F_MUL_AB][row] = s_mul_ab;
    c->fix[F_MUL_CD][row] = s_mul_cd;
    c->fix[F_SE_NEXT][row] = se_next;
    c->fix[F_CONST][row] = constant;
    c->offset++;
}
static void apply_add(rctx* c, const term* t, int n, fe constant, aval* out) { mg_apply(c, t, n, constant, fe_zero(), fe_zero(), fe_zero(), out); }
static void apply_mul(rctx* c, const term* t, int n, fe constant, aval* out) { mg_apply(c, t, n, constant, FE_ONE(), fe_zero(), fe_zero(), out); }

static aval mg_assign_constant(rctx* c, fe k) { term t[1] = {U_SUB(k)}; aval o[NADV]; apply_add(c, t, 1, k, o); return o[0]; }
static aval mg_assign_value(rctx* c, fe v) { term t[1] = {U_MUL(v)}; aval o[NADV]; apply_add(c, t, 1, fe_zero(), o); return o[0]; }
static aval mg_assign_bit(rctx* c, fe b) {
    term t[3] = {U_MUL(b), U_MUL(b), U_SUB(b)}; aval o[NADV]; apply_mul(c, t, 3, fe_zero(), o);
    add_copy(c, 0, o[0].row, 1, o[1].row); add_copy(c, 1, o[1].row, 2, o[2].row);
    return o[2];
}
static aval mg_add_with_constant(rctx* c, const aval* a, const aval* b, fe k) {
    fe s = fe_add(&FR, &a->v, &b->v); s = fe_add(&FR, &s, &k);
    term t[3] = {T_ADD(a), T_ADD(b), U_SUB(s)}; aval o[NADV]; apply_add(c, t, 3, k, o); return o[2];
}
static aval mg_add(rctx* c, const aval* a, const aval* b) { return mg_add_with_constant(c, a, b, fe_zero()); }
static aval mg_add_constant(rctx* c, const aval* a, fe k) {
    fe s = fe_add(&FR, &a->v, &k);
    term t[2] = {T_ADD(a), U_SUB(s)}; aval o[NADV]; apply_add(c, t, 2, k, o); return o[1];
}
static aval mg_sub(rctx* c, const aval* a, const aval* b) {
    fe s = fe_sub(&FR, &a->v, &b->v);
    term t[3] = {T_ADD(a), T_SUB(b), U_SUB(s)}; aval o[NADV]; apply_add(c, t, 3, fe_zero(), o); return o[2];
}
static aval mg_mul(rctx* c, const aval* a, const aval* b) {
    fe p = fe_mul(&FR, &a->v, &b->v);
    term t[3] = {T_MUL(a), T_MUL(b), U_SUB(p)}; aval o[NADV]; apply_mul(c, t, 3, fe_zero(), o); return o[2];
}
static aval mg_mul_add(rctx* c, const aval* a, const aval* b, const aval* to_add) {
    fe p = fe_mul(&FR, &a->v, &b->v); p = fe_add(&FR, &p, &to_add->v);
    term t[4] = {T_MUL(a), T_MUL(b), T_ADD(to_add), U_SUB(p)}; aval o[NADV]; apply_mul(c, t, 4, fe_zero(), o); return o[3];
}
static aval mg_and(rctx* c, const aval* a, const aval* b) { return mg_mul(c, a, b); }
static aval mg_not(rctx* c, const aval* a) {
    fe n = fe_sub(&FR, &FR.one, &a->v);
    term t[2] = {T_ADD(a), U_ADD(n)}; aval o[NADV]; apply_add(c, t, 2, FE_MINUS_ONE(), o); return o[1];
}
static aval mg_select(rctx* c, const aval* a, const aval* b, const aval* cond) {
    fe res = fe_eq(&cond->v, &FR.one) ? a->v : b->v;
    term t[5] = {T_MUL(cond), T_MUL(a), T_MUL(cond), T_ADD(b), U_SUB(res)}; aval o[NADV];
    mg_apply(c, t, 5, fe_zero(), FE_ONE(), FE_MINUS_ONE(), fe_zero(), o);
    add_copy(c, 0, o[0].row, 2, o[2].row);
    return o[4];
}
/* MainGate::is_zero via invert: r bit row, (a * a') - 1 + r = 0, r * a' - r = 0 */
static aval mg_is_zero(rctx* c, const aval* a) {
    fe r, ainv;
    if (fe_is_zero(&a->v)) { r = FR.one; ainv = FR.one; } else { r = fe_zero(); ainv = fe_inv(&FR, &a->v); }
    aval rb = mg_assign_bit(c, r);
    term t1[3] = {T_MUL(a), U_MUL(ainv), T_ADD(&rb)}; aval o[NADV]; apply_mul(c, t1, 3, FE_MINUS_ONE(), o);
    aval ai = o[1];
    term t2[3] = {T_MUL(&rb), T_MUL(&ai), T_SUB(&rb)}; apply_mul(c, t2, 3, fe_zero(), o);
    return rb;
}
static aval mg_is_equal(rctx* c, const aval* a, const aval* b) { aval d = mg_sub(c, a, b); return mg_is_zero(c, &d); }
static void mg_assert_equal(rctx* c, const aval* a, const aval* b) { add_copy(c, a->col, a->row, b->col, b->row); }
static void mg_assert_const(rctx* c, const aval* a, fe k) { term t[1] = {T_ADD(a)}; apply_add(c, t, 1, fe_neg(&FR, &k), NULL); }
static void mg_assert_zero(rctx* c, const aval* a) { mg_assert_const(c, a, fe_zero()); }
static void mg_assert_one(rctx* c, const aval* a) { mg_assert_const(c, a, FR.one); }

/* RangeChip::assign(ctx, v, limb_bit_len, bit_len) -> MainGate::decompose with lookups */
static aval range_assign(rctx* c, fe v, int limb_bits, int bit_len) {
    int nl = bit_len / limb_bits, over = bit_len % limb_bits;
    if (over) nl++;
    uint64_t cv[4]; fe_canon(&v, cv);
    big val; val.n = 8; for (int i = 0; i < 4; i++) { val.w[2 * i] = (uint32_t)cv[i]; val.w[2 * i + 1] = (uint32_t)(cv[i] >> 32); }
    fe remaining = v;
    aval result; memset(&result, 0, sizeof result);
    int nchunks = (nl - 1) / 4 + 1;
    for (int ch = 0; ch < nchunks; ch++) {
        term t[5]; int cnt = 0;
        fe composed = fe_zero();
        for (int j = 4 * ch; j < 4 * ch + 4 && j < nl; j++) {
            int bit0 = j * limb_bits;
            uint64_t sub = 0;
            for (int b = 0; b < limb_bits; b++) { int bit = bit0 + b; if (bit < 256 && ((val.w[bit >> 5] >> (bit & 31)) & 1)) sub |= (uint64_t)1 << b; }
            uint64_t bc[4] = {0, 0, 0, 0}; bc[bit0 >> 6] = (uint64_t)1 << (bit0 & 63);
            fe base = fe_of_canon(bc), sv = fe_of_u64(sub);
            t[cnt++] = t_unassigned(sv, base);
            fe pr = fe_mul(&FR, &sv, &base); composed = fe_add(&FR, &composed, &pr);
        }
        while (cnt < 4) t[cnt++] = t_zero();
        t[4] = U_SUB(remaining);
        int is_final = ch == nchunks - 1;
        size_t row = c->offset;
        if (row < c->nrows) {
            c->s_comp[row] = 1; c->tag_comp[row] = (uint8_t)c->tag_of_bits[limb_bits];
            if (is_final && over) { c->s_over[row] = 1; c->tag_over[row] = (uint8_t)c->tag_of_bits[over]; }
        }
        aval o[NADV];
        mg_apply(c, t, 5, fe_zero(), fe_zero(), fe_zero(), is_final ? fe_zero() : FE_ONE(), o);
        if (ch == 0) result = o[4];
        remaining = fe_sub(&FR, &remaining, &composed);
    }
    /* a value that does not fit leaves the last row's gate unsatisfied (caught by orc_check) */
    return result;
}

/* ---- BigIntChip ------------------------------------------------------------------------------ */
#define MAXL 140 /* max limbs of any integer here (2*64+...) */
typedef struct { aval l[MAXL]; int n; } bint; /* AssignedInteger */
typedef struct { int limb_width, num_limbs; } bigchip;

static int sublimb_bit_len(int bits) { int v = bits / 8; return v == 0 ? 1 : v; } /* chip.rs:1357-1365 */
static void u256_word_max(int limb_width, int min_n, big* out) { /* chip.rs:1368-1372 */
    big base; big_zero(&base); uint32_t one = 1; big_add_shifted_words(&base, &one, 1, 0);
    big m; m.n = limb_width / 32 + 1; for (int i = 0; i < m.n; i++) m.w[i] = 0; m.w[limb_width / 32] = 1u << (limb_width % 32);
    big mm1; big_sub(&mm1, &m, &base);
    big sq; big_mul(&sq, &mm1, &mm1);
    big nn; nn.n = 1; nn.w[0] = (uint32_t)min_n;
    big_mul(out, &nn, &sq);
    big_add_shifted_words(out, mm1.w, mm1.n, 0);
}
static fe fe_of_big(const big* b) { uint64_t c[4] = {big_limb64(b, 0), big_limb64(b, 1), big_limb64(b, 2), big_limb64(b, 3)}; return fe_of_canon(c); }
static fe limb_max_fe(int limb_width) { uint64_t c[4] = {0, 0, 0, 0}; c[limb_width >> 6] = (uint64_t)1 << (limb_width & 63); return fe_of_canon(c); }

/* AssignedInteger::to_big_uint (mod.rs:348-359) */
static void to_big_uint(const bint* a, int width, big* out) {
    big_zero(out);
    for (int i = 0; i < a->n; i++) {
        uint64_t cv[4]; fe_canon(&a->l[i].v, cv);
        uint32_t w[10] = {0};
        int sh = (width * i) % 32, wsh = (width * i) / 32;
        /* shift the 256-bit value left by sh bits into w */
        uint32_t v32[8]; for (int j = 0; j < 4; j++) { v32[2 * j] = (uint32_t)cv[j]; v32[2 * j + 1] = (uint32_t)(cv[j] >> 32); }
        for (int j = 0; j < 8; j++) { w[j] |= v32[j] << sh; if (sh) w[j + 1] |= v32[j] >> (32 - sh); }
        int nw = 9; while (nw > 0 && w[nw - 1] == 0) nw--;
        if (nw) big_add_shifted_words(out, w, nw, wsh);
    }
}
/* limb i (width bits) of a big as a field element */
static fe big_limb_fe(const big* b, int i, int width) {
    uint64_t v = 0;
    for (int bit = 0; bit < width; bit++) { int p = i * width + bit; if ((p >> 5) < b->n && ((b->w[p >> 5] >> (p & 31)) & 1)) v |= (uint64_t)1 << bit; }
    return fe_of_u64(v);
}

/* chip.rs:62-82 */
static void bi_assign_integer(rctx* c, const bigchip* ch, const fe* limbs, int n, bint* out) {
    out->n = n;
    for (int i = 0; i < n; i++) out->l[i] = range_assign(c, limbs[i], sublimb_bit_len(ch->limb_width), ch->limb_width);
}
/* chip.rs:1252-1281 */
static void bi_assign_constant(rctx* c, const bigchip* ch, const big* integer, int max_limbs, bint* out) {
    int bits = big_bits(integer), lw = ch->limb_width;
    int nl = bits % lw == 0 ? bits / lw : bits / lw + 1;
    if (nl > max_limbs) { c->failed = 1; nl = max_limbs; }
    out->n = max_limbs;
    for (int i = 0; i < nl; i++) out->l[i] = mg_assign_constant(c, big_limb_fe(integer, i, lw));
    aval zero = mg_assign_constant(c, fe_zero());
    for (int i = nl; i < max_limbs; i++) out->l[i] = zero;
}
/* chip.rs:130-147 max_value: num_limbs constants 2^limb_width - 1 */
static void bi_max_value(rctx* c, const bigchip* ch, int num_limbs, bint* out) {
    uint64_t m[4] = {0, 0, 0, 0};
    for (int b = 0; b < ch->limb_width; b++) m[b >> 6] |= (uint64_t)1 << (b & 63);
    fe limb_max = fe_of_canon(m);
    out->n = num_limbs;
    for (int i = 0; i < num_limbs; i++) out->l[i] = mg_assign_constant(c, limb_max);
}
static void bi_is_equal_fresh(rctx* c, const bint* a, const bint* b, aval* out);
/* chip.rs:245-297 */
static void bi_add(rctx* c, const bigchip* ch, const bint* a0, const bint* b0, bint* out) {
    int lw = ch->limb_width, n1 = a0->n, n2 = b0->n, max_n = n1 < n2 ? n2 : n1;
    aval zero = mg_assign_constant(c, fe_zero());
    bint a = *a0, b = *b0;
    for (int i = n1; i < max_n; i++) a.l[i] = zero;
    for (int i = n2; i < max_n; i++) b.l[i] = zero;
    aval carry = zero;
    aval limb_max = mg_assign_constant(c, limb_max_fe(lw));
    out->n = max_n + 1;
    for (int i = 0; i < max_n; i++) {
        aval a_b = mg_add(c, &a.l[i], &b.l[i]);
        aval sum = mg_add(c, &a_b, &carry);
        uint64_t sv[4]; fe_canon(&sum.v, sv);
        /* sum % 2^lw, sum >> lw on the canonical integer */
        big sb; sb.n = 8; for (int j = 0; j < 4; j++) { sb.w[2 * j] = (uint32_t)sv[j]; sb.w[2 * j + 1] = (uint32_t)(sv[j] >> 32); } big_norm(&sb);
        fe c_val = big_limb_fe(&sb, 0, lw);
        uint64_t hi[4] = {0, 0, 0, 0};
        for (int bit = lw; bit < 256; bit++) if ((sv[bit >> 6] >> (bit & 63)) & 1) hi[(bit - lw) >> 6] |= (uint64_t)1 << ((bit - lw) & 63);
        fe carry_f = fe_of_canon(hi);
        aval cc = range_assign(c, c_val, sublimb_bit_len(lw), lw);
        aval cy = range_assign(c, carry_f, sublimb_bit_len(lw), lw);
        aval c_add_carry = mg_mul_add(c, &cy, &limb_max, &cc);
        mg_assert_equal(c, &sum, &c_add_carry);
        out->l[i] = cc;
        carry = cy;
    }
    out->l[max_n] = carry;
}
/* chip.rs:1286-1318 */
static void bi_sub_unchecked(rctx* c, const bigchip* ch, const bint* a, const bint* b, bint* out) {
    int lw = ch->limb_width;
    if (a->n < b->n) { c->failed = 1; }
    int max_n = a->n;
    big ab, bb, cb; to_big_uint(a, lw, &ab); to_big_uint(b, lw, &bb);
    if (big_sub(&cb, &ab, &bb)) { c->failed = 1; big_zero(&cb); } /* BigUint underflow panics in the reference */
    out->n = max_n;
    for (int i = 0; i < max_n; i++) out->l[i] = range_assign(c, big_limb_fe(&cb, i, lw), sublimb_bit_len(lw), lw);
    bint added; bi_add(c, ch, b, out, &added);
    aval eq; bi_is_equal_fresh(c, a, &added, &eq);
    mg_assert_one(c, &eq);
}
/* chip.rs:310-373 */
static void bi_sub(rctx* c, const bigchip* ch, const bint* a, const bint* b, bint* out, aval* is_overflowed) {
    int n2 = b->n;
    bint max_int; bi_max_value(c, ch, n2, &max_int);
    bint inflated_a; bi_add(c, ch, a, &max_int, &inflated_a);
    bint inflated_subed; bi_sub_unchecked(c, ch, &inflated_a, b, &inflated_subed);
    aval one = mg_assign_bit(c, FR.one);
    aval is_not_overflowed = mg_is_equal(c, &inflated_subed.l[n2], &one);
    *is_overflowed = mg_not(c, &is_not_overflowed);
    int num_l = inflated_subed.n, num_r = a->n > n2 ? a->n : n2;
    aval zero = mg_assign_constant(c, fe_zero());
    bint sel_l, sel_r; sel_l.n = num_l; sel_r.n = num_r;
    for (int i = 0; i < num_l; i++)
        sel_l.l[i] = i >= n2 ? mg_select(c, &inflated_subed.l[i], &zero, &is_not_overflowed)
                             : mg_select(c, &inflated_subed.l[i], &b->l[i], &is_not_overflowed);
    for (int i = 0; i < num_r; i++) {
        if (i >= a->n) sel_r.l[i] = mg_select(c, &max_int.l[i], &zero, &is_not_overflowed);
        else if (i >= n2) sel_r.l[i] = mg_select(c, &zero, &a->l[i], &is_not_overflowed);
        else sel_r.l[i] = mg_select(c, &max_int.l[i], &a->l[i], &is_not_overflowed);
    }
    bi_sub_unchecked(c, ch, &sel_l, &sel_r, out);
}
/* chip.rs:386-419 */
static void bi_mul(rctx* c, const bint* a, const bint* b, bint* out) {
    int d0 = a->n, d1 = b->n, d = d0 + d1 - 1;
    out->n = d;
    for (int i = 0; i < d; i++) {
        aval acc = mg_assign_constant(c, fe_zero());
        int j = d1 >= i + 1 ? 0 : i + 1 - d1;
        while (j < d0 && j <= i) { int k = i - j; acc = mg_mul_add(c, &a->l[j], &b->l[k], &acc); j++; }
        out->l[i] = acc;
    }
}
/* chip.rs:1323-1349 */
static void bi_div_mod_main_gate(rctx* c, const aval* a, const aval* n, aval* q_out, aval* m_out) {
    uint64_t av[4], nv[4]; fe_canon(&a->v, av); fe_canon(&n->v, nv);
    big ab, nb, qb, rb; ab.n = nb.n = 8;
    for (int j = 0; j < 4; j++) { ab.w[2 * j] = (uint32_t)av[j]; ab.w[2 * j + 1] = (uint32_t)(av[j] >> 32); nb.w[2 * j] = (uint32_t)nv[j]; nb.w[2 * j + 1] = (uint32_t)(nv[j] >> 32); }
    big_norm(&ab); big_norm(&nb);
    if (big_divrem(&qb, &rb, &ab, &nb)) { c->failed = 1; big_zero(&qb); big_zero(&rb); }
    aval q = mg_assign_value(c, fe_of_big(&qb));
    aval m = mg_assign_value(c, fe_of_big(&rb));
    aval nq = mg_mul(c, n, &q);
    aval a_sub_nq = mg_sub(c, a, &nq);
    mg_assert_equal(c, &m, &a_sub_nq);
    *q_out = q; *m_out = m;
}
/* chip.rs:780-805 */
static void bi_is_equal_fresh(rctx* c, const bint* a, const bint* b, aval* out) {
    int n1 = a->n, n2 = b->n, a_larger = n1 > n2, max_n = a_larger ? n1 : n2;
    aval eq = mg_assign_bit(c, FR.one);
    for (int i = 0; i < max_n; i++) {
        aval flag;
        if (a_larger && i >= n2) flag = mg_is_zero(c, &a->l[i]);
        else if (!a_larger && i >= n1) flag = mg_is_zero(c, &b->l[i]);
        else flag = mg_is_equal(c, &a->l[i], &b->l[i]);
        eq = mg_and(c, &eq, &flag);
    }
    *out = eq;
}
/* chip.rs:822-895 */
static void bi_is_equal_muled(rctx* c, const bigchip* ch, const bint* a, const bint* b, int nl_l, int nl_r, aval* out) {
    int min_n = nl_r >= nl_l ? nl_l : nl_r, lw = ch->limb_width, num_limbs = nl_l + nl_r - 1;
    big word_max; u256_word_max(lw, min_n, &word_max);
    big wm2; big two; two.n = 1; two.w[0] = 2; big_mul(&wm2, &word_max, &two);
    int carry_bits = big_bits(&wm2) - lw;
    fe wm = fe_of_big(&word_max);
    aval limb_max = mg_assign_constant(c, limb_max_fe(lw));
    aval acc_extra = mg_assign_constant(c, fe_zero());
    aval carry = mg_assign_constant(c, fe_zero());
    aval eq = mg_assign_bit(c, FR.one);
    for (int i = 0; i < num_limbs; i++) {
        aval a_b = mg_sub(c, &a->l[i], &b->l[i]);
        aval sum = mg_add_with_constant(c, &a_b, &carry, wm);
        aval new_carry, cs; bi_div_mod_main_gate(c, &sum, &limb_max, &new_carry, &cs);
        acc_extra = mg_add_constant(c, &acc_extra, wm);
        aval q_acc, mod_acc; bi_div_mod_main_gate(c, &acc_extra, &limb_max, &q_acc, &mod_acc);
        aval cs_acc_eq = mg_is_equal(c, &cs, &mod_acc);
        eq = mg_and(c, &eq, &cs_acc_eq);
        acc_extra = q_acc;
        if (i < num_limbs - 1) {
            aval ra = range_assign(c, new_carry.v, sublimb_bit_len(carry_bits), carry_bits);
            aval range_eq = mg_is_equal(c, &new_carry, &ra);
            eq = mg_and(c, &eq, &range_eq);
        } else {
            aval fin = mg_is_equal(c, &new_carry, &acc_extra);
            eq = mg_and(c, &eq, &fin);
        }
        carry = new_carry;
    }
    *out = eq;
}
/* chip.rs:542-629 */
static void bi_mul_mod(rctx* c, const bigchip* ch, const bint* a, const bint* b, const bint* n, bint* out) {
    int lw = ch->limb_width, n1 = a->n, n2 = b->n;
    if (n1 != n->n) c->failed = 1;
    static __thread big ab, bb, nb, full, q, r;
    to_big_uint(a, lw, &ab); to_big_uint(b, lw, &bb); to_big_uint(n, lw, &nb);
    big_mul(&full, &ab, &bb);
    if (big_divrem(&q, &r, &full, &nb)) { c->failed = 1; big_zero(&q); big_zero(&r); }
    if (big_bits(&q) > lw * n2 || big_bits(&r) > lw * n1) c->failed = 1; /* chip.rs:583-584 asserts */
    bint qi, ri; qi.n = n2; ri.n = n1;
    for (int i = 0; i < n2; i++) qi.l[i] = range_assign(c, big_limb_fe(&q, i, lw), sublimb_bit_len(lw), lw);
    for (int i = 0; i < n1; i++) ri.l[i] = range_assign(c, big_limb_fe(&r, i, lw), sublimb_bit_len(lw), lw);
    bint abm, qn; bi_mul(c, a, b, &abm); bi_mul(c, &qi, n, &qn);
    int n_sum = n1 + n2;
    bint eq_a, eq_b; eq_a.n = eq_b.n = n_sum - 1;
    for (int i = 0; i < n_sum - 1; i++) {
        eq_a.l[i] = abm.l[i];
        eq_b.l[i] = i < n1 ? mg_add(c, &qn.l[i], &ri.l[i]) : qn.l[i];
    }
    aval eq; bi_is_equal_muled(c, ch, &eq_a, &eq_b, n1, n2, &eq);
    mg_assert_one(c, &eq);
    *out = ri;
}
/* chip.rs:710-742 */
static void bi_pow_mod_fixed_exp(rctx* c, const bigchip* ch, const bint* a, const uint8_t* e_le, int e_len, const bint* n, bint* out) {
    int nbits = 0;
    for (int i = e_len * 8 - 1; i >= 0; i--) if ((e_le[i >> 3] >> (i & 7)) & 1) { nbits = i + 1; break; }
    big one; one.n = 1; one.w[0] = 1;
    bint acc; bi_assign_constant(c, ch, &one, a->n, &acc);
    bint squared = *a;
    for (int i = 0; i < nbits; i++) {
        bint cur = squared;
        bi_mul_mod(c, ch, &cur, &cur, n, &squared);
        if (!((e_le[i >> 3] >> (i & 7)) & 1)) continue;
        bint t; bi_mul_mod(c, ch, &acc, &cur, n, &t); acc = t;
    }
    *out = acc;
}
/* chip.rs:908-1006, 1150-1158: assert_in_field = assert_one(is_less_than(a, n)) */
static void bi_assert_in_field(rctx* c, const bigchip* ch, const bint* a, const bint* n) {
    bint tmp; aval is_overflowed; bi_sub(c, ch, a, n, &tmp, &is_overflowed); /* is_less_than_or_equal */
    aval is_eq; bi_is_equal_fresh(c, a, n, &is_eq);
    aval not_eq = mg_not(c, &is_eq);
    aval lt = mg_and(c, &is_overflowed, &not_eq);
    mg_assert_one(c, &lt);
}

/* chip.rs:754-767 */
static aval bi_is_zero(rctx* c, const bint* a) {
    aval bit = mg_assign_bit(c, FR.one);
    for (int i = 0; i < a->n; i++) { aval z = mg_is_zero(c, &a->l[i]); bit = mg_and(c, &bit, &z); }
    return bit;
}
/* chip.rs:908-1006: the comparison family; `sub` reports its second output as 1 exactly when a <= b */
static aval bi_is_lte(rctx* c, const bigchip* ch, const bint* a, const bint* b) {
    bint tmp; aval is_overflowed; bi_sub(c, ch, a, b, &tmp, &is_overflowed); return is_overflowed;
}
static aval bi_is_lt(rctx* c, const bigchip* ch, const bint* a, const bint* b) {
    aval lte = bi_is_lte(c, ch, a, b);
    aval is_eq; bi_is_equal_fresh(c, a, b, &is_eq);
    aval not_eq = mg_not(c, &is_eq);
    return mg_and(c, &lte, &not_eq);
}
static aval bi_is_gt(rctx* c, const bigchip* ch, const bint* a, const bint* b) { aval lte = bi_is_lte(c, ch, a, b); return mg_not(c, &lte); }
static aval bi_is_gte(rctx* c, const bigchip* ch, const bint* a, const bint* b) { aval lt = bi_is_lt(c, ch, a, b); return mg_not(c, &lt); }

/* ---- MainGate::to_bits / compose (third-party maingate, restated from its published semantics) --------------
 * to_bits(composed, number_of_bits): every bit of the low number_of_bits bits of the value is assigned as a boolean
 * cell (assign_bit), least significant first; compose() lays the terms bit_i * 2^i out like decompose() does
 * (4 terms per row in a..d, running remainder with base -1 in e, rows chained by se_next = 1) and the composed cell
 * is constrained equal to the input.  A value wider than number_of_bits therefore breaks the copy constraint. */
static int mg_to_bits(rctx* c, const aval* composed, int number_of_bits, aval* bits) {
    uint64_t cv[4]; fe_canon(&composed->v, cv);
    fe total = fe_zero();
    for (int i = 0; i < number_of_bits; i++) {
        int b = (int)((cv[i >> 6] >> (i & 63)) & 1);
        bits[i] = mg_assign_bit(c, b ? FR.one : fe_zero());
        if (b) { uint64_t pc[4] = {0, 0, 0, 0}; pc[i >> 6] = (uint64_t)1 << (i & 63); fe p2 = fe_of_canon(pc); total = fe_add(&FR, &total, &p2); }
    }
    fe remaining = total;
    aval result; memset(&result, 0, sizeof result);
    int nchunks = (number_of_bits - 1) / 4 + 1;
    for (int ch = 0; ch < nchunks; ch++) {
        term t[5]; int cnt = 0;
        fe composed_chunk = fe_zero();
        for (int j = 4 * ch; j < 4 * ch + 4 && j < number_of_bits; j++) {
            uint64_t pc[4] = {0, 0, 0, 0}; pc[j >> 6] = (uint64_t)1 << (j & 63);
            fe base = fe_of_canon(pc);
            t[cnt++] = t_assigned(&bits[j], base);
            fe pr = fe_mul(&FR, &bits[j].v, &base); composed_chunk = fe_add(&FR, &composed_chunk, &pr);
        }
        while (cnt < 4) t[cnt++] = t_zero();
        t[4] = U_SUB(remaining);
        int is_final = ch == nchunks - 1;
        aval o[NADV];
        mg_apply(c, t, 5, fe_zero(), fe_zero(), fe_zero(), is_final ? fe_zero() : FE_ONE(), o);
        if (ch == 0) result = o[4];
        remaining = fe_sub(&FR, &remaining, &composed_chunk);
    }
    mg_assert_equal(c, &result, composed);
    return number_of_bits;
}

/* RefreshAux::new (src/big_integer/mod.rs:431-482): increased_limbs_vec.  Returns its length. */
static int refresh_aux(int limb_width, int num_limbs_l, int num_limbs_r, int* increased) {
    static big muled[2 * MAXL + 8];
    int nm = 0;
    big max_limb; { big m; m.n = limb_width / 32 + 1; for (int i = 0; i < m.n; i++) m.w[i] = 0; m.w[limb_width / 32] = 1u << (limb_width % 32);
                    big one; one.n = 1; one.w[0] = 1; big_sub(&max_limb, &m, &one); }
    big sq; big_mul(&sq, &max_limb, &max_limb);
    int d = num_limbs_l + num_limbs_r - 1;
    for (int i = 0; i < d; i++) {
        int j = num_limbs_r >= i + 1 ? 0 : i + 1 - num_limbs_r;
        big_zero(&muled[nm]);
        while (j < num_limbs_l && j <= i) { big_add_shifted_words(&muled[nm], sq.w, sq.n, 0); j++; }
        nm++;
    }
    int cnt = 0, cur_d = 0, max_d = d;
    while (cur_d <= max_d) {
        if (nm <= cur_d) { big_zero(&muled[nm]); nm++; }
        int bits = big_bits(&muled[cur_d]);
        int num_chunks = bits % limb_width == 0 ? bits / limb_width : bits / limb_width + 1;
        if (num_chunks == 0) return -1; /* usize underflow in the reference */
        increased[cnt++] = num_chunks - 1;
        /* cut muled[cur_d] into limb_width-bit chunks and add chunk j to muled[cur_d + j] */
        big whole = muled[cur_d];
        big_zero(&muled[cur_d]);
        for (int j = 0; j < num_chunks; j++) {
            uint32_t w[8] = {0}; int nw = 0;
            for (int b = 0; b < limb_width; b++) {
                int pbit = j * limb_width + b;
                if ((pbit >> 5) < whole.n && ((whole.w[pbit >> 5] >> (pbit & 31)) & 1)) { w[b >> 5] |= 1u << (b & 31); if ((b >> 5) + 1 > nw) nw = (b >> 5) + 1; }
            }
            if (nm <= cur_d + j) { big_zero(&muled[nm]); nm++; }
            if (nw) big_add_shifted_words(&muled[cur_d + j], w, nw, 0);
        }
        cur_d++;
    }
    return cnt;
}
/* chip.rs:168-233 */
static void bi_refresh(rctx* c, const bigchip* ch, const bint* a, int num_limbs_l, int num_limbs_r, bint* out) {
    int inc[2 * MAXL + 8];
    int lw = ch->limb_width;
    int num_fresh = refresh_aux(lw, num_limbs_l, num_limbs_r, inc);
    if (num_fresh < 0 || a->n != num_limbs_l + num_limbs_r - 1 || num_fresh > MAXL) { c->failed = 1; out->n = 0; return; }
    aval zero = mg_assign_constant(c, fe_zero());
    aval limbs[MAXL];
    for (int i = 0; i < a->n; i++) limbs[i] = a->l[i];
    for (int i = a->n; i < num_fresh; i++) limbs[i] = zero;
    aval limb_max = mg_assign_constant(c, limb_max_fe(lw));
    for (int i = 0; i < num_fresh; i++) {
        aval limb = limbs[i];
        for (int j = 0; j < inc[i] + 1; j++) {
            aval q, n; bi_div_mod_main_gate(c, &limb, &limb_max, &q, &n);
            if (j == 0) limbs[i] = n;
            else if (i + j >= num_fresh) { c->failed = 1; }
            else limbs[i + j] = mg_add(c, &limbs[i + j], &n);
            limb = q;
        }
        mg_assert_zero(c, &limb);
    }
    out->n = num_fresh;
    for (int i = 0; i < num_fresh; i++) {
        aval ra = range_assign(c, limbs[i].v, sublimb_bit_len(lw), lw);
        mg_assert_equal(c, &limbs[i], &ra);
        out->l[i] = limbs[i];
    }
}
/* chip.rs:452-481 */
static void bi_add_mod(rctx* c, const bigchip* ch, const bint* a, const bint* b, const bint* n, bint* out) {
    bint added; bi_add(c, ch, a, b, &added);
    bint subed; aval is_overflowed; bi_sub(c, ch, &added, n, &subed, &is_overflowed);
    int nl = subed.n;
    aval zero = mg_assign_constant(c, fe_zero());
    for (int i = added.n; i < nl; i++) added.l[i] = zero;
    aval res[MAXL];
    for (int i = 0; i < nl; i++) res[i] = mg_select(c, &added.l[i], &subed.l[i], &is_overflowed);
    for (int i = n->n; i < nl; i++) mg_assert_zero(c, &res[i]);
    out->n = n->n;
    for (int i = 0; i < n->n; i++) out->l[i] = res[i];
}
/* chip.rs:495-529 */
static void bi_sub_mod(rctx* c, const bigchip* ch, const bint* a, const bint* b, const bint* n, bint* out) {
    bint subed1, subed2; aval ov1, ov2;
    bi_sub(c, ch, a, b, &subed1, &ov1);
    bi_sub(c, ch, n, &subed1, &subed2, &ov2);
    mg_assert_zero(c, &ov2);
    int nl = subed2.n;
    aval zero = mg_assign_constant(c, fe_zero());
    for (int i = subed1.n; i < nl; i++) subed1.l[i] = zero;
    aval res[MAXL];
    for (int i = 0; i < nl; i++) res[i] = mg_select(c, &subed2.l[i], &subed1.l[i], &ov1);
    for (int i = n->n; i < nl; i++) mg_assert_zero(c, &res[i]);
    out->n = n->n;
    for (int i = 0; i < n->n; i++) out->l[i] = res[i];
}
/* chip.rs:664-696 */
static void bi_pow_mod(rctx* c, const bigchip* ch, const bint* a, const bint* e, const bint* n, int exp_limb_bits, bint* out) {
    static aval e_bits[MAXL * 64];
    int nb = 0;
    for (int i = 0; i < e->n; i++) nb += mg_to_bits(c, &e->l[i], exp_limb_bits, e_bits + nb);
    big one; one.n = 1; one.w[0] = 1;
    bint acc; bi_assign_constant(c, ch, &one, ch->num_limbs, &acc);
    bint squared = *a;
    for (int i = 0; i < nb; i++) {
        bint muled; bi_mul_mod(c, ch, &acc, &squared, n, &muled);
        for (int j = 0; j < acc.n; j++) acc.l[j] = mg_select(c, &muled.l[j], &acc.l[j], &e_bits[i]);
        bint sq = squared; bi_mul_mod(c, ch, &sq, &sq, n, &squared);
    }
    *out = acc;
}

/* ---- RSAChip (src/chip.rs) --------------------------------------------------------------------- */
/* chip.rs:128-199 */
static aval rsa_verify_pkcs1v15_ex(rctx* c, const bigchip* ch, int bits_len, const uint8_t* e_le, int e_len, const bint* e_var,
                                   int exp_limb_bits, const bint* n, const bint* hashed, const bint* sig) {
    aval is_eq = mg_assign_constant(c, FR.one);
    bigchip chip = *ch;
    bi_assert_in_field(c, &chip, sig, n);                     /* modpow_public_key, chip.rs:99-114 */
    bint powed;
    if (e_var) bi_pow_mod(c, &chip, sig, e_var, n, exp_limb_bits, &powed);   /* AssignedRSAPubE::Var */
    else bi_pow_mod_fixed_exp(c, &chip, sig, e_le, e_len, n, &powed);        /* AssignedRSAPubE::Fix */
    int hash_len = 4, nl = bits_len / 64;
    for (int i = 0; i < hash_len; i++) { aval e = mg_is_equal(c, &powed.l[i], &hashed->l[i]); is_eq = mg_and(c, &is_eq, &e); }
    aval p1 = mg_assign_constant(c, fe_of_u64(217300885422736416ull));
    aval p2 = mg_assign_constant(c, fe_of_u64(938447882527703397ull));
    aval e1 = mg_is_equal(c, &powed.l[hash_len], &p1);
    aval e2 = mg_is_equal(c, &powed.l[hash_len + 1], &p2);
    is_eq = mg_and(c, &is_eq, &e1); is_eq = mg_and(c, &is_eq, &e2);
    uint64_t v6[4]; fe_canon(&powed.l[hash_len + 2].v, v6);
    /* low = v % 2^32, high = v / 2^32 (full quotient of the canonical integer) */
    fe low = fe_of_u64(v6[0] & 0xffffffffull);
    uint64_t hi[4] = {(v6[0] >> 32) | (v6[1] << 32), (v6[1] >> 32) | (v6[2] << 32), (v6[2] >> 32) | (v6[3] << 32), v6[3] >> 32};
    fe high = fe_of_canon(hi);
    aval rl = range_assign(c, low, 4, 32), rh = range_assign(c, high, 4, 32);
    aval u32a = mg_assign_constant(c, fe_of_u64(1ull << 32));
    aval concat = mg_mul_add(c, &rh, &u32a, &rl);
    mg_assert_equal(c, &powed.l[hash_len + 2], &concat);
    aval p32 = mg_assign_constant(c, fe_of_u64(3158320));
    aval e3 = mg_is_equal(c, &rl, &p32); is_eq = mg_and(c, &is_eq, &e3);
    aval ff32 = mg_assign_constant(c, fe_of_u64(4294967295ull));
    aval e4 = mg_is_equal(c, &rh, &ff32); is_eq = mg_and(c, &is_eq, &e4);
    aval ff64 = mg_assign_constant(c, fe_of_u64(18446744073709551615ull));
    for (int i = hash_len + 3; i < nl - 1; i++) { aval e = mg_is_equal(c, &powed.l[i], &ff64); is_eq = mg_and(c, &is_eq, &e); }
    aval last = mg_assign_constant(c, fe_of_u64(562949953421311ull));
    aval e5 = mg_is_equal(c, &powed.l[nl - 1], &last); is_eq = mg_and(c, &is_eq, &e5);
    return is_eq;
}
static aval rsa_verify_pkcs1v15(rctx* c, const bigchip* ch, int bits_len, const uint8_t* e_le, int e_len,
                                const bint* n, const bint* hashed, const bint* sig) {
    return rsa_verify_pkcs1v15_ex(c, ch, bits_len, e_le, e_len, NULL, 0, n, hashed, sig);
}

/* ---- table lifetime / configuration --------------------------------------------------------------- */
typedef struct { rctx c; } orc_table;

/* RSAChip::compute_range_lens (chip.rs:249-254) + BigIntChip::compute_range_lens (big_integer/chip.rs:1220-1249):
 * distinct non-zero bit lengths get table tags 1.. in ascending order */
static void config_tags(rctx* c, int limb_width, int num_limbs) {
    int lens[8], nlens = 0;
    int out_comp = limb_width / 8, out_over = limb_width % out_comp;
    int fresh_carry_bits = (limb_width + 2) - limb_width; /* bits(2 * 2^lw) - lw = 2 */
    int fresh_comp = sublimb_bit_len(fresh_carry_bits), fresh_over = fresh_carry_bits % fresh_comp;
    big wm; u256_word_max(limb_width, num_limbs, &wm); big two; two.n = 1; two.w[0] = 2; big wm2; big_mul(&wm2, &wm, &two);
    int mul_carry_bits = big_bits(&wm2) - limb_width;
    int mul_comp = sublimb_bit_len(mul_carry_bits), mul_over = mul_carry_bits % mul_comp;
    int all[7] = {out_comp, fresh_comp, mul_comp, 32 / 8, out_over, fresh_over, mul_over};
    for (int i = 0; i < 7; i++) { int v = all[i], dup = 0; if (!v) continue; for (int j = 0; j < nlens; j++) if (lens[j] == v) dup = 1; if (!dup) lens[nlens++] = v; }
    for (int i = 0; i < nlens; i++) for (int j = i + 1; j < nlens; j++) if (lens[j] < lens[i]) { int t = lens[i]; lens[i] = lens[j]; lens[j] = t; }
    memset(c->tag_of_bits, 0, sizeof c->tag_of_bits);
    for (int i = 0; i < nlens; i++) c->tag_of_bits[lens[i]] = i + 1;
}

orc_table* orc_table_new(unsigned k, int limb_width, int num_limbs) {
    orc_table* t = calloc(1, sizeof *t);
    rctx* c = &t->c;
    c->k = k; c->nrows = (size_t)1 << k;
    for (int i = 0; i < NADV; i++) c->adv[i] = calloc(c->nrows, sizeof(fe));
    for (int i = 0; i < NFIX; i++) c->fix[i] = calloc(c->nrows, sizeof(fe));
    c->s_comp = calloc(c->nrows, 1); c->tag_comp = calloc(c->nrows, 1); c->s_over = calloc(c->nrows, 1); c->tag_over = calloc(c->nrows, 1);
    config_tags(c, limb_width, num_limbs);
    return t;
}
void orc_table_free(orc_table* t) {
    rctx* c = &t->c;
    for (int i = 0; i < NADV; i++) free(c->adv[i]);
    for (int i = 0; i < NFIX; i++) free(c->fix[i]);
    free(c->s_comp); free(c->tag_comp); free(c->s_over); free(c->tag_over); free(c->copies); free(t);
}
/* copies the advice columns out: out[col][row], Montgomery form (halo2curves memory format) */
void orc_table_advice(const orc_table* t, fe* out) {
    for (int i = 0; i < NADV; i++) memcpy(out + (size_t)i * t->c.nrows, t->c.adv[i], t->c.nrows * sizeof(fe));
}
uint64_t orc_table_rows(const orc_table* t) { return t->c.offset; }
int orc_table_failed(const orc_table* t) { return t->c.failed; }

/* The bench circuit's synthesize (benches/bench.rs:132-225, sha2 disabled): region 1 assigns the
 * signature then the public key, region 2 the hash + verify_pkcs1v15_signature, region 3 assert_one.
 * SimpleFloorPlanner stacks the regions (all use the same columns), so one running offset.
 * Returns the is_valid cell's value (1 / 0), or -1 if the reference would have panicked. */
int orc_rsa_synthesize(orc_table* t, int bits_len, const uint8_t* e_le, int e_len, const uint64_t* n_limbs,
                       const uint64_t* sig_limbs, const uint64_t* hash_limbs) {
    rctx* c = &t->c;
    bigchip ch = {64, bits_len / 64};
    int nl = bits_len / 64;
    fe tmp[MAXL];
    bint sig, n, hashed;
    for (int i = 0; i < nl; i++) tmp[i] = fe_of_u64(sig_limbs[i]);
    bi_assign_integer(c, &ch, tmp, nl, &sig);                 /* assign_signature, chip.rs:80-88 */
    for (int i = 0; i < nl; i++) tmp[i] = fe_of_u64(n_limbs[i]);
    bi_assign_integer(c, &ch, tmp, nl, &n);                   /* assign_public_key, chip.rs:58-70 (e fixed) */
    for (int i = 0; i < 4; i++) tmp[i] = fe_of_u64(hash_limbs[i]);
    bi_assign_integer(c, &ch, tmp, 4, &hashed);               /* bench.rs:192-202 */
    aval is_valid = rsa_verify_pkcs1v15(c, &ch, bits_len, e_le, e_len, &n, &hashed, &sig);
    mg_assert_one(c, &is_valid);                              /* bench.rs:213-221 */
    if (c->failed) return -1;
    return fe_eq(&is_valid.v, &FR.one) ? 1 : 0;
}

/* RSASignatureVerifier::verify_pkcs1v15_signature (src/lib.rs:183-248) from the digest bytes on: the 32 byte cells the
 * external SHA-256 chip returns (stood in for by one assign_value row each, in the order of `hashed_bytes` after the
 * reverse at src/lib.rs:213, i.e. least significant byte first), composed into four limbs by assign_constant / mul_add
 * (src/lib.rs:222-236), then RSAChip::verify_pkcs1v15_signature in the same region.  digest_le: the 32 bytes, least
 * significant first (= the 4 hash limbs of orc_rsa_synthesize as little-endian bytes).  Returns is_valid, -1 on panic. */
int orc_rsa_synthesize_digest(orc_table* t, int bits_len, const uint8_t* e_le, int e_len, const uint64_t* n_limbs,
                              const uint64_t* sig_limbs, const uint8_t* digest_le) {
    rctx* c = &t->c;
    bigchip ch = {64, bits_len / 64};
    int nl = bits_len / 64;
    fe tmp[MAXL];
    bint sig, n, hashed;
    for (int i = 0; i < nl; i++) tmp[i] = fe_of_u64(sig_limbs[i]);
    bi_assign_integer(c, &ch, tmp, nl, &sig);
    for (int i = 0; i < nl; i++) tmp[i] = fe_of_u64(n_limbs[i]);
    bi_assign_integer(c, &ch, tmp, nl, &n);
    aval bytes[32];
    for (int i = 0; i < 32; i++) bytes[i] = mg_assign_value(c, fe_of_u64(digest_le[i]));
    hashed.n = 4;
    for (int i = 0; i < 4; i++) {
        aval limb = mg_assign_constant(c, fe_zero());
        for (int j = 0; j < 8; j++) {
            aval coeff = mg_assign_constant(c, fe_of_u64((uint64_t)1 << (8 * j)));
            limb = mg_mul_add(c, &coeff, &bytes[8 * i + j], &limb);
        }
        hashed.l[i] = limb;
    }
    aval is_valid = rsa_verify_pkcs1v15(c, &ch, bits_len, e_le, e_len, &n, &hashed, &sig);
    if (c->failed) return -1;
    return fe_eq(&is_valid.v, &FR.one) ? 1 : 0;
}

/* the same circuit with RSAPubE::Var (src/chip.rs:58-70, :99-114): the exponent is an assigned one-limb integer,
 * pow_mod walks exp_limb_bits of its bits.  Region order as above; e is assigned right after n. */
int orc_rsa_synthesize_var(orc_table* t, int bits_len, int exp_limb_bits, uint64_t e_word, const uint64_t* n_limbs,
                           const uint64_t* sig_limbs, const uint64_t* hash_limbs) {
    rctx* c = &t->c;
    bigchip ch = {64, bits_len / 64};
    int nl = bits_len / 64;
    fe tmp[MAXL];
    bint sig, n, e, hashed;
    for (int i = 0; i < nl; i++) tmp[i] = fe_of_u64(sig_limbs[i]);
    bi_assign_integer(c, &ch, tmp, nl, &sig);
    for (int i = 0; i < nl; i++) tmp[i] = fe_of_u64(n_limbs[i]);
    bi_assign_integer(c, &ch, tmp, nl, &n);
    tmp[0] = fe_of_u64(e_word);
    bi_assign_integer(c, &ch, tmp, 1, &e);
    for (int i = 0; i < 4; i++) tmp[i] = fe_of_u64(hash_limbs[i]);
    bi_assign_integer(c, &ch, tmp, 4, &hashed);
    aval is_valid = rsa_verify_pkcs1v15_ex(c, &ch, bits_len, NULL, 0, &e, exp_limb_bits, &n, &hashed, &sig);
    mg_assert_one(c, &is_valid);
    if (c->failed) return -1;
    return fe_eq(&is_valid.v, &FR.one) ? 1 : 0;
}

/* ---- single BigIntChip operations, as the reference's in-file test circuits drive them -------------
 * (src/big_integer/chip.rs:1470-3264: impl_bigint_test_circuit! bodies).  Integers come in as
 * little-endian 64-bit words.  out receives the result limbs as canonical 4 x u64 each.
 *   op 0  test_mul_case*:     a,b = assign_constant_fresh; ab = mul(a,b); ans = assign_constant_muled(n);
 *                              assert_equal_muled(ab, ans)               out = 2*nl-1 unreduced limbs of ab
 *   op 1  test_mulmod_case*:   a,b,n = assign_integer; r = mul_mod(a,b,n)  out = nl limbs of r
 *   op 2  test_pow_mod_fixed_exp: a,n = assign_integer; r = pow_mod_fixed_exp(a, e=b_words[0], n)
 *   op 3  test_add:            c = add(a,b)                               out = nl+1 limbs
 *   op 4  test_sub:            (c, overflow) = sub(a,b)                   out = nl limbs, then overflow bit
 *   op 5  test_assert_in_field: assert_in_field(a, n)
 *   op 6  test_refresh (chip.rs:1861-1899): ab = mul(a,b), ba = mul(b,a), both refreshed with RefreshAux(64, nl, nl),
 *                              assert_equal_fresh                         out = 2*nl refreshed limbs of ab
 *   op 7  test_add_mod (chip.rs:1948-1986): a,b,n = assign_integer; r = add_mod(a,b,n)   out = nl limbs
 *   op 8  test_sub_mod (chip.rs:2027-2070): r = sub_mod(a,b,n)                            out = nl limbs
 *   op 9  test_pow_mod (chip.rs:2229-2271): a = assign_integer, e = assign_integer([e_word]) (as RSAPubE::Var),
 *                              n = assign_integer; r = pow_mod(a, e, n, exp_limb_bits)    out = nl limbs
 *         (e_word = b_words[0], exp_limb_bits = n_words_len for this op)
 * returns the number of out limbs, or -1 when the reference would have panicked. */
static void words_to_big(const uint64_t* w, int nwords, big* out) {
    out->n = 2 * nwords;
    for (int i = 0; i < nwords; i++) { out->w[2 * i] = (uint32_t)w[i]; out->w[2 * i + 1] = (uint32_t)(w[i] >> 32); }
    big_norm(out);
}
int orc_bigint_op(orc_table* t, int op, int bits_len, const uint64_t* a_words, const uint64_t* b_words,
                  const uint64_t* n_words, int n_words_len, uint64_t* out) {
    rctx* c = &t->c;
    int nl = bits_len / 64;
    bigchip ch = {64, nl};
    fe tmp[MAXL];
    bint a, b, n, r;
    int nout = 0;
    if (op == 0) {
        big ab_, bb_, nb_;
        words_to_big(a_words, nl, &ab_); words_to_big(b_words, nl, &bb_); words_to_big(n_words, n_words_len, &nb_);
        bi_assign_constant(c, &ch, &ab_, nl, &a);
        bi_assign_constant(c, &ch, &bb_, nl, &b);
        bi_mul(c, &a, &b, &r);
        bint ans; bi_assign_constant(c, &ch, &nb_, 2 * nl - 1, &ans);
        aval eq; bi_is_equal_muled(c, &ch, &r, &ans, nl, nl, &eq);
        mg_assert_one(c, &eq);
        nout = r.n;
    } else {
        for (int i = 0; i < nl; i++) tmp[i] = fe_of_u64(a_words[i]);
        bi_assign_integer(c, &ch, tmp, nl, &a);
        if (op == 9) { tmp[0] = fe_of_u64(b_words[0]); bi_assign_integer(c, &ch, tmp, 1, &b); }
        else if (op != 2) { for (int i = 0; i < nl; i++) tmp[i] = fe_of_u64(b_words[i]); bi_assign_integer(c, &ch, tmp, nl, &b); }
        if (op == 1 || op == 2 || op == 5 || (op >= 7 && op <= 9) || op == 18) { for (int i = 0; i < nl; i++) tmp[i] = fe_of_u64(n_words[i]); bi_assign_integer(c, &ch, tmp, nl, &n); }
        if (op == 1) { bi_mul_mod(c, &ch, &a, &b, &n, &r); nout = r.n; }
        else if (op == 2) {
            uint8_t e_le[8]; for (int i = 0; i < 8; i++) e_le[i] = (uint8_t)(b_words[0] >> (8 * i));
            bi_pow_mod_fixed_exp(c, &ch, &a, e_le, 8, &n, &r); nout = r.n;
        } else if (op == 3) { bi_add(c, &ch, &a, &b, &r); nout = r.n; }
        else if (op == 4) { aval ov; bi_sub(c, &ch, &a, &b, &r, &ov); r.l[r.n] = ov; nout = r.n + 1; }
        else if (op == 5) { bi_assert_in_field(c, &ch, &a, &n); nout = 0; }
        else if (op == 6) {
            bint ab, ba, ba_r; bi_mul(c, &a, &b, &ab); bi_mul(c, &b, &a, &ba);
            bi_refresh(c, &ch, &ab, nl, nl, &r); bi_refresh(c, &ch, &ba, nl, nl, &ba_r);
            aval eq; bi_is_equal_fresh(c, &r, &ba_r, &eq); mg_assert_one(c, &eq);
            nout = r.n;
        }
        else if (op == 7) { bi_add_mod(c, &ch, &a, &b, &n, &r); nout = r.n; }
        else if (op == 8) { bi_sub_mod(c, &ch, &a, &b, &n, &r); nout = r.n; }
        else if (op == 9) { bi_pow_mod(c, &ch, &a, &b, &n, n_words_len, &r); nout = r.n; }
        else if (op >= 10 && op <= 16) {   /* the predicates: one result cell */
            r.n = 1; nout = 1;
            switch (op) {
                case 10: r.l[0] = bi_is_zero(c, &a); break;
                case 11: bi_is_equal_fresh(c, &a, &b, &r.l[0]); break;
                case 12: case 16: r.l[0] = bi_is_lt(c, &ch, &a, &b); break;   /* is_in_field(a, n) = is_less_than(a, n) */
                case 13: r.l[0] = bi_is_lte(c, &ch, &a, &b); break;
                case 14: r.l[0] = bi_is_gt(c, &ch, &a, &b); break;
                default: r.l[0] = bi_is_gte(c, &ch, &a, &b); break;
            }
        }
        else if (op == 17) { bi_mul(c, &a, &a, &r); nout = r.n; }            /* square, chip.rs:430-437 */
        else if (op == 18) { bi_mul_mod(c, &ch, &a, &a, &n, &r); nout = r.n; } /* square_mod, chip.rs:642-649 */
        else return -2;
    }
    for (int i = 0; i < nout; i++) fe_canon(&r.l[i].v, out + 4 * i);
    return c->failed ? -1 : nout;
}

/* keygen-side exports for the oracle prover (oracle/plonk.py): fixed columns, range selectors/tags,
 * copy constraints, bits per lookup tag */
void orc_table_fixed(const orc_table* t, fe* out) {
    for (int i = 0; i < NFIX; i++) memcpy(out + (size_t)i * t->c.nrows, t->c.fix[i], t->c.nrows * sizeof(fe));
}
void orc_table_range(const orc_table* t, uint8_t* out /* [4][nrows]: s_comp, tag_comp, s_over, tag_over */) {
    size_t n = t->c.nrows;
    memcpy(out, t->c.s_comp, n); memcpy(out + n, t->c.tag_comp, n); memcpy(out + 2 * n, t->c.s_over, n); memcpy(out + 3 * n, t->c.tag_over, n);
}
size_t orc_table_copies(const orc_table* t, uint32_t* out, size_t cap) {
    size_t n = t->c.ncopies < cap ? t->c.ncopies : cap;
    if (out) memcpy(out, t->c.copies, n * 4 * sizeof(uint32_t));
    return t->c.ncopies;
}
void orc_table_tag_bits(const orc_table* t, int* out /* [16], bits of tag i, 0 = unused */) {
    for (int i = 0; i < 16; i++) out[i] = 0;
    for (int b = 1; b < 80; b++) if (t->c.tag_of_bits[b] && t->c.tag_of_bits[b] < 16) out[t->c.tag_of_bits[b]] = b;
}

/* layout digest, compared with the product's host-side circuit recorder (tests/test_host_circuit.py):
 * out[0] = rows used, out[1] = order-independent hash of all fixed cells (canonical values),
 * out[2] = number of copy constraints, out[3] = order-independent hash of the copy constraints,
 * out[4] = hash of the range selectors / tags */
static uint64_t mix64(uint64_t h, uint64_t v) { h ^= v; h *= 0x100000001B3ull; h ^= h >> 29; return h; }
void orc_table_layout_digest(const orc_table* t, uint64_t* out) {
    const rctx* c = &t->c;
    uint64_t hf = 0, hc = 0, hr = 0;
    for (size_t r = 0; r < c->offset && r < c->nrows; r++) {
        for (int f = 0; f < NFIX; f++) {
            uint64_t v[4]; fe_canon(&c->fix[f][r], v);
            if (!(v[0] | v[1] | v[2] | v[3])) continue;
            uint64_t h = mix64(mix64(0xcbf29ce484222325ull, r), f);
            for (int i = 0; i < 4; i++) h = mix64(h, v[i]);
            hf += h;
        }
        uint64_t h = mix64(mix64(mix64(mix64(mix64(0xcbf29ce484222325ull, r), c->s_comp[r]), c->tag_comp[r]), c->s_over[r]), c->tag_over[r]);
        if (c->s_comp[r] | c->s_over[r]) hr += h;
    }
    for (size_t i = 0; i < c->ncopies; i++) {
        const uint32_t* e = c->copies[i];
        hc += mix64(mix64(mix64(mix64(0xcbf29ce484222325ull, e[0]), e[1]), e[2]), e[3]);
    }
    out[0] = c->offset; out[1] = hf; out[2] = c->ncopies; out[3] = hc; out[4] = hr;
}

/* ---- MockProver-style check ----------------------------------------------------------------------------- */
/* returns the number of violated constraints (0 = satisfied); first few are described in msg */
long orc_check(const orc_table* t, char* msg, size_t msg_cap) {
    const rctx* c = &t->c;
    long bad = 0; size_t mlen = 0;
    if (msg_cap) msg[0] = 0;
#define REPORT(...) do { if (bad < 8 && mlen + 160 < msg_cap) mlen += snprintf(msg + mlen, msg_cap - mlen, __VA_ARGS__); bad++; } while (0)
    size_t usable = c->nrows - 6; /* blinding_factors 5 + 1 */
    if (c->offset > usable) REPORT("rows used %zu > usable %zu\n", c->offset, usable);
    /* lookup table: tag -> 2^bits values; (0,0) for disabled rows */
    int bits_of_tag[16] = {0}; int ntags = 0;
    for (int b = 1; b < 80; b++) if (c->tag_of_bits[b]) { bits_of_tag[c->tag_of_bits[b]] = b; if (c->tag_of_bits[b] > ntags) ntags = c->tag_of_bits[b]; }
    size_t table_rows = 1; for (int tg = 1; tg <= ntags; tg++) table_rows += (size_t)1 << bits_of_tag[tg];
    if (table_rows > usable) REPORT("lookup table rows %zu > usable\n", table_rows);
    for (size_t r = 0; r < c->offset && r < c->nrows; r++) {
        fe acc = c->fix[F_CONST][r];
        for (int i = 0; i < NADV; i++) { fe p = fe_mul(&FR, &c->adv[i][r], &c->fix[F_SA + i][r]); acc = fe_add(&FR, &acc, &p); }
        fe ab = fe_mul(&FR, &c->adv[0][r], &c->adv[1][r]); ab = fe_mul(&FR, &ab, &c->fix[F_MUL_AB][r]); acc = fe_add(&FR, &acc, &ab);
        fe cd = fe_mul(&FR, &c->adv[2][r], &c->adv[3][r]); cd = fe_mul(&FR, &cd, &c->fix[F_MUL_CD][r]); acc = fe_add(&FR, &acc, &cd);
        if (r + 1 < c->nrows) { fe en = fe_mul(&FR, &c->adv[4][r + 1], &c->fix[F_SE_NEXT][r]); acc = fe_add(&FR, &acc, &en); }
        if (!fe_is_zero(&acc)) REPORT("main gate violated at row %zu\n", r);
        if (c->s_comp[r]) {
            int bits = bits_of_tag[c->tag_comp[r]];
            for (int i = 0; i < 4; i++) { uint64_t v[4]; fe_canon(&c->adv[i][r], v); if (v[1] | v[2] | v[3] || (bits < 64 && v[0] >> bits)) REPORT("composition lookup failed row %zu col %d\n", r, i); }
        }
        if (c->s_over[r]) {
            int bits = bits_of_tag[c->tag_over[r]]; uint64_t v[4]; fe_canon(&c->adv[0][r], v);
            if (v[1] | v[2] | v[3] || (v[0] >> bits)) REPORT("overflow lookup failed row %zu\n", r);
        }
    }
    for (size_t i = 0; i < c->ncopies; i++) {
        const uint32_t* e = c->copies[i];
        if (!fe_eq(&c->adv[e[0]][e[1]], &c->adv[e[2]][e[3]])) REPORT("copy constraint violated (%u,%u)=(%u,%u)\n", e[0], e[1], e[2], e[3]);
    }
    return bad;
}
