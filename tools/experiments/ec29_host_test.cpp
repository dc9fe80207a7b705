// Host check of the 29-bit-limb XYZZ point arithmetic (ec29.cuh) against ec.cuh (32-bit limbs), both
// compiled for the CPU: random signed accumulations, XYZZ + XYZZ, doublings, and the exceptional cases
// (P + P, P - P, identity operands).  Prints "ok <n>" or "FAIL ...".
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ec29.cuh"

using namespace b2r;

static uint64_t st = 88172645463325252ull;
static uint64_t rnd() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; }

static affine_t to_table(const affine_t& p) {  // Montgomery-2^256 -> packed canonical Montgomery-2^261
    fe_t c = Fq::zero();
    c.l[0] = 32;
    c = Fq::to_mont(c);
    affine_t r;
    r.x = Fq::mul(p.x, c);
    r.y = Fq::mul(p.y, c);
    return r;
}
static bool same(const xyzz_t& a, const xyzz29_t& b) {
    affine_t x = xyzz_to_affine(a), y = xyzz29_to_affine256(b);
    return Fq::eq(x.x, y.x) && Fq::eq(x.y, y.y);
}
int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 300;
    const int NP = 64;
    std::vector<affine_t> pts(NP), tab(NP);
    affine_t g;
    g.x = Fq::one();
    g.y = Fq::dbl(Fq::one());
    xyzz_t run = xyzz_identity();
    for (int i = 0; i < NP; i++) {
        xyzz_madd(run, g, false);
        pts[i] = xyzz_to_affine(run);
        tab[i] = to_table(pts[i]);
    }
    affine_t idp;
    idp.x = Fq::zero();
    idp.y = Fq::zero();
    long checks = 0;
    for (int it = 0; it < iters; it++) {
        xyzz_t a = xyzz_identity(), b = xyzz_identity();
        xyzz29_t a9 = xyzz29_identity(), b9 = xyzz29_identity();
        const int len = 1 + (int)(rnd() % 40);
        for (int j = 0; j < len; j++) {
            int idx = (int)(rnd() % NP);
            bool neg = rnd() & 1;
            int special = (int)(rnd() % 16);
            affine_t p = pts[idx], t = tab[idx];
            if (special == 0) { p = idp; t = idp; }
            xyzz_madd_ls(a, p, neg);
            xyzz29_madd_ls(a9, affine29_unpack(t), neg);
            if (special == 1) {  // force P + P (doubling path) then P - P - P
                xyzz_t s = xyzz_from_affine_signed(p, neg);
                xyzz29_t s9 = xyzz29_from_affine_signed(affine29_unpack(t), neg);
                xyzz_madd_ls(s, p, neg);
                xyzz29_madd_ls(s9, affine29_unpack(t), neg);
                if (!same(s, s9)) { printf("FAIL P+P it=%d\n", it); return 1; }
                xyzz_t z = xyzz_from_affine_signed(p, neg);
                xyzz29_t z9 = xyzz29_from_affine_signed(affine29_unpack(t), neg);
                xyzz_madd_ls(z, p, !neg);
                xyzz29_madd_ls(z9, affine29_unpack(t), !neg);
                if (!xyzz_is_identity(z) || !xyzz29_is_identity(z9)) { printf("FAIL P-P it=%d\n", it); return 1; }
                xyzz_add_ls(a, s);
                xyzz29_add_ls(a9, s9);
                checks += 2;
            }
            if (!same(a, a9)) { printf("FAIL madd it=%d j=%d\n", it, j); return 1; }
            checks++;
            if (j == len / 2) { b = a; b9 = a9; }
        }
        // XYZZ + XYZZ, including acc + acc (doubling) and acc + identity
        xyzz_t c = a; xyzz29_t c9 = a9;
        xyzz_add_ls(c, b); xyzz29_add_ls(c9, b9);
        if (!same(c, c9)) { printf("FAIL add it=%d\n", it); return 1; }
        xyzz_t d = a; xyzz29_t d9 = a9;
        xyzz_add_ls(d, a); xyzz29_add_ls(d9, a9);
        if (!same(d, d9) || !same(xyzz_double(a), xyzz29_double(a9))) { printf("FAIL double it=%d\n", it); return 1; }
        xyzz_t e = xyzz_identity(); xyzz29_t e9 = xyzz29_identity();
        xyzz_add_ls(e, a); xyzz29_add_ls(e9, a9);
        xyzz_add_ls(e, xyzz_identity()); xyzz29_add_ls(e9, xyzz29_identity());
        if (!same(e, e9) || !same(a, e9)) { printf("FAIL identity add it=%d\n", it); return 1; }
        // long doubling chain keeps the bounds
        xyzz_t f = a; xyzz29_t f9 = a9;
        for (int j = 0; j < 20; j++) { f = xyzz_double(f); f9 = xyzz29_double(f9); xyzz_add_ls(f, c); xyzz29_add_ls(f9, c9); }
        if (!same(f, f9)) { printf("FAIL double chain it=%d\n", it); return 1; }
        checks += 5;
    }
    printf("ok %ld\n", checks);
    return 0;
}
