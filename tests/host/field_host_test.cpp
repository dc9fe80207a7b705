// Host-side exerciser for the device field algorithm in csrc/field.cuh (compiled by g++:
// the row primitives fall back to their C emulation).  Prints hex vectors; the pytest
// (tests/test_field_host.py) checks them against Python integers.
#include <cstdio>
#include <cstdlib>
#include "../../halo2-rsa_b200/csrc/field.cuh"
using namespace b2r;

static uint64_t st = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() {
    st ^= st << 13; st ^= st >> 7; st ^= st << 17;
    return (uint32_t)(st >> 16);
}
template <class P>
static fe_t rand_fe(int mode) {
    fe_t x;
    for (;;) {
        for (int i = 0; i < 8; i++) x.l[i] = rnd();
        if (mode == 1) for (int i = 0; i < 8; i++) x.l[i] = 0xffffffffu;
        x.l[7] &= 0x3fffffffu;
        if (mode == 2) { for (int i = 0; i < 8; i++) x.l[i] = P::MOD(i); x.l[0] -= 1 + (rnd() & 3); }
        if (mode == 3) { for (int i = 1; i < 8; i++) x.l[i] = 0; x.l[0] = rnd() & 3; }
        uint32_t t[8], m[8];
        for (int i = 0; i < 8; i++) m[i] = P::MOD(i);
        if (sub8(t, x.l, m)) return x;  // x < p
        if (mode == 1) mode = 0;
    }
}
static void pr(const fe_t& x) {
    for (int i = 7; i >= 0; i--) printf("%08x", x.l[i]);
    printf(" ");
}
template <class P>
static void run(const char* name, int n) {
    using F = Field<P>;
    for (int it = 0; it < n; it++) {
        fe_t a = rand_fe<P>(it % 5 == 4 ? 2 : (it % 7 == 6 ? 3 : 0));
        fe_t b = rand_fe<P>(it % 11 == 10 ? 2 : 0);
        printf("%s ", name);
        pr(a); pr(b); pr(F::mul(a, b)); pr(F::add(a, b)); pr(F::sub(a, b)); pr(F::neg(a));
        if (it < 8) pr(F::inv(a)); else pr(it < 200 ? F::inv_vartime(a) : F::zero());
        pr(F::sqr(a));
        // fused sums of products (one reduction): extremes p - 1..p - 4 in every operand every 13th vector
        fe_t c = rand_fe<P>(it % 13 == 12 ? 2 : (it % 17 == 16 ? 3 : 0)), d = rand_fe<P>(it % 13 == 12 ? 2 : 0);
        if (it % 13 == 12) { a = rand_fe<P>(2); b = rand_fe<P>(2); }
        pr(a); pr(b); pr(c); pr(d);
        pr(F::mul_add_mul(a, b, c, d)); pr(F::mul_sub_mul(a, b, c, d)); pr(F::dot4(a, b, c, d, a, c, b, d));
        printf("\n");
    }
}
int main(int argc, char** argv) {
    int n = argc > 1 ? atoi(argv[1]) : 1000;
    run<FrP>("fr", n);
    run<FqP>("fq", n);
    return 0;
}
