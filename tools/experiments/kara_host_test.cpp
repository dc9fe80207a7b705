// host check of tools/experiments/field_kara.cuh against Field<P>::mul:  g++ -O1 -std=c++17 kara_host_test.cpp && ./a.out
#include <cstdio>
#include <cstdlib>
#include "field_kara.cuh"
using namespace b2r;
static uint64_t st = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (uint32_t)(st >> 16); }
template <class P> static fe_t rand_fe(int mode) {
    fe_t x;
    for (;;) {
        for (int i = 0; i < 8; i++) x.l[i] = mode == 1 ? 0xffffffffu : rnd();
        x.l[7] &= 0x3fffffffu;
        if (mode == 2) { for (int i = 0; i < 8; i++) x.l[i] = P::MOD(i); x.l[0] -= 1 + (rnd() & 3); }
        if (mode == 3) { for (int i = 0; i < 8; i++) x.l[i] = 0; x.l[rnd() & 7] = rnd(); }
        uint32_t t[8], m[8];
        for (int i = 0; i < 8; i++) m[i] = P::MOD(i);
        if (sub8(t, x.l, m)) return x;
        if (mode == 1) mode = 0;
    }
}
template <class P> static int run(int n) {
    int bad = 0;
    for (int it = 0; it < n; it++) {
        fe_t a = rand_fe<P>(it % 4), b = rand_fe<P>((it / 4) % 4);
        fe_t w = Field<P>::mul(a, b), g = FieldKara<P>::mul(a, b);
        if (!Field<P>::eq(w, g)) bad++;
    }
    return bad;
}
int main() {
    int b1 = run<FrP>(200000), b2 = run<FqP>(200000);
    printf("mismatches: fr %d fq %d\n", b1, b2);
    return b1 || b2;
}
