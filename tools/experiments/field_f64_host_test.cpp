// Host check of the FP64-assisted Montgomery product (csrc/field_f64.cuh) against Field<P>::mul (csrc/field.cuh).
// The two DFMA steps are emulated with 128-bit integers on the host; everything else (limb split, column sums from
// the bit patterns, digit layout, word-serial reduction) is the device code.  Prints "ok <n>" or "FAIL ...".
#include <cstdio>
#include <cstdlib>
#include "field_f64.cuh"

using namespace b2r;

static uint64_t st = 0x243F6A8885A308D3ull;
static uint64_t rnd() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; }

template <class P>
static fe_t rand_fe(int kind) {
    fe_t r;
    for (int i = 0; i < 4; i++) { uint64_t w = rnd(); r.l[2 * i] = (uint32_t)w; r.l[2 * i + 1] = (uint32_t)(w >> 32); }
    if (kind == 1) for (int i = 1; i < 8; i++) r.l[i] = 0;
    if (kind == 2) { for (int i = 0; i < 8; i++) r.l[i] = P::MOD(i); r.l[0] -= 1 + (uint32_t)(rnd() % 3); }
    if (kind == 3) for (int i = 0; i < 8; i++) r.l[i] = 0;
    if (kind == 4) for (int i = 0; i < 8; i++) r.l[i] = 0xffffffffu;   // all-ones limbs before the clamp below
    r.l[7] &= 0x3fffffffu;
    Field<P>::final_sub(r.l);
    return r;
}
template <class P>
static long run(const char* name, int iters) {
    using F = Field<P>;
    using G = FieldF64<P>;
    for (int it = 0; it < iters; it++) {
        fe_t a = rand_fe<P>(it % 7 == 0 ? 1 : it % 11 == 0 ? 2 : it % 13 == 0 ? 3 : it % 17 == 0 ? 4 : 0);
        fe_t b = rand_fe<P>(it % 5 == 0 ? 2 : it % 19 == 0 ? 4 : 0);
        if (!F::eq(G::mul(a, b), F::mul(a, b))) { printf("FAIL %s mul it=%d\n", name, it); return -1; }
        if (!F::eq(G::sqr(a), F::sqr(a))) { printf("FAIL %s sqr it=%d\n", name, it); return -1; }
        if (!F::eq(G::mul(a, F::one()), a)) { printf("FAIL %s mul by one it=%d\n", name, it); return -1; }
    }
    return 3L * iters;
}
int main(int argc, char** argv) {
    int iters = argc > 1 ? atoi(argv[1]) : 100000;
    long a = run<FrP>("Fr", iters);
    if (a < 0) return 1;
    long b = run<FqP>("Fq", iters);
    if (b < 0) return 1;
    printf("ok %ld\n", a + b);
    return 0;
}
