// Host check of the 29-bit-limb lazy Montgomery arithmetic (field29.cuh) against the 32-bit-limb
// reference implementation (field.cuh), both compiled for the CPU.  Prints "ok <n>" or "FAIL ...".
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include "field29.cuh"

using namespace b2r;

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() {
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return rng_state;
}
template <class P>
static fe_t rand_fe(int kind) {
    fe_t r;
    for (int i = 0; i < 4; i++) { uint64_t w = rnd(); r.l[2 * i] = (uint32_t)w; r.l[2 * i + 1] = (uint32_t)(w >> 32); }
    if (kind == 1) for (int i = 1; i < 8; i++) r.l[i] = 0;           // small
    if (kind == 2) { for (int i = 0; i < 8; i++) r.l[i] = P::MOD(i); r.l[0] -= 1 + (uint32_t)(rnd() % 3); }  // p - small
    if (kind == 3) for (int i = 0; i < 8; i++) r.l[i] = 0;           // zero
    r.l[7] &= 0x3fffffffu;
    Field<P>::final_sub(r.l);
    return r;
}
template <class P>
static int run(const char* name, int iters) {
    using F = Field<P>;
    using G = F29<P>;
    int checks = 0;
    for (int it = 0; it < iters; it++) {
        fe_t a = rand_fe<P>(it % 7 == 0 ? 1 : it % 11 == 0 ? 2 : it % 13 == 0 ? 3 : 0), b = rand_fe<P>(it % 5 == 0 ? 2 : 0), c = rand_fe<P>(0);
        // pack / unpack
        if (!F::eq(G::pack(G::unpack(a)), a)) { printf("FAIL %s pack/unpack\n", name); return -1; }
        typename G::el A = G::from_mont256(a), B = G::from_mont256(b), C = G::from_mont256(c);
        if (!F::eq(G::to_mont256(A), a)) { printf("FAIL %s domain round trip\n", name); return -1; }
        if (!F::eq(G::to_mont256(G::mul(A, B)), F::mul(a, b))) { printf("FAIL %s mul it=%d\n", name, it); return -1; }
        if (!F::eq(G::to_mont256(G::sqr(A)), F::sqr(a))) { printf("FAIL %s sqr it=%d\n", name, it); return -1; }
        // lazy chain: ((a + b) - c) * (c - a) + 2 b - (a b)   with bounds: A,B,C < 2p
        typename G::el s1 = G::add(A, B);                    // < 4p
        typename G::el s2 = G::template sub<2>(s1, C);       // < 6p
        typename G::el s3 = G::template sub<2>(C, A);        // < 4p
        typename G::el pr = G::mul(s2, s3);                  // 24/128 -> < 2p
        typename G::el s4 = G::add(pr, G::dbl(B));           // < 6p
        typename G::el s5 = G::template sub<2>(s4, G::mul(A, B));  // < 8p
        fe_t want = F::sub(F::add(F::mul(F::sub(F::add(a, b), c), F::sub(c, a)), F::dbl(b)), F::mul(a, b));
        if (!F::eq(G::to_mont256(s5), want)) { printf("FAIL %s lazy chain it=%d\n", name, it); return -1; }
        // squares of lazy values, neg
        typename G::el n8 = G::template neg<8>(s5);
        if (!F::eq(G::to_mont256(n8), F::neg(want))) { printf("FAIL %s neg it=%d\n", name, it); return -1; }
        if (!F::eq(G::to_mont256(G::sqr(s5)), F::sqr(want))) { printf("FAIL %s lazy sqr it=%d\n", name, it); return -1; }
        // zero tests: x - x in lazy form is a non-trivial multiple of p
        typename G::el z = G::template sub<8>(s5, s5);
        if (!G::is_zero_mod_p(z)) { printf("FAIL %s zero test (8p) it=%d\n", name, it); return -1; }
        typename G::el z2 = G::template sub<2>(A, A);
        if (!G::is_zero_mod_p(z2) || !G::is_zero_limbs(G::reduce(z2))) { printf("FAIL %s zero test (2p) it=%d\n", name, it); return -1; }
        if (G::is_zero_mod_p(s5) != F::is_zero(want)) { printf("FAIL %s zero test value it=%d\n", name, it); return -1; }
        // extreme bound: 64 * (p - 1)-ish accumulations then multiply by a 2p-bounded value (64 * 2 = 128)
        typename G::el big = A;
        fe_t bigw = a;
        for (int j = 0; j < 5; j++) { big = G::dbl(big); bigw = F::dbl(bigw); }   // < 64p
        if (!F::eq(G::to_mont256(G::mul(big, B)), F::mul(bigw, b))) { printf("FAIL %s big-bound mul it=%d\n", name, it); return -1; }
        if (!F::eq(G::pack(G::reduce(big)), G::pack(G::reduce(G::from_mont256(bigw))))) { printf("FAIL %s reduce it=%d\n", name, it); return -1; }
        // mul by exact zero stays exactly zero
        if (!G::is_zero_limbs(G::mul(G::zero(), s5))) { printf("FAIL %s zero mul\n", name); return -1; }
        checks += 14;
    }
    return checks;
}
int main(int argc, char** argv) {
    int iters = argc > 1 ? atoi(argv[1]) : 20000;
    int a = run<FrP>("Fr", iters);
    if (a < 0) return 1;
    int b = run<FqP>("Fq", iters);
    if (b < 0) return 1;
    printf("ok %d\n", a + b);
    return 0;
}
