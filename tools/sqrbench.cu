// Development tool: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sqrbench sqrbench.cu
// Device parity of the dedicated Montgomery squaring (field.cuh Field::sqr: 36 + 64 wide MADs) against mul(a, a) for
// Fq and Fr, the throughput of both, and the throughput of the fused sums of products (mul_add_mul: 192 wide MADs for two
// products, dot4: 320 for four).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "../halo2-rsa_b200/csrc/field.cuh"
using namespace b2r;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <class F, int ILP, bool SQR>
__global__ void k_chain(fe_t* out, const fe_t* in, int iters) {
    fe_t x[ILP];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int j = 0; j < ILP; j++) x[j] = in[(t + j) & 1023];
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) x[j] = SQR ? F::sqr(x[j]) : F::mul(x[j], x[j]);
    }
    fe_t r = x[0];
    for (int j = 1; j < ILP; j++) r = F::add(r, x[j]);
    out[t] = r;
}
// OP 2: x = x*a + x*b (mul_add_mul), OP 3: x = x*a + x*b + x*a + x*b (dot4): dependent chains like k_chain
template <class F, int OP>
__global__ void k_fused(fe_t* out, const fe_t* in, int iters) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    fe_t x = in[t & 1023];
    const fe_t a = in[(t + 7) & 1023], b = in[(t + 13) & 1023];
#pragma unroll 1
    for (int i = 0; i < iters; i++) x = OP == 2 ? F::mul_add_mul(x, a, x, b) : F::dot4(x, a, x, b, x, a, x, b);
    out[t] = x;
}
template <class F>
__global__ void k_seed(fe_t* io, int n) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint32_t)n) return;
    fe_t x = F::zero();
    for (int i = 0; i < 8; i++) x.l[i] = (t + 1) * 0x9E3779B1u + i * 0x85EBCA77u;
    x.l[7] &= 0x0fffffffu;
    if (t % 97 == 0) { for (int i = 0; i < 8; i++) x.l[i] = F::one().l[i]; }
    if (t % 101 == 0) x = F::zero();
    if (t % 103 == 0) { x = F::neg(F::one()); }
    io[t] = x;
}
template <class Fn> static double time_ms(Fn f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < 3; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}
template <class F> static void run(const char* name, int sms) {
    const int grid = sms * 8, th = 128;
    const size_t cnt = (size_t)grid * th;
    fe_t *in, *o1, *o2;
    CK(cudaMalloc(&in, 1024 * 32)); CK(cudaMalloc(&o1, cnt * 32)); CK(cudaMalloc(&o2, cnt * 32));
    k_seed<F><<<4, 256>>>(in, 1024);
    for (int iters : {1, 2, 37}) {
        k_chain<F, 1, false><<<grid, th>>>(o1, in, iters);
        k_chain<F, 1, true><<<grid, th>>>(o2, in, iters);
        CK(cudaDeviceSynchronize());
        fe_t* h1 = (fe_t*)malloc(cnt * 32); fe_t* h2 = (fe_t*)malloc(cnt * 32);
        CK(cudaMemcpy(h1, o1, cnt * 32, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h2, o2, cnt * 32, cudaMemcpyDeviceToHost));
        size_t bad = 0;
        for (size_t i = 0; i < cnt; i++) bad += memcmp(&h1[i], &h2[i], 32) != 0;
        printf("%s sqr parity vs mul(a,a), %d chained: %zu mismatches of %zu\n", name, iters, bad, cnt);
        free(h1); free(h2);
    }
    for (int cps : {2, 4, 8}) {
        const int g = sms * cps;
        const double m = (double)g * th * 512;
        double tm1 = time_ms([&] { k_chain<F, 1, false><<<g, th>>>(o1, in, 512); });
        double ts1 = time_ms([&] { k_chain<F, 1, true><<<g, th>>>(o1, in, 512); });
        double tm2 = time_ms([&] { k_chain<F, 2, false><<<g, th>>>(o1, in, 512); });
        double ts2 = time_ms([&] { k_chain<F, 2, true><<<g, th>>>(o1, in, 512); });
        printf("%s warps/SM %2d  mul ILP1 %6.1f ILP2 %6.1f   sqr ILP1 %6.1f ILP2 %6.1f  G/s\n", name, cps * th / 32, m / tm1 / 1e6, 2 * m / tm2 / 1e6,
               m / ts1 / 1e6, 2 * m / ts2 / 1e6);
    }
    for (int cps : {4, 8}) {
        const int g = sms * cps;
        const double m = (double)g * th * 512;
        double t2 = time_ms([&] { k_fused<F, 2><<<g, th>>>(o1, in, 512); });
        double t4 = time_ms([&] { k_fused<F, 3><<<g, th>>>(o1, in, 512); });
        printf("%s warps/SM %2d  mul_add_mul %6.1f G/s = %6.1f G products/s   dot4 %6.1f G/s = %6.1f G products/s\n", name, cps * th / 32, m / t2 / 1e6,
               2 * m / t2 / 1e6, m / t4 / 1e6, 4 * m / t4 / 1e6);
    }
    cudaFree(in); cudaFree(o1); cudaFree(o2);
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    run<Fq>("Fq", p.multiProcessorCount);
    run<Fr>("Fr", p.multiProcessorCount);
    return 0;
}
