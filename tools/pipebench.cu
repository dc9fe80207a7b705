// Instruction-pipe microbenchmarks that size the field-arithmetic roofline on sm_100a:
// IMAD / IMAD.WIDE (plain and with carry chains) / DFMA issue rates per SM, and the throughput of
// the library's Montgomery product (Fq::mul) at several ILP / occupancy points.
// Development tool: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipebench pipebench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "experiments/ec29.cuh"
#include "experiments/field_f64.cuh"
#include "experiments/field_kara.cuh"

using namespace b2r;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("cuda error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ITERS = 2048;

__global__ void k_imad(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x0) : "r"(a), "r"(b));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x1) : "r"(a), "r"(b));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x2) : "r"(a), "r"(b));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x3) : "r"(a), "r"(b));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x4) : "r"(a), "r"(b));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x5) : "r"(a), "r"(b));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x6) : "r"(a), "r"(b));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x7) : "r"(a), "r"(b));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
}

__global__ void k_imad_wide(uint64_t* out, uint32_t a, uint32_t b) {
    uint64_t x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x0) : "r"(a), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x1) : "r"(a), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x2) : "r"(a), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x3) : "r"(a), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x4) : "r"(a), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x5) : "r"(a), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x6) : "r"(a), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x7) : "r"(a), "r"(b));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
}

// the carry-chained row pattern of field.cuh: two independent 8-word accumulators, 4 wide MADs each per row
__global__ void k_imad_wide_cc(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t X[8], Y[8], av[8];
    for (int i = 0; i < 8; i++) X[i] = threadIdx.x + i, Y[i] = threadIdx.x * 3 + i, av[i] = a + i;
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            row_mad_nc(X, &av[0], b + j);
            row_mad_nc(Y, &av[1], b + j);
        }
    }
    uint32_t r = 0;
    for (int i = 0; i < 8; i++) r ^= X[i] ^ Y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__global__ void k_dfma(double* out, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// DFMA and IMAD.WIDE interleaved: do the two pipes run concurrently?
__global__ void k_mixed(double* out, double a, double b, uint32_t ia, uint32_t ib) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    uint64_t y0 = threadIdx.x, y1 = y0 + 1, y2 = y0 + 2, y3 = y0 + 3;
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            x0 = fma(x0, a, b);
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y0) : "r"(ia), "r"(ib));
            x1 = fma(x1, a, b);
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y1) : "r"(ia), "r"(ib));
            x2 = fma(x2, a, b);
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y2) : "r"(ia), "r"(ib));
            x3 = fma(x3, a, b);
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y3) : "r"(ia), "r"(ib));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + (double)(y0 ^ y1 ^ y2 ^ y3);
}

// IMAD.WIDE interleaved 1:1 with an ALU-pipe instruction (independent streams): do fma and alu pipes overlap?
template <int KIND>
__global__ void k_mix_alu(uint64_t* out, uint32_t a, uint32_t b) {
    uint64_t y0 = threadIdx.x, y1 = y0 + 1, y2 = y0 + 2, y3 = y0 + 3;
    uint32_t x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
#define ALUOP(x)                                                                                          \
    if (KIND == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(a));                              \
    else if (KIND == 1) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));     \
    else if (KIND == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b));     \
    else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y0) : "r"(a), "r"(b));
            ALUOP(x0)
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y1) : "r"(a), "r"(b));
            ALUOP(x1)
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y2) : "r"(a), "r"(b));
            ALUOP(x2)
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y3) : "r"(a), "r"(b));
            ALUOP(x3)
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = y0 ^ y1 ^ y2 ^ y3 ^ x0 ^ x1 ^ x2 ^ x3;
}
// ALU-only streams
template <int KIND>
__global__ void k_alu(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            ALUOP(x0) ALUOP(x1) ALUOP(x2) ALUOP(x3) ALUOP(x4) ALUOP(x5) ALUOP(x6) ALUOP(x7)
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
}

template <int ILP>
__global__ void k_fqmul(fe_t* out, const fe_t* in, int iters) {
    fe_t x[ILP], y;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    y = in[t & 1023];
    for (int j = 0; j < ILP; j++) x[j] = in[(t + j + 1) & 1023];
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) x[j] = Fq::mul(x[j], y);
    }
    fe_t r = x[0];
    for (int j = 1; j < ILP; j++) r = Fq::add(r, x[j]);
    out[t] = r;
}


template <int ILP, bool SQR>
__global__ void k_f29mul(fe_t* out, const fe_t* in, int iters) {
    Fq29::el x[ILP], y;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    y = Fq29::unpack(in[t & 1023]);
    for (int j = 0; j < ILP; j++) x[j] = Fq29::unpack(in[(t + j + 1) & 1023]);
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) x[j] = SQR ? Fq29::sqr(x[j]) : Fq29::mul(x[j], y);
    }
    Fq29::el r = x[0];
    for (int j = 1; j < ILP; j++) r = Fq29::add(r, x[j]);
    out[t] = Fq29::pack(Fq29::reduce(r));
}

// mixed additions into a running XYZZ accumulator from a small L1-resident point table (random field elements:
// the formulas do not care whether the operands are on the curve)
__global__ void __launch_bounds__(128, 4) k_madd_old(fe_t* out, const affine_t* pts, int iters) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    xyzz_t acc = xyzz_from_affine_signed(pts[t & 255], false);
#pragma unroll 1
    for (int i = 0; i < iters; i++) xyzz_madd_ls(acc, pts[(t + i + 1) & 255], (i & 1) != 0);
    out[t] = Fq::add(Fq::add(acc.x, acc.y), Fq::add(acc.zz, acc.zzz));
}
__global__ void __launch_bounds__(128, 4) k_madd_29(fe_t* out, const affine_t* pts, int iters) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    xyzz29_t acc = xyzz29_from_affine_signed(affine29_unpack(pts[t & 255]), false);
#pragma unroll 1
    for (int i = 0; i < iters; i++) xyzz29_madd_ls(acc, affine29_unpack(pts[(t + i + 1) & 255]), (i & 1) != 0);
    out[t] = Fq29::pack(Fq29::reduce(Fq29::add(Fq29::add(acc.x, acc.y), Fq29::add(acc.zz, acc.zzz))));
}

template <int ILP, bool SQR>
__global__ void k_f64mul(fe_t* out, const fe_t* in, int iters) {
    fe_t x[ILP], y;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    y = in[t & 1023];
    for (int j = 0; j < ILP; j++) x[j] = in[(t + j + 1) & 1023];
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) x[j] = SQR ? FqF64::sqr(x[j]) : FqF64::mul(x[j], y);
    }
    fe_t r = x[0];
    for (int j = 1; j < ILP; j++) r = Fq::add(r, x[j]);
    out[t] = r;
}

template <int ILP>
__global__ void k_karamul(fe_t* out, const fe_t* in, int iters) {
    fe_t x[ILP], y;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    y = in[t & 1023];
    for (int j = 0; j < ILP; j++) x[j] = in[(t + j + 1) & 1023];
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) x[j] = FqKara::mul(x[j], y);
    }
    fe_t r = x[0];
    for (int j = 1; j < ILP; j++) r = Fq::add(r, x[j]);
    out[t] = r;
}

template <class F>
static double time_ms(F launch, int reps = 5) {
    cudaEvent_t s, e;
    CK(cudaEventCreate(&s));
    CK(cudaEventCreate(&e));
    launch();
    launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(s));
        launch();
        CK(cudaEventRecord(e));
        CK(cudaEventSynchronize(e));
        float ms;
        CK(cudaEventElapsedTime(&ms, s, e));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main(int argc, char** argv) {
    (void)argv;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    const double ghz = clk_khz / 1e6;
    printf("device %s, %d SMs, max clock %.3f GHz (rates below assume this clock)\n", prop.name, sms, ghz);
    void* buf;
    CK(cudaMalloc(&buf, (size_t)sms * 16 * 1024 * 32));
    CK(cudaMemset(buf, 0x11, (size_t)sms * 16 * 1024 * 32));
    if (argc > 1 && !strcmp(argv[1], "kara")) {
        fe_t* in0 = (fe_t*)buf;
        // valid field elements: clear the top bits of every element of the input block
        {
            fe_t* h = (fe_t*)malloc(1024 * 32 * 2);
            uint64_t st = 88172645463325252ull;
            for (int i = 0; i < 2048; i++) {
                for (int k = 0; k < 8; k++) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; h[i].l[k] = (uint32_t)(st >> 11); }
                h[i].l[7] &= 0x0fffffffu;
            }
            CK(cudaMemcpy(in0, h, 2048 * 32, cudaMemcpyHostToDevice));
            free(h);
        }
        fe_t* o1 = in0 + 4096;
        const int grid0 = sms * 8;
        fe_t* o2 = o1 + (size_t)grid0 * 256;
        k_fqmul<1><<<grid0, 256>>>(o1, in0, 64);
        k_karamul<1><<<grid0, 256>>>(o2, in0, 64);
        CK(cudaDeviceSynchronize());
        const size_t cnt = (size_t)grid0 * 256;
        fe_t* h1 = (fe_t*)malloc(cnt * 32);
        fe_t* h2 = (fe_t*)malloc(cnt * 32);
        CK(cudaMemcpy(h1, o1, cnt * 32, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h2, o2, cnt * 32, cudaMemcpyDeviceToHost));
        size_t bad = 0;
        for (size_t i = 0; i < cnt; i++) bad += memcmp(&h1[i], &h2[i], 32) != 0;
        printf("Karatsuba mul parity vs Fq::mul: %zu mismatches of %zu (64 chained products each)\n", bad, cnt);
        for (int cps : {2, 4, 8}) {
            for (int th : {128, 256}) {
                const int grid = sms * cps;
                const double m = (double)grid * th * 512;
                double tb1 = time_ms([&] { k_fqmul<1><<<grid, th>>>(o1, in0, 512); });
                double tb2 = time_ms([&] { k_fqmul<2><<<grid, th>>>(o1, in0, 512); });
                double t1 = time_ms([&] { k_karamul<1><<<grid, th>>>(o1, in0, 512); });
                double t2 = time_ms([&] { k_karamul<2><<<grid, th>>>(o1, in0, 512); });
                printf("warps/SM %2d  Fq::mul ILP1 %6.1f ILP2 %6.1f   Karatsuba ILP1 %6.1f ILP2 %6.1f  Gmul/s\n", cps * th / 32, m / tb1 / 1e6, 2 * m / tb2 / 1e6,
                       m / t1 / 1e6, 2 * m / t2 / 1e6);
            }
        }
        return 0;
    }
    if (argc > 1) {
        // profiling mode (ncu): one launch of each kernel of interest at full occupancy
        fe_t* in0 = (fe_t*)buf;
        fe_t* out0 = in0 + 1024;
        const int grid = sms * 8;
        k_imad_wide<<<grid, 256>>>((uint64_t*)buf, 3, 5);
        k_imad_wide_cc<<<grid, 256>>>((uint32_t*)buf, 3, 5);
        CK(cudaMemset(buf, 0x11, 1024 * 64));
        k_fqmul<1><<<grid, 256>>>(out0, in0, 512);
        k_f29mul<1, false><<<grid, 256>>>(out0, in0, 512);
        k_f64mul<1, false><<<grid, 256>>>(out0, in0, 512);
        k_f64mul<2, false><<<grid, 256>>>(out0, in0, 512);
        k_f29mul<2, true><<<grid, 256>>>(out0, in0, 512);
        k_madd_old<<<sms * 4, 128>>>(out0, (const affine_t*)buf, 256);
        k_madd_29<<<sms * 4, 128>>>(out0, (const affine_t*)buf, 256);
        CK(cudaDeviceSynchronize());
        return 0;
    }
    const int threads = 256;
    for (int cps : {2, 8}) {
        const int grid = sms * cps;
        const double ops = (double)grid * threads * ITERS * 64;
        double t;
        t = time_ms([&] { k_imad<<<grid, threads>>>((uint32_t*)buf, 3, 5); });
        printf("warps/SM %2d  IMAD          %7.1f /clk/SM\n", cps * 8, ops / (t * 1e-3) / (ghz * 1e9) / sms);
        t = time_ms([&] { k_imad_wide<<<grid, threads>>>((uint64_t*)buf, 3, 5); });
        printf("warps/SM %2d  IMAD.WIDE     %7.1f /clk/SM\n", cps * 8, ops / (t * 1e-3) / (ghz * 1e9) / sms);
        t = time_ms([&] { k_imad_wide_cc<<<grid, threads>>>((uint32_t*)buf, 3, 5); });
        printf("warps/SM %2d  IMAD.WIDE.cc  %7.1f /clk/SM (2 chains of 4)\n", cps * 8, ops / (t * 1e-3) / (ghz * 1e9) / sms);
        t = time_ms([&] { k_dfma<<<grid, threads>>>((double*)buf, 1.0000001, 0.5); });
        printf("warps/SM %2d  DFMA          %7.1f /clk/SM\n", cps * 8, ops / (t * 1e-3) / (ghz * 1e9) / sms);
        t = time_ms([&] { k_mixed<<<grid, threads>>>((double*)buf, 1.0000001, 0.5, 3, 5); });
        printf("warps/SM %2d  DFMA+IMAD.W   %7.1f /clk/SM (sum of both)\n", cps * 8, ops / (t * 1e-3) / (ghz * 1e9) / sms);
    }
    {
        const int grid = sms * 8;
        const double ops = (double)grid * threads * ITERS * 64;
        double t;
        t = time_ms([&] { k_alu<0><<<grid, threads>>>((uint32_t*)buf, 3, 5); });
        printf("IADD alone            %7.1f /clk/SM\n", ops / (t * 1e-3) / (ghz * 1e9) / sms);
        t = time_ms([&] { k_alu<1><<<grid, threads>>>((uint32_t*)buf, 3, 5); });
        printf("SHF alone             %7.1f /clk/SM\n", ops / (t * 1e-3) / (ghz * 1e9) / sms);
        t = time_ms([&] { k_alu<2><<<grid, threads>>>((uint32_t*)buf, 3, 5); });
        printf("LOP3 alone            %7.1f /clk/SM\n", ops / (t * 1e-3) / (ghz * 1e9) / sms);
        t = time_ms([&] { k_mix_alu<0><<<grid, threads>>>((uint64_t*)buf, 3, 5); });
        printf("IMAD.WIDE + IADD 1:1  %7.1f /clk/SM (sum)\n", ops / (t * 1e-3) / (ghz * 1e9) / sms);
        t = time_ms([&] { k_mix_alu<1><<<grid, threads>>>((uint64_t*)buf, 3, 5); });
        printf("IMAD.WIDE + SHF 1:1   %7.1f /clk/SM (sum)\n", ops / (t * 1e-3) / (ghz * 1e9) / sms);
        t = time_ms([&] { k_mix_alu<2><<<grid, threads>>>((uint64_t*)buf, 3, 5); });
        printf("IMAD.WIDE + LOP3 1:1  %7.1f /clk/SM (sum)\n", ops / (t * 1e-3) / (ghz * 1e9) / sms);
        t = time_ms([&] { k_mix_alu<3><<<grid, threads>>>((uint64_t*)buf, 3, 5); });
        printf("IMAD.WIDE + IMAD 1:1  %7.1f /clk/SM (sum)\n", ops / (t * 1e-3) / (ghz * 1e9) / sms);
    }
    fe_t* in = (fe_t*)buf;
    fe_t* out = in + 1024;
    const int iters = 512;
    for (int th : {128, 256}) {
        for (int cps : {1, 2, 3, 4, 6, 8}) {
            const int grid = sms * cps;
            double t1 = time_ms([&] { k_fqmul<1><<<grid, th>>>(out, in, iters); });
            double t2 = time_ms([&] { k_fqmul<2><<<grid, th>>>(out, in, iters); });
            double t4 = time_ms([&] { k_fqmul<4><<<grid, th>>>(out, in, iters); });
            const double m = (double)grid * th * iters;
            printf("Fq::mul  warps/SM %2d  ILP1 %6.1f  ILP2 %6.1f  ILP4 %6.1f  Gmul/s\n", cps * th / 32, m / t1 / 1e6, 2 * m / t2 / 1e6,
                   4 * m / t4 / 1e6);
        }
    }
    for (int th : {128, 256}) {
        for (int cps : {1, 2, 3, 4, 6, 8}) {
            const int grid = sms * cps;
            double t1 = time_ms([&] { k_f29mul<1, false><<<grid, th>>>(out, in, iters); });
            double t2 = time_ms([&] { k_f29mul<2, false><<<grid, th>>>(out, in, iters); });
            double t4 = time_ms([&] { k_f29mul<4, false><<<grid, th>>>(out, in, iters); });
            double s2 = time_ms([&] { k_f29mul<2, true><<<grid, th>>>(out, in, iters); });
            const double m = (double)grid * th * iters;
            printf("F29::mul warps/SM %2d  ILP1 %6.1f  ILP2 %6.1f  ILP4 %6.1f  Gmul/s   sqr ILP2 %6.1f Gsqr/s\n", cps * th / 32, m / t1 / 1e6,
                   2 * m / t2 / 1e6, 4 * m / t4 / 1e6, 2 * m / s2 / 1e6);
        }
    }
    {
        // canonical values so that the old arithmetic's preconditions hold
        CK(cudaMemset(buf, 0x11, 1024 * 64));
        const affine_t* pts = (const affine_t*)buf;
        const int mi = 256;
        for (int cps : {1, 2, 3, 4}) {
            const int grid = sms * cps;
            double to = time_ms([&] { k_madd_old<<<grid, 128>>>(out, pts, mi); });
            double tn = time_ms([&] { k_madd_29<<<grid, 128>>>(out, pts, mi); });
            const double m = (double)grid * 128 * mi;
            printf("madd  warps/SM %2d  32-bit limbs %6.2f   29-bit limbs %6.2f  Gadd/s\n", cps * 4, m / to / 1e6, m / tn / 1e6);
        }
    }
    {
        // FP64-assisted product: parity with Fq::mul on the device, then throughput
        CK(cudaMemset(buf, 0x11, 1024 * 64));
        fe_t* i0 = (fe_t*)buf;
        fe_t* o1 = i0 + 4096;
        fe_t* o2 = o1 + (size_t)sms * 8 * 256;
        const int grid = sms * 8;
        // pseudo-random canonical inputs: repeated squaring of the 0x11.. pattern
        k_fqmul<1><<<4, 256>>>(i0, i0, 3);
        CK(cudaDeviceSynchronize());
        k_fqmul<1><<<grid, 256>>>(o1, i0, 64);
        k_f64mul<1, false><<<grid, 256>>>(o2, i0, 64);
        CK(cudaDeviceSynchronize());
        const size_t cnt = (size_t)grid * 256;
        fe_t* h1 = (fe_t*)malloc(cnt * 32);
        fe_t* h2 = (fe_t*)malloc(cnt * 32);
        CK(cudaMemcpy(h1, o1, cnt * 32, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h2, o2, cnt * 32, cudaMemcpyDeviceToHost));
        size_t bad = 0;
        for (size_t i = 0; i < cnt; i++) bad += memcmp(&h1[i], &h2[i], 32) != 0;
        printf("F64-assisted mul parity vs Fq::mul: %zu mismatches of %zu (64 chained products each)\n", bad, cnt);
        free(h1); free(h2);
        for (int th : {128, 256}) {
            for (int cps : {1, 2, 3, 4, 6, 8}) {
                const int g2 = sms * cps;
                double t1 = time_ms([&] { k_f64mul<1, false><<<g2, th>>>(o2, i0, 512); });
                double t2 = time_ms([&] { k_f64mul<2, false><<<g2, th>>>(o2, i0, 512); });
                double s1 = time_ms([&] { k_f64mul<1, true><<<g2, th>>>(o2, i0, 512); });
                const double m = (double)g2 * th * 512;
                printf("F64::mul warps/SM %2d  ILP1 %6.1f  ILP2 %6.1f  Gmul/s   sqr ILP1 %6.1f Gsqr/s\n", cps * th / 32, m / t1 / 1e6, 2 * m / t2 / 1e6,
                       m / s1 / 1e6);
            }
        }
    }
    return 0;
}
