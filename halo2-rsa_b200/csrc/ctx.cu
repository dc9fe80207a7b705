// Context, error reporting and device-memory helpers of the C ABI (include/b2rsa.h).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>

#include "ctx.hpp"

namespace b2r {

int32_t fail(b2r_ctx* ctx, int32_t code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}

int32_t cuda_fail(b2r_ctx* ctx, cudaError_t e, const char* what) {
    if (ctx) {
        ctx->err = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    }
    return e == cudaErrorMemoryAllocation ? B2R_ERR_NOMEM : B2R_ERR_CUDA;
}

int32_t scratch_get(b2r_ctx* ctx, int slot, size_t bytes, void** out) {
    Scratch& s = ctx->scratch[slot];
    if (s.cap < bytes) {
        if (s.p) {
            // stream-ordered: earlier kernels on the stream may still use the old block
            B2R_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            B2R_CUDA(ctx, cudaFree(s.p));
            s.p = nullptr;
            s.cap = 0;
        }
        size_t cap = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
        B2R_CUDA(ctx, cudaMalloc(&s.p, cap));
        s.cap = cap;
    }
    *out = s.p;
    return 0;
}

}  // namespace b2r

using namespace b2r;

static thread_local std::string g_create_err;

extern "C" {

const char* b2r_version(void) { return "b2rsa 0.1 (sm_100a)"; }

int32_t b2r_ctx_create(int32_t device, b2r_ctx** out) {
    if (!out) return B2R_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_err = "no CUDA device: libb2rsa has no CPU fallback";
        return B2R_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) return B2R_ERR_INVALID;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return B2R_ERR_CUDA;
    if (prop.major != 10) {
        g_create_err = "device is not sm_100 (kernels are built for sm_100a only)";
        return B2R_ERR_NO_DEVICE;
    }
    b2r::DeviceGuard guard(device);   // the stream is created on `device`; the caller's current device is restored on return
    b2r_ctx* ctx = new (std::nothrow) b2r_ctx();
    if (!ctx) return B2R_ERR_NOMEM;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return B2R_ERR_CUDA;
    }
    ctx->own_stream = true;
    *out = ctx;
    return 0;
}

int32_t b2r_ctx_destroy(b2r_ctx* ctx) {
    if (!ctx) return B2R_ERR_INVALID;
    b2r::DeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& kv : ctx->twiddles) cudaFree(kv.second);
    for (auto& s : ctx->scratch)
        if (s.p) cudaFree(s.p);
    for (auto& r : ctx->prof) { cudaEventDestroy(r.start); cudaEventDestroy(r.stop); }
    for (auto& e : ctx->prof_pool) cudaEventDestroy(e);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->commit_log) cudaFree(ctx->commit_log);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}

int32_t b2r_ctx_set_stream(b2r_ctx* ctx, void* cuda_stream) try {
    B2R_ENTER(ctx);
    B2R_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (cuda_stream) {
        ctx->stream = (cudaStream_t)cuda_stream;
        ctx->own_stream = false;
    } else {
        B2R_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    return 0;
} B2R_ABI_CATCH(ctx)

int32_t b2r_ctx_sync(b2r_ctx* ctx) try {
    B2R_ENTER(ctx);
    B2R_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
} B2R_ABI_CATCH(ctx)

const char* b2r_last_error(const b2r_ctx* ctx) {
    if (!ctx) return g_create_err.c_str();
    return ctx->err.c_str();
}

uint64_t b2r_launch_count(const b2r_ctx* ctx) { return ctx ? ctx->launches : 0; }

int32_t b2r_profile_enable(b2r_ctx* ctx, int32_t on) try {
    B2R_ENTER(ctx);
    ctx->profile = on != 0;
    return 0;
} B2R_ABI_CATCH(ctx)

// sums the recorded launches whose name matches `name` exactly; clears nothing
int32_t b2r_profile_read(b2r_ctx* ctx, const char* name, double* total_ms, uint64_t* launches, double* units) try {
    B2R_ENTER(ctx);
    if (!ctx || !name) return B2R_ERR_INVALID;
    B2R_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double ms = 0, u = 0;
    uint64_t cnt = 0;
    for (auto& r : ctx->prof) {
        if (strcmp(r.name, name) != 0) continue;
        float t = 0;
        B2R_CUDA(ctx, cudaEventElapsedTime(&t, r.start, r.stop));
        ms += t;
        u += r.units;
        cnt++;
    }
    if (total_ms) *total_ms = ms;
    if (launches) *launches = cnt;
    if (units) *units = u;
    return 0;
} B2R_ABI_CATCH(ctx)

// writes "name:ms:launches;..." for every distinct name, then clears the records
int32_t b2r_profile_dump(b2r_ctx* ctx, char* buf, size_t cap, int32_t clear) try {
    B2R_ENTER(ctx);
    if (!ctx || !buf || cap == 0) return B2R_ERR_INVALID;
    B2R_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::map<std::string, std::pair<double, uint64_t>> acc;
    for (auto& r : ctx->prof) {
        float t = 0;
        B2R_CUDA(ctx, cudaEventElapsedTime(&t, r.start, r.stop));
        auto& a = acc[r.name];
        a.first += t;
        a.second++;
    }
    std::string out;
    for (auto& kv : acc) {
        char tmp[160];
        snprintf(tmp, sizeof tmp, "%s:%.6f:%llu;", kv.first.c_str(), kv.second.first, (unsigned long long)kv.second.second);
        out += tmp;
    }
    if (out.size() + 1 > cap) return fail(ctx, B2R_ERR_INVALID, "profile_dump: buffer too small");
    memcpy(buf, out.c_str(), out.size() + 1);
    if (clear) {
        for (auto& r : ctx->prof) {
            ctx->prof_pool.push_back(r.start);
            ctx->prof_pool.push_back(r.stop);
        }
        ctx->prof.clear();
    }
    return 0;
} B2R_ABI_CATCH(ctx)

int32_t b2r_dev_alloc(b2r_ctx* ctx, size_t bytes, void** dptr) try {
    B2R_ENTER(ctx);
    if (!ctx || !dptr) return B2R_ERR_INVALID;
    B2R_CUDA(ctx, cudaMalloc(dptr, bytes ? bytes : 1));
    return 0;
} B2R_ABI_CATCH(ctx)
int32_t b2r_dev_free(b2r_ctx* ctx, void* dptr) try {
    B2R_ENTER(ctx);
    B2R_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    B2R_CUDA(ctx, cudaFree(dptr));
    return 0;
} B2R_ABI_CATCH(ctx)
int32_t b2r_h2d(b2r_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) try {
    B2R_ENTER(ctx);
    B2R_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    B2R_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
} B2R_ABI_CATCH(ctx)
int32_t b2r_d2h(b2r_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) try {
    B2R_ENTER(ctx);
    B2R_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    B2R_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
} B2R_ABI_CATCH(ctx)

}  // extern "C"
