// Internal context shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b2rsa.h"
#include "field.cuh"

namespace b2r {

struct TwiddleKey {
    uint32_t log_n;
    uint64_t w[4];
    bool operator<(const TwiddleKey& o) const {
        if (log_n != o.log_n) return log_n < o.log_n;
        for (int i = 0; i < 4; i++)
            if (w[i] != o.w[i]) return w[i] < o.w[i];
        return false;
    }
};

struct ProfRec {
    const char* name;
    cudaEvent_t start, stop;
    double units;  // caller-defined work units of this launch (e.g. MSM terms)
};

struct Scratch {
    void* p = nullptr;
    size_t cap = 0;
};

}  // namespace b2r

struct b2r_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    std::string err;
    uint64_t launches = 0;
    uint64_t prove_calls = 0;  // nonce of the 64-bit-seed prove entry points: one blinding stream per call
    // twiddle tables: omega^e, e < n/2, Montgomery form, device resident
    std::map<b2r::TwiddleKey, b2r::fe_t*> twiddles;
    // grow-only scratch arenas (ping-pong buffers, MSM work arrays, staging)
    b2r::Scratch scratch[8];
    // optional per-kernel timing (b2r_profile_*): event pairs around selected launches
    bool profile = false;
    std::vector<b2r::ProfRec> prof;
    std::vector<cudaEvent_t> prof_pool;
    void* pinned = nullptr;  // small pinned staging block
    size_t pinned_cap = 0;
    // commitments of the last prove call, device resident: [batch][31] affine points in transcript order
    // (b2r_last_commitments; the block a multi-GPU host all-gathers)
    void* commit_log = nullptr;
    size_t commit_log_cap = 0;    // proofs
    size_t commit_log_batch = 0;  // proofs of the last completed call (0 = none)
};

namespace b2r {

int32_t fail(b2r_ctx* ctx, int32_t code, const std::string& msg);
int32_t cuda_fail(b2r_ctx* ctx, cudaError_t e, const char* what);
// returns device pointer of at least `bytes` from arena `slot` (contents undefined)
int32_t scratch_get(b2r_ctx* ctx, int slot, size_t bytes, void** out);

// Every ABI entry point runs on ITS context's device whatever the calling thread's current device is (a host that
// keeps one context per GPU in one process - INTEGRATION.md section 4 - must not have to cudaSetDevice itself), and
// restores the caller's device on the way out.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess || prev != device) {
            cudaSetDevice(device);
            switched = prev >= 0;
        }
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define B2R_ENTER(ctx)                        \
    if (!(ctx)) return B2R_ERR_INVALID;       \
    b2r::DeviceGuard _b2r_device_guard((ctx)->device)
// the ABI never unwinds into the caller (a Rust / C host): entry points are function-try-blocks closed by this
#define B2R_ABI_CATCH(ctx)                                                                                   \
    catch (const std::bad_alloc&) { return b2r::fail((b2r_ctx*)(ctx), B2R_ERR_NOMEM, "out of host memory"); }  \
    catch (const std::exception& e) { return b2r::fail((b2r_ctx*)(ctx), B2R_ERR_INVALID, e.what()); }          \
    catch (...) { return b2r::fail((b2r_ctx*)(ctx), B2R_ERR_INVALID, "unknown exception"); }

#define B2R_CUDA(ctx, call)                                          \
    do {                                                             \
        cudaError_t _e = (call);                                     \
        if (_e != cudaSuccess) return b2r::cuda_fail(ctx, _e, #call); \
    } while (0)

#define B2R_TRY(expr)            \
    do {                         \
        int32_t _r = (expr);     \
        if (_r != 0) return _r;  \
    } while (0)

#define B2R_LAUNCH_CHECK(ctx)                                                  \
    do {                                                                       \
        (ctx)->launches++;                                                     \
        cudaError_t _e = cudaGetLastError();                                   \
        if (_e != cudaSuccess) return b2r::cuda_fail(ctx, _e, "kernel launch"); \
    } while (0)

// RAII scope that brackets kernel launches with events when profiling is on
struct KTimer {
    b2r_ctx* ctx;
    cudaEvent_t start = nullptr, stop = nullptr;
    const char* name;
    double units;
    KTimer(b2r_ctx* c, const char* n, double u = 0) : ctx(c), name(n), units(u) {
        if (!ctx->profile) return;
        auto get = [&]() {
            cudaEvent_t e;
            if (!ctx->prof_pool.empty()) { e = ctx->prof_pool.back(); ctx->prof_pool.pop_back(); }
            else cudaEventCreate(&e);
            return e;
        };
        start = get();
        stop = get();
        cudaEventRecord(start, ctx->stream);
    }
    ~KTimer() {
        if (!start) return;
        cudaEventRecord(stop, ctx->stream);
        ctx->prof.push_back({name, start, stop, units});
    }
};

// scratch arena slots
enum { SC_NTT_PING = 0, SC_NTT_PONG = 1, SC_MSM_A = 2, SC_MSM_B = 3, SC_MSM_C = 4, SC_STAGE = 5, SC_WIT = 6, SC_MISC = 7 };

// sha256.cu
int32_t sha256_msgs_dev(b2r_ctx* ctx, const uint8_t* d_msgs, const uint64_t* d_offsets, size_t batch, uint64_t* d_hash_limbs, uint32_t limb_stride,
                        uint8_t* d_digests);
int32_t sha256_check_offsets(b2r_ctx* ctx, const uint64_t* offsets, size_t batch);

// ntt.cu
int32_t ntt_get_twiddles(b2r_ctx* ctx, const fe_t& omega, uint32_t log_n, const fe_t** out);

}  // namespace b2r
