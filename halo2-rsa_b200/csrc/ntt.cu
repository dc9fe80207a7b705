// Radix-2 Stockham (autosort) NTT over BN254 Fr for sm_100a.
//
// Replaces halo2_proofs::arithmetic::best_fft and the EvaluationDomain wrappers around it
// (lagrange_to_coeff / coeff_to_extended / extended_to_coeff); call sites in the reference:
// benches/bench.rs:236-237 (keygen) and :321-329 (create_proof).
//
// Structure: log_n radix-2 Stockham stages are grouped into passes of S <= 9 stages.  A
// Stockham stage with length l and stride s (l*s = n) maps
//     a = x[q + s*p], b = x[q + s*(p + l/2)]  ->  y[q + s*2p] = a + b,  y[q + s*(2p+1)] = (a - b) * w_l^p.
// S consecutive stages only mix the R = 2^S elements x[c + (n/R)*r], r < R, of one "column"
// c = q + s*p', so one CTA stages C adjacent columns (C*32 B contiguous per row, 128-bit
// loads) in shared memory, runs the S stages there and writes every element once:
// out index q + s*(R*p' + bitrev_S(r)).  No bit-reversal pass, natural order in and out.
// HBM traffic = 2 * n * 32 B per pass (+ the twiddle table, L2 resident for n <= 2^20).
// Scaling (1/n) and the coset ZETA-power twists are fused into the first load / last store.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "ctx.hpp"

namespace b2r {

// ---- twiddle table: tw[e] = omega^e, e < n/2 ------------------------------------------
__global__ void k_twiddles(fe_t* tw, fe_t omega, uint32_t count) {
    const uint32_t CH = 64;
    uint32_t chunk = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t e0 = chunk * CH;
    if (e0 >= count) return;
    uint32_t ew[8] = {e0, 0, 0, 0, 0, 0, 0, 0};
    fe_t cur = Fr::pow(omega, ew);
    for (uint32_t j = 0; j < CH && e0 + j < count; j++) {
        tw[e0 + j] = cur;
        cur = Fr::mul(cur, omega);
    }
}

struct NttPassArgs {
    const fe_t* x;
    fe_t* y;
    const fe_t* tw;       // omega^e, e < n/2
    uint64_t in_stride;   // elements between consecutive vectors of the batch (input)
    uint64_t in_stride2;  // two-level input addressing: vector v starts at (v / in_inner) * in_stride2 + (v % in_inner) * in_stride
    uint32_t in_inner;    // (0 = flat: v * in_stride)
    uint64_t out_stride;  // same for output
    uint32_t in_len;      // valid input elements per vector (rest read as zero)
    uint32_t log_n;
    uint32_t log_s;  // stride s = 2^log_s = product of the radices of earlier passes
    uint32_t log_c;  // columns per CTA
    uint32_t pre;    // multiply input element i by pre3[i % 3]
    uint32_t post;   // multiply output element i by post3[i % 3]
    uint32_t vec_fast;  // 0: grid (tiles, vectors); else 1-D grid, vector = blockIdx.x % vec_fast, tile = blockIdx.x / vec_fast
    fe_t pre3[3];
    fe_t post3[3];
};

__device__ __forceinline__ fe_t ld_fe(const fe_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    fe_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ fe_t ld_fe_nc(const fe_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    fe_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_fe(fe_t* p, const fe_t& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

// Shared-memory tile accessors.  A tile holds R rows x C columns of field elements, element (r, c) at index r * C + c.
//  SplitTile    the tile split into the low and the high 16 bytes of every element: a warp's 128-bit accesses then touch
//               consecutive banks (32-byte elements accessed whole are a 2-way bank conflict on every load and store)
//  SwizzledTile the image a TMA tensor copy with CU_TENSOR_MAP_SWIZZLE_128B leaves: 128-byte lines of 4 elements, the
//               16-byte chunk j of line L stored at chunk j ^ (L & 7) - also conflict free for 32 consecutive elements
struct SplitTile {
    uint4 *lo, *hi;
    __device__ __forceinline__ fe_t ld(uint32_t idx) const {
        const uint4 a = lo[idx], b = hi[idx];
        fe_t r;
        r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
        r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
        return r;
    }
    __device__ __forceinline__ void st(uint32_t idx, const fe_t& v) const {
        lo[idx] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
        hi[idx] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
    }
};
struct SwizzledTile {
    uint4* base;  // 1024-byte aligned
    __device__ __forceinline__ uint32_t chunk(uint32_t idx) const {
        const uint32_t line = idx >> 2;
        return (line << 3) + (((idx & 3u) << 1) ^ (line & 7u));
    }
    __device__ __forceinline__ fe_t ld(uint32_t idx) const {
        const uint32_t c = chunk(idx);
        const uint4 a = base[c], b = base[c ^ 1u];
        fe_t r;
        r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
        r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
        return r;
    }
    __device__ __forceinline__ void st(uint32_t idx, const fe_t& v) const {
        const uint32_t c = chunk(idx);
        base[c] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
        base[c ^ 1u] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
    }
};

// ---- S radix-2 stages on a tile in shared memory (decimation in frequency, in place), two stages per round trip: a
// thread takes rows r0, r0+q, r0+2q, r0+3q (q = a quarter of the current block), does the two butterflies of stage i and
// the two of stage i+1 in registers (4 twiddle products, as two radix-2 stages would) and writes the 4 rows back:
// half the shared-memory traffic and barriers of the plain radix-2 loop.  Ends with a barrier.
// TWS: the pass's twiddles were staged in shared memory by ntt_stage_twiddles (passes behind the first one, where every
// column of the CTA shares its twiddles): `tws` holds, round after round, w^ea / w^eb for rp < 2q and then w^ec for rp < q.
template <int S, class Tile, bool TWS = false>
__device__ __forceinline__ void ntt_tile_stages(const NttPassArgs& A, const Tile& tile, uint32_t c0, const fe_t* tws = nullptr) {
    constexpr uint32_t R = 1u << S;
    const uint32_t T = blockDim.x, tid = threadIdx.x;
    const uint32_t log_c = A.log_c, cmask = (1u << log_c) - 1;
    const uint32_t log_cols = A.log_n - S;  // columns = n / R
    uint32_t i = 0, tw_off = 0;
#pragma unroll 1
    for (; i + 1 < S; i += 2) {
        const uint32_t log_q = S - 2 - i, q = 1u << log_q;
        const fe_t* tw_round = TWS ? tws + tw_off : nullptr;
        tw_off += 3u << log_q;
        for (uint32_t g = tid; g < ((R / 4) << log_c); g += T) {
            const uint32_t c = g & cmask, gf = g >> log_c;
            const uint32_t blk = gf >> log_q, rp = gf & (q - 1);
            const uint32_t i0 = ((((blk << (log_q + 2)) + rp)) << log_c) + c, st = q << log_c;
            const uint32_t base = ((c0 + c) >> A.log_s) << A.log_s;  // s * p'
            const uint32_t ea = (base + (rp << log_cols)) << i;          // stage i, rows (r0, r0 + 2q)
            const uint32_t eb = (base + ((rp + q) << log_cols)) << i;    // stage i, rows (r0 + q, r0 + 3q)
            const uint32_t ec = ea << 1;                                 // stage i + 1, both pairs
            const fe_t a0 = tile.ld(i0), a1 = tile.ld(i0 + st), a2 = tile.ld(i0 + 2 * st), a3 = tile.ld(i0 + 3 * st);
            const fe_t u0 = Fr::add(a0, a2), u1 = Fr::add(a1, a3);
            fe_t d0 = Fr::sub(a0, a2), d1 = Fr::sub(a1, a3);
            if (ea) d0 = Fr::mul(d0, TWS ? ld_fe(tw_round + rp) : ld_fe_nc(A.tw + ea));
            // zero-padded first pass (coeff_to_extended: 3/4 of the rows are zero): a1 = a3 = 0 in the first round, 0 * w = 0
            if (!Fr::is_zero(d1)) d1 = Fr::mul(d1, TWS ? ld_fe(tw_round + rp + q) : ld_fe_nc(A.tw + eb));
            const fe_t wc = TWS ? ld_fe(tw_round + 2 * q + rp) : ld_fe_nc(A.tw + ec);
            tile.st(i0, Fr::add(u0, u1));
            tile.st(i0 + 2 * st, Fr::add(d0, d1));
            fe_t v1 = Fr::sub(u0, u1), v3 = Fr::sub(d0, d1);
            if (ec) {
                v1 = Fr::mul(v1, wc);
                v3 = Fr::mul(v3, wc);
            }
            tile.st(i0 + st, v1);
            tile.st(i0 + 3 * st, v3);
        }
        __syncthreads();
    }
    if (i < S) {  // odd S: one radix-2 stage left (half = 1)
        for (uint32_t b = tid; b < ((R / 2) << log_c); b += T) {
            const uint32_t c = b & cmask, bf = b >> log_c;
            const uint32_t i0 = ((bf << 1) << log_c) + c, i1 = i0 + (1u << log_c);
            const uint32_t e = (((c0 + c) >> A.log_s) << A.log_s) << i;
            const fe_t a = tile.ld(i0), bb = tile.ld(i1);
            fe_t d = Fr::sub(a, bb);
            if (e) d = Fr::mul(d, ld_fe_nc(A.tw + e));
            tile.st(i0, Fr::add(a, bb));
            tile.st(i1, d);
        }
        __syncthreads();
    }
}

// ---- store: local row r holds output digit t = bitrev_S(r)
template <int S, class Tile>
__device__ __forceinline__ void ntt_tile_store(const NttPassArgs& A, const Tile& tile, uint32_t c0, fe_t* y) {
    constexpr uint32_t R = 1u << S;
    const uint32_t T = blockDim.x, tid = threadIdx.x;
    const uint32_t log_c = A.log_c, cmask = (1u << log_c) - 1;
    const uint32_t smask = (1u << A.log_s) - 1;
    if (A.log_s == 0) {
        // first pass: out index = R*col + t, make t the fastest index for contiguous stores
        for (uint32_t idx = tid; idx < (R << log_c); idx += T) {
            uint32_t t = idx & (R - 1), c = idx >> S;
            uint32_t r = __brev(t) >> (32 - S);
            fe_t v = tile.ld((r << log_c) + c);
            uint32_t oi = ((c0 + c) << S) + t;
            if (A.post) v = Fr::mul(v, A.post3[oi % 3u]);
            st_fe(y + oi, v);
        }
    } else {
        for (uint32_t idx = tid; idx < (R << log_c); idx += T) {
            uint32_t r = idx >> log_c, c = idx & cmask;
            uint32_t t = __brev(r) >> (32 - S);
            uint32_t col = c0 + c, q = col & smask, pp = col >> A.log_s;
            uint32_t oi = q + (pp << (A.log_s + S)) + (t << A.log_s);
            fe_t v = tile.ld(idx);
            if (A.post) v = Fr::mul(v, A.post3[oi % 3u]);
            st_fe(y + oi, v);
        }
    }
}

// Twiddles of one pass behind the first (log_c <= log_s: all columns of the CTA share base = s * p'): about R values,
// fetched once per CTA next to the tile load instead of three 32-byte L2 / L1 reads per radix-4 unit and round.
template <int S>
__device__ __forceinline__ void ntt_stage_twiddles(const NttPassArgs& A, uint32_t c0, fe_t* tws) {
    const uint32_t log_cols = A.log_n - S;
    const uint32_t base = (c0 >> A.log_s) << A.log_s;
    uint32_t off = 0;
    for (uint32_t i = 0; i + 1 < (uint32_t)S; i += 2) {
        const uint32_t q = 1u << (S - 2 - i);
        for (uint32_t j = threadIdx.x; j < 3 * q; j += blockDim.x) {
            const uint32_t e = j < 2 * q ? (base + (j << log_cols)) << i : ((base + ((j - 2 * q) << log_cols)) << i) << 1;
            st_fe(tws + off + j, ld_fe_nc(A.tw + e));
        }
        off += 3 * q;
    }
}
template <int S>
constexpr uint32_t ntt_staged_twiddles() {   // entries ntt_stage_twiddles writes
    uint32_t n = 0;
    for (int i = 0; i + 1 < S; i += 2) n += 3u << (S - 2 - i);
    return n;
}

template <int S, bool TWS>
__global__ void __launch_bounds__(256, 4) k_ntt_pass(const NttPassArgs A) {
    constexpr uint32_t R = 1u << S;
    extern __shared__ uint4 smem_raw[];
    SplitTile tile;
    tile.lo = smem_raw;
    tile.hi = smem_raw + ((size_t)(1u << S) << A.log_c);
    fe_t* tws = reinterpret_cast<fe_t*>(smem_raw + ((size_t)(2u << S) << A.log_c));
    const uint32_t T = blockDim.x, tid = threadIdx.x;
    const uint32_t log_c = A.log_c, cmask = (1u << log_c) - 1;
    const uint32_t log_cols = A.log_n - S;  // columns = n / R
    // vec_fast: CTAs that are resident together work on the SAME column tile of different vectors, i.e. on the same twiddles
    const uint32_t vec = A.vec_fast ? blockIdx.x % A.vec_fast : blockIdx.y;
    const uint32_t c0 = (A.vec_fast ? blockIdx.x / A.vec_fast : blockIdx.x) << log_c;
    const fe_t* x = A.in_inner ? A.x + (uint64_t)(vec / A.in_inner) * A.in_stride2 + (uint64_t)(vec % A.in_inner) * A.in_stride
                               : A.x + (uint64_t)vec * A.in_stride;
    fe_t* y = A.y + (uint64_t)vec * A.out_stride;

    // ---- load R rows x C columns (row r of column c lives at c + (n/R)*r)
    for (uint32_t idx = tid; idx < (R << log_c); idx += T) {
        uint32_t r = idx >> log_c, c = idx & cmask;
        uint32_t gi = c0 + c + (r << log_cols);
        fe_t v;
        if (gi < A.in_len) {
            v = ld_fe(x + gi);
            if (A.pre) {
                uint32_t m3 = gi % 3u;
                if (m3) v = Fr::mul(v, A.pre3[m3]);
            }
        } else {
            v = Fr::zero();
        }
        tile.st(idx, v);
    }
    if (TWS) ntt_stage_twiddles<S>(A, c0, tws);
    __syncthreads();
    ntt_tile_stages<S, SplitTile, TWS>(A, tile, c0, tws);
    ntt_tile_store<S>(A, tile, c0, y);
}

// ---- the same pass with the tile loads done by the TMA engine ----------------------------------------------------------
// Persistent CTAs (grid = resident CTAs) walk the (vector, column tile) list; the R x C tile of the NEXT work item is
// fetched by ONE cp.async.bulk.tensor (a 5-D tensor map over [outer vector][inner vector][row][128-byte line][word],
// 128-byte swizzle) into the other half of a two-stage shared-memory ring and completes on that stage's mbarrier while
// the 256 threads run the butterflies of the current tile: the compute warps issue no global loads for the tile and
// never wait for them.  Rows beyond in_len (the zero padding of coeff_to_extended) lie outside the tensor and are
// zero-filled by the copy engine without touching memory.
struct NttTmaArgs {
    uint32_t tiles_per_vec;  // n / R / C
    uint32_t total_tiles;    // tiles_per_vec * batch
    uint32_t inner;          // vectors per outer group (tensor dims 3 / 4)
    uint32_t tile_bytes;
};
__device__ __forceinline__ uint32_t ntt_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ntt_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
template <int S>
__global__ void __launch_bounds__(256, 3) k_ntt_pass_tma(const NttPassArgs A, const NttTmaArgs M, const __grid_constant__ CUtensorMap tmap) {
    constexpr uint32_t R = 1u << S;
    extern __shared__ uint4 smem_raw[];
    __shared__ __align__(8) uint64_t bars[2];
    // the swizzle pattern is a function of the shared-memory address: align the ring to 1024 bytes
    uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t tid = threadIdx.x, log_c = A.log_c, cmask = (1u << log_c) - 1;
    const uint32_t log_cols = A.log_n - S;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ntt_smem_addr(&bars[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ntt_smem_addr(&bars[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    }
    __syncthreads();
    auto issue = [&](uint32_t work, uint32_t stage) {   // thread 0 only
        const uint32_t vec = work / M.tiles_per_vec, tilei = work % M.tiles_per_vec;
        const uint32_t bar = ntt_smem_addr(&bars[stage]), dst = ntt_smem_addr(ring + (size_t)stage * M.tile_bytes);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic-proxy accesses of this stage before the async write
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(M.tile_bytes) : "memory");
        // coordinates: word in line, line = first column of the tile / 4, row, inner vector, outer vector
        asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
                     "l"(&tmap), "r"(bar), "r"(0), "r"((int)((tilei << log_c) >> 2)), "r"(0), "r"((int)(vec % M.inner)), "r"((int)(vec / M.inner))
                     : "memory");
    };
    uint32_t work = blockIdx.x, k = 0;
    if (tid == 0 && work < M.total_tiles) issue(work, 0);
    for (; work < M.total_tiles; work += gridDim.x, k++) {
        const uint32_t stage = k & 1u;
        const uint32_t next = work + gridDim.x;
        if (tid == 0 && next < M.total_tiles) issue(next, stage ^ 1u);
        ntt_mbar_wait(ntt_smem_addr(&bars[stage]), (k >> 1) & 1u);
        SwizzledTile tile;
        tile.base = reinterpret_cast<uint4*>(ring + (size_t)stage * M.tile_bytes);
        const uint32_t vec = work / M.tiles_per_vec, c0 = (work % M.tiles_per_vec) << log_c;
        fe_t* y = A.y + (uint64_t)vec * A.out_stride;
        if (A.pre) {   // coset twist of the coefficients: element gi times pre3[gi % 3] (the rows past in_len are zero)
            const uint32_t live = min((uint32_t)R, (A.in_len + (1u << log_cols) - 1) >> log_cols) << log_c;
            for (uint32_t idx = tid; idx < live; idx += blockDim.x) {
                const uint32_t r = idx >> log_c, c = idx & cmask;
                const uint32_t m3 = (c0 + c + (r << log_cols)) % 3u;
                if (m3) tile.st(idx, Fr::mul(tile.ld(idx), A.pre3[m3]));
            }
            __syncthreads();
        }
        ntt_tile_stages<S>(A, tile, c0);
        ntt_tile_store<S>(A, tile, c0, y);
        __syncthreads();   // every thread is done with this stage before thread 0 hands it back to the copy engine
    }
}

typedef void (*ntt_kernel_t)(const NttPassArgs);
typedef void (*ntt_tma_kernel_t)(const NttPassArgs, const NttTmaArgs, const CUtensorMap);
static ntt_tma_kernel_t pass_kernel_tma(int S) {
    switch (S) {
        case 4: return k_ntt_pass_tma<4>;
        case 5: return k_ntt_pass_tma<5>;
        case 6: return k_ntt_pass_tma<6>;
        case 7: return k_ntt_pass_tma<7>;
        case 8: return k_ntt_pass_tma<8>;
        default: return nullptr;
    }
}
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn get_encode_tiled() {
    static encode_tiled_fn fn = []() -> encode_tiled_fn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
        return (encode_tiled_fn)p;
    }();
    return fn;
}
static ntt_kernel_t pass_kernel(int S, bool tws) {
    switch (S) {
        case 1: return k_ntt_pass<1, false>;
        case 2: return tws ? k_ntt_pass<2, true> : k_ntt_pass<2, false>;
        case 3: return tws ? k_ntt_pass<3, true> : k_ntt_pass<3, false>;
        case 4: return tws ? k_ntt_pass<4, true> : k_ntt_pass<4, false>;
        case 5: return tws ? k_ntt_pass<5, true> : k_ntt_pass<5, false>;
        case 6: return tws ? k_ntt_pass<6, true> : k_ntt_pass<6, false>;
        case 7: return tws ? k_ntt_pass<7, true> : k_ntt_pass<7, false>;
        case 8: return tws ? k_ntt_pass<8, true> : k_ntt_pass<8, false>;
        case 9: return tws ? k_ntt_pass<9, true> : k_ntt_pass<9, false>;
        default: return nullptr;
    }
}
static uint32_t staged_twiddles(int S) {
    uint32_t n = 0;
    for (int i = 0; i + 1 < S; i += 2) n += 3u << (S - 2 - i);
    return n;
}

int32_t ntt_get_twiddles(b2r_ctx* ctx, const fe_t& omega, uint32_t log_n, const fe_t** out) {
    TwiddleKey key;
    key.log_n = log_n;
    for (int i = 0; i < 4; i++) key.w[i] = (uint64_t)omega.l[2 * i] | ((uint64_t)omega.l[2 * i + 1] << 32);
    auto it = ctx->twiddles.find(key);
    if (it != ctx->twiddles.end()) {
        *out = it->second;
        return 0;
    }
    uint32_t count = log_n == 0 ? 1 : (1u << (log_n - 1));
    fe_t* d = nullptr;
    B2R_CUDA(ctx, cudaMalloc(&d, (size_t)count * sizeof(fe_t)));
    uint32_t chunks = (count + 63) / 64;
    k_twiddles<<<(chunks + 127) / 128, 128, 0, ctx->stream>>>(d, omega, count);
    B2R_LAUNCH_CHECK(ctx);
    ctx->twiddles[key] = d;
    *out = d;
    return 0;
}

// domain constants (host): omega_k = ROOT_OF_UNITY^(2^(28-k))
static fe_t root_of_unity() {
    // 0x03ddb9f5166d18b798865ea93dd31f743215cf6dd39329c8d34f1ed960c37c9c, canonical
    fe_t c;
    const uint32_t w[8] = {0x60c37c9cu, 0xd34f1ed9u, 0xd39329c8u, 0x3215cf6du,
                           0x3dd31f74u, 0x98865ea9u, 0x166d18b7u, 0x03ddb9f5u};
    for (int i = 0; i < 8; i++) c.l[i] = w[i];
    return Fr::to_mont(c);
}
fe_t fr_omega(uint32_t k) {
    fe_t w = root_of_unity();
    for (uint32_t i = k; i < 28; i++) w = Fr::sqr(w);
    return w;
}
fe_t fr_zeta() {
    // 0x30644e72e131a029048b6e193fd84104cc37a73fec2bc5e9b8ca0b2d36636f23, canonical
    fe_t c;
    const uint32_t w[8] = {0x36636f23u, 0xb8ca0b2du, 0xec2bc5e9u, 0xcc37a73fu,
                           0x3fd84104u, 0x048b6e19u, 0xe131a029u, 0x30644e72u};
    for (int i = 0; i < 8; i++) c.l[i] = w[i];
    return Fr::to_mont(c);
}
fe_t fr_from_u64(uint64_t v) {
    fe_t c = Fr::zero();
    c.l[0] = (uint32_t)v;
    c.l[1] = (uint32_t)(v >> 32);
    return Fr::to_mont(c);
}

enum NttMode { MODE_PLAIN = 0, MODE_INV = 1, MODE_COSET_FWD = 2, MODE_COSET_INV = 3 };

// Runs the passes.  in: `batch` vectors of in_len valid elements (stride in_stride);
// out: 2^log_n elements each (stride out_stride).  `in` may equal `out` (in place).
static int32_t ntt_run(b2r_ctx* ctx, const fe_t* in, uint64_t in_stride, uint32_t in_len, fe_t* out,
                       uint64_t out_stride, size_t batch, const fe_t& omega, uint32_t log_n, int mode, uint32_t in_inner = 0,
                       uint64_t in_stride2 = 0) {
    if (log_n > 27) return fail(ctx, B2R_ERR_INVALID, "ntt: log_n > 27");
    if (batch == 0) return 0;
    if (batch > 65535) return fail(ctx, B2R_ERR_INVALID, "ntt: batch > 65535");
    const uint64_t n = 1ull << log_n;
    if (log_n == 0) {
        if (in != out) B2R_CUDA(ctx, cudaMemcpy2DAsync(out, out_stride * 32, in, in_stride * 32, 32, batch, cudaMemcpyDeviceToDevice, ctx->stream));
        return 0;
    }
    const fe_t* tw = nullptr;
    B2R_TRY(ntt_get_twiddles(ctx, omega, log_n, &tw));

    // stages per pass: 9 (64 KiB tiles, fewest passes) for batches; a single short transform is a latency chain of a few
    // hundred CTAs, where three passes of 32 KiB tiles at 4 CTAs/SM finish sooner than two of 64 KiB (2^17: 57 vs 72 us)
    int max_s = (batch << log_n) <= ((size_t)1 << 18) ? 8 : 9;
    if (const char* ov = getenv("B2R_NTT_MAXS")) max_s = atoi(ov) >= 4 && atoi(ov) <= 9 ? atoi(ov) : 9;   // tuning hook: stages per pass
    int npass = (log_n + max_s - 1) / max_s;
    int S[8];
    {
        int base = log_n / npass, extra = log_n % npass;
        for (int p = 0; p < npass; p++) S[p] = base + (p < extra ? 1 : 0);
    }
    // Stockham passes are out of place.  Destinations alternate so that the LAST pass writes `out` and no pass writes
    // the buffer it reads: out of place (in != out) alternates out / ping backwards from the end; in place uses
    // ping (and pong when the pass count is odd).  No copy-back pass (the first version always started in the scratch
    // buffer and copied 16 MiB per 2^19 transform back: 22 GB per 64-proof step).
    fe_t *ping = nullptr, *pong = nullptr;
    const bool in_place = (const void*)in == (const void*)out;
    if (npass > 1 || in_place) B2R_TRY(scratch_get(ctx, SC_NTT_PING, batch * n * sizeof(fe_t), (void**)&ping));
    if (in_place && npass > 1 && (npass & 1)) B2R_TRY(scratch_get(ctx, SC_NTT_PONG, batch * n * sizeof(fe_t), (void**)&pong));
    fe_t n_inv = Fr::zero();
    fe_t zeta = fr_zeta(), zeta2 = Fr::sqr(zeta);
    if (mode == MODE_INV || mode == MODE_COSET_INV) n_inv = Fr::inv(fr_from_u64(n));

    const fe_t* src = in;
    uint64_t src_stride = in_stride;
    uint32_t src_len = in_len;
    uint32_t log_s = 0;
    for (int p = 0; p < npass; p++) {
        bool last = (p == npass - 1);
        fe_t* dst;
        if (!in_place) {
            dst = ((npass - 1 - p) & 1) ? ping : out;
        } else if (npass == 1) {
            dst = ping;  // single pass in place: through the scratch buffer, copied back below
        } else if (npass & 1) {
            dst = last ? out : (p & 1 ? pong : ping);  // in -> ping -> pong -> ... -> out
        } else {
            dst = (p & 1) ? out : ping;                // in -> ping -> out -> ping -> out
        }
        const uint64_t dst_stride = (dst == out) ? out_stride : n;
        NttPassArgs A;
        A.x = src;
        A.y = dst;
        A.tw = tw;
        A.in_stride = src_stride;
        A.in_inner = p == 0 ? in_inner : 0;   // only the first pass reads the caller's layout
        A.in_stride2 = in_stride2;
        A.out_stride = dst_stride;
        A.in_len = src_len;
        A.log_n = log_n;
        A.log_s = log_s;
        A.pre = 0;
        A.vec_fast = 0;
        A.post = 0;
        for (int j = 0; j < 3; j++) A.pre3[j] = A.post3[j] = Fr::one();
        if (p == 0 && mode == MODE_COSET_FWD) {
            A.pre = 1;
            A.pre3[1] = zeta;
            A.pre3[2] = zeta2;
        }
        if (last && mode == MODE_INV) {
            A.post = 1;
            A.post3[0] = A.post3[1] = A.post3[2] = n_inv;
        }
        if (last && mode == MODE_COSET_INV) {
            A.post = 1;
            A.post3[0] = n_inv;
            A.post3[1] = Fr::mul(n_inv, zeta2);
            A.post3[2] = Fr::mul(n_inv, zeta);
        }
        uint32_t log_cols = log_n - S[p];
        // columns per CTA: 256 threads want >= 256 four-row groups (2^(S-2) per column); measured best on B200
        // (tools/microbench.py, B2R_NTT_LOGC sweep)
        uint32_t log_c = (S[p] >= 8) ? 2 : (S[p] == 7 ? 3 : 4);
        {   // tuning hook (tools/microbench.py): columns per CTA
            const char* ov = getenv("B2R_NTT_LOGC");
            if (ov) log_c = (uint32_t)atoi(ov);
            if ((((size_t)sizeof(fe_t) << S[p]) << log_c) > 200 * 1024) log_c = 2;
        }
        if (log_c > log_cols) log_c = log_cols;
        if (p > 0 && log_c > log_s) log_c = log_s;
        // CTA size: 128 threads on half the columns (up to 8 CTAs per SM instead of 4 x 256 threads).  Same registers, same
        // shared memory per SM, same work per thread - but a barrier now waits for 4 warps instead of 8 and twice as many CTAs
        // are in different phases (load / butterflies / store) at any time: 64 x 2^17 1.323 -> 1.254 ms, 64 x 2^19 5.644 ->
        // 5.378 ms (tools/microbench.py; quarter-size CTAs: 1.325 / 5.346).  B2R_NTT_HALF = 0 / 1 / 2 is the tuning hook.
        unsigned threads = 256;
        {
            const char* ov = getenv("B2R_NTT_HALF");
            const uint32_t h = ov ? (uint32_t)atoi(ov) : 1u;
            if (h >= 1 && h <= 2 && log_c >= h + 1) { threads = 256u >> h; log_c -= h; }
        }
        A.log_c = log_c;
        size_t smem = ((size_t)sizeof(fe_t) << S[p]) << log_c;
        // passes behind the first: the CTA's columns share their twiddles (log_c <= log_s) and can stage them in shared memory
        // once.  Opt-in ("1"): measured 3 % SLOWER on B200 (64 x 2^19: 5.79 vs 5.64 ms) - at the 64-register cap of 4 CTAs/SM
        // the extra addressing spills (92 vs 36 bytes), and 32 warps per SM already hide the L1-resident twiddle reads
        bool tws = false;
        if (const char* ov = getenv("B2R_NTT_TWS")) tws = ov[0] == '1' && p > 0 && S[p] >= 2 && log_c <= log_s;
        const size_t tile_smem = smem;
        if (tws) smem += (size_t)staged_twiddles(S[p]) * sizeof(fe_t);
        ntt_kernel_t kern = pass_kernel(S[p], tws);
        if (smem > 48 * 1024) B2R_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // TMA path: tiles of at most 32 KiB (two ring stages, 3 CTAs per SM), at least one 128-byte line per row, rows <= 256
        bool use_tma = false;
        {
            // opt-in ("1"): measured on B200 the plain-load kernel is faster (4 CTAs/SM already hide the tile loads behind
            // other CTAs' butterflies; the two-stage ring costs a CTA of occupancy) - profiles/r02_tma_experiments.md
            const char* ov = getenv("B2R_NTT_TMA");
            const bool want = ov && ov[0] == '1';
            const uint32_t ncols = 1u << log_cols;
            const uint32_t vec_rows = (uint32_t)((src_len + ncols - 1) / ncols);   // rows that exist in memory (rest: zero fill)
            use_tma = want && get_encode_tiled() && pass_kernel_tma(S[p]) && log_c >= 2 && tile_smem <= 32 * 1024 && vec_rows >= 1 &&
                      (src_len % ncols == 0) && (src_stride % 4 == 0) && ((uintptr_t)src % 16 == 0);
            if (use_tma) {
                const uint32_t inner = A.in_inner ? A.in_inner : (uint32_t)batch, outer = A.in_inner ? (uint32_t)(batch / A.in_inner) : 1u;
                const uint64_t inner_stride = src_stride, outer_stride = A.in_inner ? A.in_stride2 : src_stride * batch;
                CUtensorMap tmap;
                const cuuint64_t gdim[5] = {32, ncols / 4, vec_rows, inner, outer};
                const cuuint64_t gstr[4] = {128, (cuuint64_t)ncols * 32, inner_stride * 32, (outer > 1 ? outer_stride : inner_stride * inner) * 32};
                const cuuint32_t box[5] = {32, 1u << (log_c - 2), 1u << S[p], 1, 1};
                const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
                CUresult cr = get_encode_tiled()(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 5, (void*)src, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (cr != CUDA_SUCCESS) {
                    use_tma = false;   // shapes the encoder refuses go through the plain kernel
                } else {
                    NttTmaArgs M;
                    M.tiles_per_vec = 1u << (log_cols - log_c);
                    M.total_tiles = M.tiles_per_vec * (uint32_t)batch;
                    M.inner = inner;
                    M.tile_bytes = (uint32_t)tile_smem;
                    const size_t dyn = 2 * tile_smem + 1024;
                    ntt_tma_kernel_t tk = pass_kernel_tma(S[p]);
                    B2R_CUDA(ctx, cudaFuncSetAttribute(tk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
                    const uint32_t resident = 3u * (uint32_t)ctx->sm_count;
                    const uint32_t gridx = M.total_tiles < resident ? M.total_tiles : resident;
                    { KTimer kt(ctx, "ntt_pass", (double)batch * n);
                    tk<<<gridx, 256, dyn, ctx->stream>>>(A, M, tmap); }
                    B2R_LAUNCH_CHECK(ctx);
                }
            }
        }
        if (!use_tma) {
            dim3 grid(1u << (log_cols - log_c), (unsigned)batch);
            A.vec_fast = 0;
            {   // tuning hook: B2R_NTT_ORDER=1 walks the batch first (twiddles shared by the resident CTAs); =2 only in the first pass
                const char* ov = getenv("B2R_NTT_ORDER");
                const int o = ov ? atoi(ov) : 0;
                if (batch > 1 && (o == 1 || (o == 2 && p == 0) || (o == 3 && p > 0)) && ((uint64_t)batch << (log_cols - log_c)) < (1ull << 31)) {
                    A.vec_fast = (uint32_t)batch;
                    grid = dim3((unsigned)(batch << (log_cols - log_c)), 1);
                }
            }
            { KTimer kt(ctx, "ntt_pass", (double)batch * n);
            kern<<<grid, threads, smem, ctx->stream>>>(A); }
            B2R_LAUNCH_CHECK(ctx);
        }
        if (last && dst != out) {
            B2R_CUDA(ctx, cudaMemcpy2DAsync(out, out_stride * 32, dst, dst_stride * 32, n * 32, batch, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        src = dst;
        src_stride = dst_stride;
        src_len = (uint32_t)n;
        log_s += S[p];
    }
    return 0;
}

static inline fe_t fe_from_abi(const b2r_fr* p) {
    fe_t r;
    for (int i = 0; i < 4; i++) {
        r.l[2 * i] = (uint32_t)p->l[i];
        r.l[2 * i + 1] = (uint32_t)(p->l[i] >> 32);
    }
    return r;
}

// host-buffer wrapper: stage through the SC_STAGE arena
static int32_t ntt_host(b2r_ctx* ctx, const b2r_fr* in, uint32_t in_len, b2r_fr* out, const fe_t& omega,
                        uint32_t log_n, int mode) {
    size_t n = (size_t)1 << log_n;
    fe_t* d = nullptr;
    B2R_TRY(scratch_get(ctx, SC_STAGE, n * sizeof(fe_t), (void**)&d));
    B2R_CUDA(ctx, cudaMemcpyAsync(d, in, (size_t)in_len * 32, cudaMemcpyHostToDevice, ctx->stream));
    B2R_TRY(ntt_run(ctx, d, n, in_len, d, n, 1, omega, log_n, mode));
    B2R_CUDA(ctx, cudaMemcpyAsync(out, d, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    B2R_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// coeff_to_extended for `outer` x `inner` vectors in one batched launch per pass: vector (o, i) starts at
// coeffs + o * outer_stride + i * 2^k (elements); outputs are contiguous, (o * inner + i) * 2^ext_k
int32_t coset_ntt_grouped_dev(b2r_ctx* ctx, const fe_t* coeffs, size_t outer, uint64_t outer_stride, size_t inner, uint32_t k, uint32_t ext_k,
                              fe_t* out) {
    return ntt_run(ctx, coeffs, 1ull << k, 1u << k, out, 1ull << ext_k, outer * inner, fr_omega(ext_k), ext_k, MODE_COSET_FWD, (uint32_t)inner,
                   outer_stride);
}

}  // namespace b2r

using namespace b2r;

extern "C" {

int32_t b2r_ntt_fr(b2r_ctx* ctx, b2r_fr* a, const b2r_fr* omega, uint32_t log_n) try {
    B2R_ENTER(ctx);
    if (!a || !omega) return fail(ctx, B2R_ERR_INVALID, "ntt: null pointer");
    if (log_n > 27) return fail(ctx, B2R_ERR_INVALID, "ntt: log_n > 27");
    return ntt_host(ctx, a, 1u << log_n, a, fe_from_abi(omega), log_n, MODE_PLAIN);
} B2R_ABI_CATCH(ctx)
int32_t b2r_ntt_fr_dev(b2r_ctx* ctx, b2r_fr* a_dev, const b2r_fr* omega_host, uint32_t log_n) try {
    B2R_ENTER(ctx);
    return b2r_ntt_fr_batch_dev(ctx, a_dev, 1, omega_host, log_n);
} B2R_ABI_CATCH(ctx)
int32_t b2r_ntt_fr_batch_dev(b2r_ctx* ctx, b2r_fr* a_dev, size_t batch, const b2r_fr* omega_host, uint32_t log_n) try {
    B2R_ENTER(ctx);
    if (!a_dev || !omega_host) return fail(ctx, B2R_ERR_INVALID, "ntt: null pointer");
    uint64_t n = 1ull << log_n;
    return ntt_run(ctx, (fe_t*)a_dev, n, (uint32_t)n, (fe_t*)a_dev, n, batch, fe_from_abi(omega_host), log_n, MODE_PLAIN);
} B2R_ABI_CATCH(ctx)
int32_t b2r_intt_fr(b2r_ctx* ctx, b2r_fr* a, uint32_t k) try {
    B2R_ENTER(ctx);
    if (!a) return fail(ctx, B2R_ERR_INVALID, "intt: null pointer");
    if (k > 27) return fail(ctx, B2R_ERR_INVALID, "intt: k > 27");
    return ntt_host(ctx, a, 1u << k, a, Fr::inv(fr_omega(k)), k, MODE_INV);
} B2R_ABI_CATCH(ctx)
int32_t b2r_intt_fr_batch_dev(b2r_ctx* ctx, b2r_fr* a_dev, size_t batch, uint32_t k) try {
    B2R_ENTER(ctx);
    if (!a_dev) return fail(ctx, B2R_ERR_INVALID, "intt: null pointer");
    if (k > 27) return fail(ctx, B2R_ERR_INVALID, "intt: k > 27");
    uint64_t n = 1ull << k;
    return ntt_run(ctx, (fe_t*)a_dev, n, (uint32_t)n, (fe_t*)a_dev, n, batch, Fr::inv(fr_omega(k)), k, MODE_INV);
} B2R_ABI_CATCH(ctx)
int32_t b2r_coset_ntt_fr(b2r_ctx* ctx, const b2r_fr* coeffs, uint32_t k, uint32_t ext_k, b2r_fr* out) try {
    B2R_ENTER(ctx);
    if (!coeffs || !out) return fail(ctx, B2R_ERR_INVALID, "coset_ntt: null pointer");
    if (ext_k > 27 || k > ext_k) return fail(ctx, B2R_ERR_INVALID, "coset_ntt: need k <= ext_k <= 27");
    return ntt_host(ctx, coeffs, 1u << k, out, fr_omega(ext_k), ext_k, MODE_COSET_FWD);
} B2R_ABI_CATCH(ctx)
int32_t b2r_coset_ntt_fr_batch_dev(b2r_ctx* ctx, const b2r_fr* coeffs_dev, size_t batch, uint32_t k, uint32_t ext_k,
                                   b2r_fr* out_dev) try {
    B2R_ENTER(ctx);
    if (!coeffs_dev || !out_dev) return fail(ctx, B2R_ERR_INVALID, "coset_ntt: null pointer");
    if (ext_k > 27 || k > ext_k) return fail(ctx, B2R_ERR_INVALID, "coset_ntt: need k <= ext_k <= 27");
    if ((const void*)coeffs_dev == (void*)out_dev && k != ext_k)
        return fail(ctx, B2R_ERR_INVALID, "coset_ntt: in place needs k == ext_k");
    return ntt_run(ctx, (const fe_t*)coeffs_dev, 1ull << k, 1u << k, (fe_t*)out_dev, 1ull << ext_k, batch,
                   fr_omega(ext_k), ext_k, MODE_COSET_FWD);
} B2R_ABI_CATCH(ctx)
int32_t b2r_coset_intt_fr(b2r_ctx* ctx, b2r_fr* a, uint32_t ext_k) try {
    B2R_ENTER(ctx);
    if (!a) return fail(ctx, B2R_ERR_INVALID, "coset_intt: null pointer");
    if (ext_k > 27) return fail(ctx, B2R_ERR_INVALID, "coset_intt: ext_k > 27");
    return ntt_host(ctx, a, 1u << ext_k, a, Fr::inv(fr_omega(ext_k)), ext_k, MODE_COSET_INV);
} B2R_ABI_CATCH(ctx)
int32_t b2r_coset_intt_fr_batch_dev(b2r_ctx* ctx, b2r_fr* a_dev, size_t batch, uint32_t ext_k) try {
    B2R_ENTER(ctx);
    if (!a_dev) return fail(ctx, B2R_ERR_INVALID, "coset_intt: null pointer");
    if (ext_k > 27) return fail(ctx, B2R_ERR_INVALID, "coset_intt: ext_k > 27");
    uint64_t n = 1ull << ext_k;
    return ntt_run(ctx, (fe_t*)a_dev, n, (uint32_t)n, (fe_t*)a_dev, n, batch, Fr::inv(fr_omega(ext_k)), ext_k,
                   MODE_COSET_INV);
} B2R_ABI_CATCH(ctx)

}  // extern "C"
