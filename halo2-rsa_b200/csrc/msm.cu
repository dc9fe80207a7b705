// Pippenger bucket MSM over BN254 G1 for sm_100a, resident pre-processed bases.
//
// Replaces halo2_proofs::arithmetic::best_multiexp as reached through
// ParamsKZG::{commit, commit_lagrange} (reference call sites benches/bench.rs:235-237 and
// :321-329).  The reference re-derives window sums from the raw bases on every call; the
// SRS never changes, so here registration stores T[j][i] = 2^(c*j) * P_i (affine) once
// and every signed c-bit digit of every scalar lands in ONE bucket set of 2^(c-1)
// buckets: no per-window reduction and no window-combining doublings at MSM time.
//
// Per call (m scalar vectors sharing the base set, grid.y = vector):
//   1. k_count    digits of each scalar (Montgomery -> canonical -> signed windows),
//                 per-bucket histogram (warp-aggregated atomics: advice columns are
//                 dominated by a few values)
//   2. k_scan     exclusive scan of the histogram -> bucket offsets
//   3. k_scatter  counting sort of (table index, sign) entries by bucket
//   4. k_accum_entries / k_accum_slots   load-balanced segmented sum: every thread adds a
//                 fixed-length chunk of the sorted entry list (XYZZ mixed adds, 64 B
//                 gathers from the table), complete buckets are stored, chunk-boundary
//                 partial sums go to a slot list that the next level reduces the same
//                 way.  Work per thread is uniform whatever the scalar distribution.
//   5. k_br_level1/2/3   sum_b (b+1) * B_b by three levels of per-thread running sums, one
//                 inversion to return the canonical affine point.
// Algorithmic HBM bytes: 32 B per scalar + 64 B per gathered table point.
#include <cuda_runtime.h>

#include <cstdlib>
#include <vector>

#include "ctx.hpp"
#include "ec.cuh"

struct b2r_bases {
    size_t n = 0;
    uint32_t c = 0;
    uint32_t W = 0;
    b2r::affine_t* table = nullptr;  // W * n
};

namespace b2r {

// entries of the bucket-sorted list per thread of k_accum_entries (64: half the chunk-boundary partial sums of 32)
#ifndef MSM_L1
#define MSM_L1 64
#endif
static constexpr uint32_t SLOT_INVALID = 0xffffffffu;
static constexpr uint32_t SLOT_BEGINS = 0x80000000u;
static constexpr uint32_t SLOT_ENDS = 0x40000000u;
static constexpr uint32_t SLOT_KEY = 0x3fffffffu;

__device__ __forceinline__ fe_t ldg_fe(const fe_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    fe_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void stg_fe(fe_t* p, const fe_t& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ affine_t ldg_affine(const affine_t* p) {
    affine_t r;
    r.x = ldg_fe(&p->x);
    r.y = ldg_fe(&p->y);
    return r;
}
__device__ __forceinline__ xyzz_t ld_xyzz(const xyzz_t* p) {
    xyzz_t r;
    r.x = ldg_fe(&p->x);
    r.y = ldg_fe(&p->y);
    r.zz = ldg_fe(&p->zz);
    r.zzz = ldg_fe(&p->zzz);
    return r;
}
__device__ __forceinline__ void st_xyzz(xyzz_t* p, const xyzz_t& v) {
    stg_fe(&p->x, v.x);
    stg_fe(&p->y, v.y);
    stg_fe(&p->zz, v.zz);
    stg_fe(&p->zzz, v.zzz);
}

// ---- registration: T[j][i] = 2^(c*j) * P_i -------------------------------------------
static constexpr int MAX_W = 32;

__global__ void k_precompute(const affine_t* __restrict__ bases, affine_t* __restrict__ table, xyzz_t* tmp,
                             uint32_t n, uint32_t i0, uint32_t cnt, uint32_t c, uint32_t W) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt) return;
    uint32_t i = i0 + t;
    affine_t P = bases[i];
    table[i] = P;
    xyzz_t acc = xyzz_from_affine(P);
    fe_t pref[MAX_W];
    fe_t run = Fq::one();
    bool ident = xyzz_is_identity(acc);
    for (uint32_t j = 1; j < W; j++) {
        for (uint32_t d = 0; d < c; d++) acc = xyzz_double(acc);
        tmp[(size_t)(j - 1) * cnt + t] = acc;
        pref[j] = run;
        if (!ident) run = Fq::mul(run, Fq::mul(acc.zz, acc.zzz));
    }
    fe_t inv = Fq::inv(run);
    for (uint32_t j = W - 1; j >= 1; j--) {
        xyzz_t q = tmp[(size_t)(j - 1) * cnt + t];
        affine_t o;
        if (ident) {
            o.x = Fq::zero();
            o.y = Fq::zero();
        } else {
            fe_t d = Fq::mul(q.zz, q.zzz);
            fe_t di = Fq::mul(inv, pref[j]);  // 1 / (zz * zzz)
            inv = Fq::mul(inv, d);
            o.x = Fq::mul(q.x, Fq::mul(di, q.zzz));
            o.y = Fq::mul(q.y, Fq::mul(di, q.zz));
        }
        table[(size_t)j * n + i] = o;
    }
}

// ---- suffix sums of a base set: S_j = sum_{i >= j} P_i ---------------------------------------
// A vector that is constant over long row runs commits as  sum_i a_i P_i = sum_j (a_j - a_{j-1}) S_j : an MSM whose
// non-zero scalars are only the rows where the vector changes.  The grand-product columns Z of the prover are like
// that (the ratio is 1 on every row that takes part in no copy constraint / lookup, e.g. all rows behind the circuit:
// 43 % of the column at RSA-2048 / k = 17), so they are committed against the S_j of g_lagrange (prover.cu).
static constexpr uint32_t SFX_CHUNK = 32;
__global__ void __launch_bounds__(128) k_sfx_local(const affine_t* __restrict__ in, xyzz_t* __restrict__ chunk_tot, uint32_t n) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t * SFX_CHUNK >= n) return;
    const uint32_t lo = t * SFX_CHUNK, hi = min(lo + SFX_CHUNK, n);
    xyzz_t acc = xyzz_identity();
    for (uint32_t i = lo; i < hi; i++) xyzz_madd(acc, in[i], false);
    chunk_tot[t] = acc;
}
// exclusive suffix scan of the chunk totals; one thread (one-off per SRS, n / 32 full additions)
__global__ void k_sfx_scan(xyzz_t* chunk_tot, uint32_t nchunks) {
    if (blockIdx.x || threadIdx.x) return;
    xyzz_t run = xyzz_identity();
    for (uint32_t t = nchunks; t-- > 0;) {
        xyzz_t v = chunk_tot[t];
        chunk_tot[t] = run;
        xyzz_add(run, v);
    }
}
__global__ void __launch_bounds__(128) k_sfx_write(const affine_t* __restrict__ in, const xyzz_t* __restrict__ carry, xyzz_t* __restrict__ out,
                                                   uint32_t n) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t * SFX_CHUNK >= n) return;
    const uint32_t lo = t * SFX_CHUNK, hi = min(lo + SFX_CHUNK, n);
    xyzz_t acc = carry[t];
    for (uint32_t i = hi; i-- > lo;) {
        xyzz_madd(acc, in[i], false);
        out[i] = acc;
    }
}
__global__ void __launch_bounds__(128) k_normalize(const xyzz_t* __restrict__ in, affine_t* __restrict__ out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = xyzz_to_affine(in[i]);
}

// ---- digits -----------------------------------------------------------------------------
// signed window j of canonical scalar k: digit in (-2^(c-1), 2^(c-1)].  C is a template parameter and the window
// loop is fully unrolled, so every word index and shift below is a compile-time constant (the first version indexed
// the limb array at run time: k_digits was alu-pipe bound at 80 %, profiles/r01_ncu_prover.md).
template <int C, int J>
__device__ __forceinline__ int32_t signed_digit(const fe_t& k, uint32_t& carry) {
    constexpr uint32_t bit = (uint32_t)J * C, w = bit >> 5, s = bit & 31;
    uint32_t raw = 0;
    if (w < 8) {
        raw = k.l[w] >> s;
        if (s + C > 32 && w + 1 < 8) raw |= k.l[w + 1] << (32 - s);
    }
    raw = (raw & ((1u << C) - 1)) + carry;
    if (raw > (1u << (C - 1))) {
        carry = 1;
        return (int32_t)raw - (int32_t)(1u << C);
    }
    carry = 0;
    return (int32_t)raw;
}

// AGG: lanes of a warp that hit the same bucket are combined into one atomic (MATCH.ANY).  Columns of small or repeated
// values need it (thousands of rows share a bucket); for scalars the caller knows to be uniformly random it is pure
// overhead (k_digits spent 3x longer per vector on them than on run-structured columns), so those take one atomic per lane.
template <bool SCATTER, bool AGG>
__device__ __forceinline__ void digit_emit(int32_t d, bool live, uint32_t j, uint32_t i, uint32_t n_table, uint32_t lane, uint32_t* cnt,
                                           uint32_t* ent, uint32_t* key) {
    const bool has = live && d != 0;
    const uint32_t b = has ? (uint32_t)(d < 0 ? -d : d) - 1 : 0xffffffffu;
    if (!AGG) {
        if (has) {
            if (SCATTER) {
                const uint32_t pos = atomicAdd(&cnt[b], 1u);
                ent[pos] = (j * n_table + i) | (d < 0 ? 0x80000000u : 0u);
                key[pos] = b;
            } else {
                atomicAdd(&cnt[b], 1u);
            }
        }
        return;
    }
    // warp-aggregate lanes that hit the same bucket
    const uint32_t act = __ballot_sync(0xffffffffu, has);
    if (has) {
        const uint32_t peers = __match_any_sync(act, b);
        const uint32_t leader = __ffs(peers) - 1;
        const uint32_t rank = __popc(peers & ((1u << lane) - 1));
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&cnt[b], __popc(peers));
        if (SCATTER) {
            base = __shfl_sync(peers, base, leader);
            ent[base + rank] = (j * n_table + i) | (d < 0 ? 0x80000000u : 0u);
            key[base + rank] = b;
        }
    }
}

template <bool SCATTER, bool AGG, int C, int J, int W>
struct DigitLoop {
    static __device__ __forceinline__ void run(const fe_t& k, uint32_t& carry, bool live, uint32_t i, uint32_t n_table, uint32_t lane,
                                               uint32_t* cnt, uint32_t* ent, uint32_t* key) {
        const int32_t d = signed_digit<C, J>(k, carry);
        digit_emit<SCATTER, AGG>(d, live, J, i, n_table, lane, cnt, ent, key);
        DigitLoop<SCATTER, AGG, C, J + 1, W>::run(k, carry, live, i, n_table, lane, cnt, ent, key);
    }
};
template <bool SCATTER, bool AGG, int C, int W>
struct DigitLoop<SCATTER, AGG, C, W, W> {
    static __device__ __forceinline__ void run(const fe_t&, uint32_t&, bool, uint32_t, uint32_t, uint32_t, uint32_t*, uint32_t*, uint32_t*) {}
};

template <bool SCATTER, int C, bool AGG>
__global__ void __launch_bounds__(256)
k_digits(const fe_t* __restrict__ scalars, uint32_t n, uint32_t n_table, uint32_t B,
         uint32_t* counts /*[G][B] (count) or cursor (scatter)*/, uint32_t* entries, uint32_t* keys, size_t ent_stride) {
    constexpr int W = (255 + C - 1) / C;
    uint32_t g = blockIdx.y;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool live = i < n;
    fe_t s = Fr::zero();
    if (live) s = ldg_fe(scalars + (size_t)g * n + i);
    // the sparse columns this kernel serves (first differences of A', S', unused advice rows) are mostly zero: a warp
    // whose 32 scalars are all zero has nothing to convert or emit (the whole warp leaves: the votes below stay complete)
    if (__all_sync(0xffffffffu, Fr::is_zero(s))) return;
    s = Fr::from_mont(s);
    uint32_t* cnt = counts + (size_t)g * B;
    uint32_t* ent = SCATTER ? entries + (size_t)g * ent_stride : nullptr;
    uint32_t* key = SCATTER ? keys + (size_t)g * ent_stride : nullptr;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t carry = 0;
    DigitLoop<SCATTER, AGG, C, 0, W>::run(s, carry, live, i, n_table, lane, cnt, ent, key);
}

template <bool SCATTER>
static void launch_digits(uint32_t c, bool agg, dim3 grid, cudaStream_t st, const fe_t* scalars, uint32_t n, uint32_t n_table, uint32_t B,
                          uint32_t* counts, uint32_t* entries, uint32_t* keys, size_t ent_stride) {
#define B2R_DIG(C, A) k_digits<SCATTER, C, A><<<grid, 256, 0, st>>>(scalars, n, n_table, B, counts, entries, keys, ent_stride)
    switch (c) {
        case 10: if (agg) B2R_DIG(10, true); else B2R_DIG(10, false); break;
        case 13: if (agg) B2R_DIG(13, true); else B2R_DIG(13, false); break;
        default: if (agg) B2R_DIG(16, true); else B2R_DIG(16, false); break;
    }
#undef B2R_DIG
}

// digit loop of k_bin_scatter: window J of the scalar -> packed entry (table index | sign << idx_bits | low LB bucket bits
// << (idx_bits + 1)) and its rank among the CTA's entries of the same high-bits bin
template <int C, int J, int W, int LB>
struct DigitLoopBins {
    static __device__ __forceinline__ void run(const fe_t& k, uint32_t& carry, uint32_t i, uint32_t n_table, uint32_t idx_bits, uint32_t* hist,
                                               uint32_t* packed, uint32_t* where) {
        const int32_t d = signed_digit<C, J>(k, carry);
        if (d != 0) {
            const uint32_t b = (uint32_t)(d < 0 ? -d : d) - 1, hi = b >> LB, lo = b & ((1u << LB) - 1);
            packed[J] = ((uint32_t)J * n_table + i) | ((d < 0 ? 1u : 0u) << idx_bits) | (lo << (idx_bits + 1));
            where[J] = (hi << 16) | atomicAdd(&hist[hi], 1u);
        } else {
            packed[J] = 0;
            where[J] = 0xffffffffu;
        }
        DigitLoopBins<C, J + 1, W, LB>::run(k, carry, i, n_table, idx_bits, hist, packed, where);
    }
};
template <int C, int W, int LB>
struct DigitLoopBins<C, W, W, LB> {
    static __device__ __forceinline__ void run(const fe_t&, uint32_t&, uint32_t, uint32_t, uint32_t, uint32_t*, uint32_t*, uint32_t*) {}
};

// ---- binned counting sort for vectors the caller declares uniform (c = 16) ---------------------------------------
// The per-entry global atomics and scattered 4-byte stores of k_digits are what bound it on dense random scalars
// (56 us per 2^17 vector against 350 us of accumulation).  Uniform digits make every group of 256 buckets receive
// ~16 K entries, so the sort is done in two coalesced steps: k_bin_scatter partitions the entries by the high 7 (or 8) bits
// of the bucket (shared-memory ranks, one global reservation per CTA and bin, runs of ~32 entries written together into
// fixed-capacity bins), k_bin_sort finishes each bin in shared memory by the low 8 bits and writes the final sorted
// list, the keys and the bucket offsets.  A bin that overflows its capacity raises a flag and the group is redone by
// the general kernels.
// HB = high bits: 7 (128 bins of 256 buckets) up to 2^17 scalars, 8 (256 bins of 128 buckets) for 2^18, 9 for 2^19 and
// 10 (1024 bins of 32 buckets) for 2^20 (BASELINE configs[4]), so that a bin holds ~16 K entries every time.
static constexpr uint32_t BIN_CAP = 20480, BIN_SORT_T = 512, BIN_MAX = 256 /* buckets per bin */, BIN_HI_MAX = 1024 /* bins per vector */;

template <int C, int HB>
__global__ void __launch_bounds__(256)
k_bin_scatter(const fe_t* __restrict__ scalars, uint32_t n, uint32_t n_table, uint32_t idx_bits, uint32_t* __restrict__ bin_fill /*[G][HI]*/,
              uint32_t* __restrict__ bins /*[G][HI][CAP]*/, uint32_t* __restrict__ flags) {
    constexpr int W = (255 + C - 1) / C;
    constexpr uint32_t HI = 1u << HB, PER = HI / 32;
    __shared__ uint32_t hist[HI], base[HI], lstart[HI + 1];
    __shared__ uint32_t stage[256 * W];      // the CTA's entries grouped by bin, so that every bin's run leaves as one contiguous store
    __shared__ uint16_t stage_bin[256 * W];  // bin of every staged entry
    const uint32_t g = blockIdx.y, tid = threadIdx.x, i = blockIdx.x * 256 + tid, lane = tid & 31, wid = tid >> 5;
    for (uint32_t h = tid; h < HI; h += 256) hist[h] = 0;
    __syncthreads();
    fe_t s = Fr::zero();
    if (i < n) s = ldg_fe(scalars + (size_t)g * n + i);
    uint32_t packed[W], where[W];   // where = hi << 16 | rank inside the CTA's share of the bin; 0xffffffff = no entry
    // the grand-product differences are zero on most rows: a warp of 32 zero scalars skips the conversion and the digits
    if (__all_sync(0xffffffffu, Fr::is_zero(s))) {
#pragma unroll
        for (int j = 0; j < W; j++) { packed[j] = 0; where[j] = 0xffffffffu; }
    } else {
        s = Fr::from_mont(s);
        uint32_t carry = 0;
        DigitLoopBins<C, 0, W, C - 1 - HB>::run(s, carry, i, n_table, idx_bits, hist, packed, where);
    }
    __syncthreads();
    for (uint32_t h = tid; h < HI; h += 256) base[h] = hist[h] ? atomicAdd(&bin_fill[(size_t)g * HI + h], hist[h]) : 0;
    if (wid == 7) {   // exclusive scan of the local counts by one warp, PER bins per lane
        uint32_t tot = 0;
#pragma unroll
        for (uint32_t q = 0; q < PER; q++) tot += hist[lane * PER + q];
        uint32_t x = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= (uint32_t)d) x += y;
        }
        uint32_t e = x - tot;
#pragma unroll
        for (uint32_t q = 0; q < PER; q++) { lstart[lane * PER + q] = e; e += hist[lane * PER + q]; }
        if (lane == 31) lstart[HI] = x;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < W; j++) {
        if (where[j] == 0xffffffffu) continue;
        const uint32_t h = where[j] >> 16, pos = lstart[h] + (where[j] & 0xffffu);
        stage[pos] = packed[j];
        stage_bin[pos] = (uint16_t)h;
    }
    __syncthreads();
    uint32_t* mine = bins + (size_t)g * HI * BIN_CAP;
    bool over = false;
    const uint32_t total = lstart[HI];
    for (uint32_t p = tid; p < total; p += 256) {   // neighbours in the stage are neighbours in their bin
        const uint32_t h = stage_bin[p], dst = base[h] + (p - lstart[h]);
        if (dst < BIN_CAP) mine[(size_t)h * BIN_CAP + dst] = stage[p];
        else over = true;
    }
    if (over) atomicOr(flags + g, 1u);
}

// exclusive scan of the HI bin sizes of every vector -> bin_base[g][0..HI] (bin_base[g][HI] = number of entries)
__global__ void __launch_bounds__(BIN_HI_MAX) k_bin_prefix(const uint32_t* __restrict__ bin_fill, uint32_t* __restrict__ bin_base, uint32_t* __restrict__ off,
                                                           uint32_t B, uint32_t HI) {
    __shared__ uint32_t sh[BIN_HI_MAX];
    const uint32_t g = blockIdx.x, t = threadIdx.x;
    const uint32_t v = t < HI ? min(bin_fill[(size_t)g * HI + t], BIN_CAP) : 0;
    sh[t] = v;
    __syncthreads();
    for (uint32_t d = 1; d < BIN_HI_MAX; d <<= 1) {
        const uint32_t o = t >= d ? sh[t - d] : 0;
        __syncthreads();
        sh[t] += o;
        __syncthreads();
    }
    if (t < HI) bin_base[(size_t)g * (HI + 1) + t] = sh[t] - v;
    if (t == HI - 1) {
        bin_base[(size_t)g * (HI + 1) + HI] = sh[t];
        off[(size_t)g * (B + 1) + B] = sh[t];
    }
}

__global__ void __launch_bounds__(BIN_SORT_T)
k_bin_sort(const uint32_t* __restrict__ bins, const uint32_t* __restrict__ bin_fill, const uint32_t* __restrict__ bin_base, uint32_t idx_bits,
           uint32_t B, uint32_t HI, uint32_t* __restrict__ off, uint32_t* __restrict__ entries, uint32_t* __restrict__ keys, size_t ent_stride) {
    extern __shared__ uint32_t sh_ent[];   // BIN_CAP packed entries
    __shared__ uint32_t cnt[BIN_MAX], start[BIN_MAX], cur[BIN_MAX];
    const uint32_t h = blockIdx.x, g = blockIdx.y, tid = threadIdx.x, LO = B / HI;
    const uint32_t m = min(bin_fill[(size_t)g * HI + h], BIN_CAP);
    const uint32_t gbase = bin_base[(size_t)g * (HI + 1) + h];
    const uint32_t* src = bins + ((size_t)g * HI + h) * BIN_CAP;
    if (tid < BIN_MAX) cnt[tid] = 0;
    __syncthreads();
    for (uint32_t k = tid; k < m; k += BIN_SORT_T) atomicAdd(&cnt[src[k] >> (idx_bits + 1)], 1u);
    __syncthreads();
    if (tid < BIN_MAX) start[tid] = cnt[tid];
    __syncthreads();
    for (uint32_t d = 1; d < BIN_MAX; d <<= 1) {   // inclusive scan of the counts
        uint32_t o = 0;
        if (tid < BIN_MAX && tid >= d) o = start[tid - d];
        __syncthreads();
        if (tid < BIN_MAX) start[tid] += o;
        __syncthreads();
    }
    if (tid < BIN_MAX) {
        const uint32_t excl = start[tid] - cnt[tid];
        cur[tid] = excl;
        if (tid < LO) off[(size_t)g * (B + 1) + h * LO + tid] = gbase + excl;
    }
    __syncthreads();
    // second read of the bin (64-80 KB, L2-resident) places every entry at its sorted position in shared memory; the
    // final list and keys then leave as fully coalesced stores
    for (uint32_t k = tid; k < m; k += BIN_SORT_T) {
        const uint32_t p = src[k];
        sh_ent[atomicAdd(&cur[p >> (idx_bits + 1)], 1u)] = p;
    }
    __syncthreads();
    uint32_t* ent = entries + (size_t)g * ent_stride + gbase;
    uint32_t* key = keys + (size_t)g * ent_stride + gbase;
    const uint32_t idx_mask = (1u << idx_bits) - 1;
    for (uint32_t k = tid; k < m; k += BIN_SORT_T) {
        const uint32_t p = sh_ent[k];
        ent[k] = (p & idx_mask) | (((p >> idx_bits) & 1u) << 31);
        key[k] = h * LO + (p >> (idx_bits + 1));
    }
}

// ---- single-CTA exclusive scan per vector: offsets[g][0..B], cursor[g][b] = offsets[g][b]
__global__ void __launch_bounds__(1024) k_scan(const uint32_t* counts, uint32_t* offsets, uint32_t* cursor, uint32_t B) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    uint32_t g = blockIdx.x;
    const uint32_t* cnt = counts + (size_t)g * B;
    uint32_t* off = offsets + (size_t)g * (B + 1);
    uint32_t* cur = cursor + (size_t)g * B;
    uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < B; base += 1024) {
        uint32_t idx = base + tid;
        uint32_t v = idx < B ? cnt[idx] : 0;
        uint32_t x = v;
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= (uint32_t)d) x += y;
        }
        if (lane == 31) warp_tot[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = warp_tot[lane], wx = w;
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, wx, d);
                if (lane >= (uint32_t)d) wx += y;
            }
            warp_tot[lane] = wx - w;  // exclusive
        }
        __syncthreads();
        uint32_t excl = carry_s + warp_tot[wid] + x - v;
        if (idx < B) {
            off[idx] = excl;
            cur[idx] = excl;
        }
        __syncthreads();
        if (tid == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (tid == 0) off[B] = carry_s;
}

// ---- load-balanced segmented bucket sums -------------------------------------------------
__device__ __forceinline__ void emit_run(uint32_t b, const xyzz_t& acc, bool begins, bool ends, bool first_in_chunk,
                                         xyzz_t* buckets, uint32_t* slot_keys, xyzz_t* slot_pts, size_t slot0) {
    if (begins && ends) {
        st_xyzz(buckets + b, acc);
    } else {
        size_t s = slot0 + (first_in_chunk ? 0 : 1);
        slot_keys[s] = b | (begins ? SLOT_BEGINS : 0u) | (ends ? SLOT_ENDS : 0u);
        st_xyzz(slot_pts + s, acc);
    }
}

// Lock-step formulation: every thread owns exactly L consecutive entries of the bucket-sorted list
// and walks them with ONE flat loop, so all lanes of a warp execute the same mixed add in the same
// iteration whatever the bucket boundaries are (the nested run loops of the first version left 9.8
// of 32 lanes active, profiles/r01_ncu_full_baseline.md).  A bucket change costs a predicated flush.
template <int L, int MINB, bool PF>
__global__ void __launch_bounds__(128, MINB)
k_accum_entries(const affine_t* __restrict__ table, const uint32_t* __restrict__ entries, const uint32_t* __restrict__ keys,
                size_t ent_stride, const uint32_t* __restrict__ offsets, uint32_t B, uint32_t nchunks, xyzz_t* buckets,
                uint32_t* slot_keys, xyzz_t* slot_pts, size_t slot_stride) {
    uint32_t g = blockIdx.y;
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nchunks) return;
    const uint32_t* off = offsets + (size_t)g * (B + 1);
    const uint32_t* ent = entries + (size_t)g * ent_stride;
    const uint32_t* key = keys + (size_t)g * ent_stride;
    xyzz_t* bk = buckets + (size_t)g * B;
    uint32_t* sk = slot_keys + (size_t)g * slot_stride;
    xyzz_t* sp = slot_pts + (size_t)g * slot_stride;
    size_t slot0 = 2 * (size_t)t;
    sk[slot0] = SLOT_INVALID;
    sk[slot0 + 1] = SLOT_INVALID;
    const uint32_t E = off[B];
    const uint32_t start = t * L;
    if (start >= E) return;
    const uint32_t end = min(start + (uint32_t)L, E);
    uint32_t cur = key[start];
    bool begins = (off[cur] == start), first = true;
    uint32_t e = ent[start];
    xyzz_t acc = xyzz_from_affine_signed(ldg_affine(table + (e & 0x7fffffffu)), (e >> 31) != 0);
    uint32_t e_next = 0, k_next = cur;
    affine_t p_next;
    p_next.x = Fq::zero();
    p_next.y = Fq::zero();
    if (start + 1 < end) {
        e_next = ent[start + 1];
        k_next = key[start + 1];
        if (PF) p_next = ldg_affine(table + (e_next & 0x7fffffffu));
    }
#pragma unroll 1
    for (uint32_t pos = start + 1; pos < end; pos++) {
        const uint32_t ec = e_next, kc = k_next;
        // PF: the table point of the next entry is fetched one iteration ahead (16 more live registers);
        // otherwise only its index is, and the 64 B gather is issued at the top of its own iteration
        const affine_t p = PF ? p_next : ldg_affine(table + (ec & 0x7fffffffu));
        if (pos + 1 < end) {
            e_next = ent[pos + 1];
            k_next = key[pos + 1];
            if (PF) p_next = ldg_affine(table + (e_next & 0x7fffffffu));
        }
        if (kc != cur) {
            emit_run(cur, acc, begins, true, first, bk, sk, sp, slot0);
            first = false;
            begins = true;
            cur = kc;
            acc = xyzz_from_affine_signed(p, (ec >> 31) != 0);
        } else {
            xyzz_madd_ls(acc, p, (ec >> 31) != 0);
        }
    }
    emit_run(cur, acc, begins, end == off[cur + 1], first, bk, sk, sp, slot0);
}

// ---- the same walk with the table points staged through shared memory by the TMA engine ---------------------------
// The 64-byte table point of entry pos + 1 is fetched while entry pos is being added: every lane issues one
// cp.async.bulk (global -> its own shared-memory slot) that completes on its warp's mbarrier for that ring stage, and
// reads the point back with four conflict-free LDS.128 right where the addition needs it.  Unlike the register prefetch
// (PF above: 16 more live registers, spills at 127) the point in flight costs no registers, so the DRAM / L2 latency of
// the gather (the table is 128 MiB, 37 % L2 hits) is hidden behind a full mixed addition instead of being exposed at
// the top of every iteration.  Ring: RING_D stages x 128 lanes x 80 bytes (64-byte point + 16 bytes of padding: eight
// consecutive lanes then start in eight different 16-byte bank groups).
static constexpr uint32_t RING_D = 2, RING_SLOT = 80;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ affine_t lds_affine(uint32_t addr) {
    affine_t r;
    uint32_t* w = r.x.l;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(addr));
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "r"(addr + 16));
    w = r.y.l;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(addr + 32));
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "r"(addr + 48));
    return r;
}

__device__ __forceinline__ void ldgsts64(uint32_t dst, const void* src) {
    const char* s = (const char*)src;
#pragma unroll
    for (int j = 0; j < 4; j++) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * j), "l"(s + 16 * j) : "memory");
}
template <int L, int MINB, bool BULK>
__global__ void __launch_bounds__(128, MINB)
k_accum_entries_ring(const affine_t* __restrict__ table, const uint32_t* __restrict__ entries, const uint32_t* __restrict__ keys,
                     size_t ent_stride, const uint32_t* __restrict__ offsets, uint32_t B, uint32_t nchunks, xyzz_t* buckets,
                     uint32_t* slot_keys, xyzz_t* slot_pts, size_t slot_stride) {
    __shared__ __align__(16) uint8_t ring[RING_D * 128 * RING_SLOT];
    __shared__ __align__(8) uint64_t bars[RING_D * 4];
    const uint32_t g = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t t = blockIdx.x * blockDim.x + tid;
    if (BULK) {
        if (lane == 0) {
            for (uint32_t s = 0; s < RING_D; s++) mbar_init(smem_addr(&bars[s * 4 + wid]), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    const uint32_t* off = offsets + (size_t)g * (B + 1);
    const uint32_t* ent = entries + (size_t)g * ent_stride;
    const uint32_t* key = keys + (size_t)g * ent_stride;
    xyzz_t* bk = buckets + (size_t)g * B;
    uint32_t* sk = slot_keys + (size_t)g * slot_stride;
    xyzz_t* sp = slot_pts + (size_t)g * slot_stride;
    const size_t slot0 = 2 * (size_t)t;
    if (t < nchunks) {
        sk[slot0] = SLOT_INVALID;
        sk[slot0 + 1] = SLOT_INVALID;
    }
    const uint32_t E = off[B];
    const uint32_t start = t * L;
    // no early return: the warp-wide votes below need every lane; a lane without entries has end == start
    const bool live = t < nchunks && start < E;
    const uint32_t end = live ? min(start + (uint32_t)L, E) : start;
    uint32_t ring_base[RING_D], bar_addr[RING_D];
#pragma unroll
    for (uint32_t s = 0; s < RING_D; s++) {
        ring_base[s] = smem_addr(ring) + (s * 128 + tid) * RING_SLOT;
        bar_addr[s] = smem_addr(&bars[s * 4 + wid]);
    }
    uint32_t cur = 0, e_next = 0, k_next = 0;
    bool begins = false, first = true;
    xyzz_t acc = xyzz_identity();
    if (live) {
        cur = key[start];
        begins = (off[cur] == start);
        const uint32_t e = ent[start];
        acc = xyzz_from_affine_signed(ldg_affine(table + (e & 0x7fffffffu)), (e >> 31) != 0);
        k_next = cur;
    }
    // stage for i = 1 (entry start + 1)
    {
        const bool v1 = start + 1 < end;
        if (v1) {
            e_next = ent[start + 1];
            k_next = key[start + 1];
        }
        if (BULK) {
            const uint32_t bal = __ballot_sync(0xffffffffu, v1);
            if (lane == 0 && bal) mbar_expect_tx(bar_addr[1 % RING_D], 64u * __popc(bal));
            __syncwarp();
            if (v1) bulk_g2s(ring_base[1 % RING_D], table + (e_next & 0x7fffffffu), 64u, bar_addr[1 % RING_D]);
        } else {
            if (v1) ldgsts64(ring_base[1 % RING_D], table + (e_next & 0x7fffffffu));
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    }
#pragma unroll 1
    for (uint32_t i = 1; i < (uint32_t)L; i++) {
        const uint32_t pos = start + i;
        const bool v = pos < end, vn = pos + 1 < end;
        const uint32_t bal = __ballot_sync(0xffffffffu, v);
        if (!bal) break;
        const uint32_t ec = e_next, kc = k_next;
        // issue the gather of entry pos + 1 into the other stage (its previous content, entry pos - 1, was consumed in
        // the previous iteration), then wait for this iteration's stage
        if (vn) {
            e_next = ent[pos + 1];
            k_next = key[pos + 1];
        }
        const uint32_t sn = (i + 1) % RING_D, sc = i % RING_D;
        if (BULK) {
            const uint32_t baln = __ballot_sync(0xffffffffu, vn);
            if (lane == 0 && baln) mbar_expect_tx(bar_addr[sn], 64u * __popc(baln));
            __syncwarp();
            if (vn) bulk_g2s(ring_base[sn], table + (e_next & 0x7fffffffu), 64u, bar_addr[sn]);
            mbar_wait(bar_addr[sc], ((i - (sc ? sc : RING_D)) / RING_D) & 1u);   // parity of this stage's use count
        } else {
            if (vn) ldgsts64(ring_base[sn], table + (e_next & 0x7fffffffu));
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 1;" ::: "memory");   // everything but the group just committed has landed
        }
        if (v) {
            const affine_t p = lds_affine(ring_base[sc]);
            if (kc != cur) {
                emit_run(cur, acc, begins, true, first, bk, sk, sp, slot0);
                first = false;
                begins = true;
                cur = kc;
                acc = xyzz_from_affine_signed(p, (ec >> 31) != 0);
            } else {
                xyzz_madd_ls(acc, p, (ec >> 31) != 0);
            }
        }
    }
    if (live) emit_run(cur, acc, begins, end == off[cur + 1], first, bk, sk, sp, slot0);
}

// upper levels: the same walk over a slot list (key + XYZZ partial sum per slot)
template <int L>
__global__ void __launch_bounds__(128)
k_accum_slots(const uint32_t* __restrict__ in_keys, const xyzz_t* __restrict__ in_pts, size_t in_stride, uint32_t M,
              uint32_t nchunks, uint32_t B, xyzz_t* buckets, uint32_t* slot_keys, xyzz_t* slot_pts, size_t slot_stride) {
    uint32_t g = blockIdx.y;
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nchunks) return;
    const uint32_t* ik = in_keys + (size_t)g * in_stride;
    const xyzz_t* ip = in_pts + (size_t)g * in_stride;
    xyzz_t* bk = buckets + (size_t)g * B;
    uint32_t* sk = slot_keys + (size_t)g * slot_stride;
    xyzz_t* sp = slot_pts + (size_t)g * slot_stride;
    size_t slot0 = 2 * (size_t)t;
    sk[slot0] = SLOT_INVALID;
    sk[slot0 + 1] = SLOT_INVALID;
    const uint32_t start = t * L, end = min(start + (uint32_t)L, M);
    bool have = false, first = true, begins = false, ends = false;
    uint32_t cb = 0;
    xyzz_t acc = xyzz_identity();
#pragma unroll 1
    for (uint32_t s = start; s < end; s++) {
        const uint32_t key = ik[s];
        if (key == SLOT_INVALID) continue;
        const uint32_t b = key & SLOT_KEY;
        const xyzz_t q = ld_xyzz(ip + s);
        if (have && b == cb) {
            xyzz_add_ls(acc, q);
            ends = (key & SLOT_ENDS) != 0;
        } else {
            if (have) {
                emit_run(cb, acc, begins, ends, first, bk, sk, sp, slot0);
                first = false;
            }
            have = true;
            cb = b;
            acc = q;
            begins = (key & SLOT_BEGINS) != 0;
            ends = (key & SLOT_ENDS) != 0;
        }
    }
    if (have) emit_run(cb, acc, begins, ends, first, bk, sk, sp, slot0);
}

// One thread per bucket gathers the chunk-boundary partial sums k_accum_entries left in the level-A slot list, in one
// launch instead of the log_16 levels of k_accum_slots (seven launches of latency-bound 16-addition chains for a
// standalone 2^20 MSM: 0.39 ms).  Bucket b covers entries [off[b], off[b+1]) = chunks t0 .. t1 of L entries; it was
// stored directly if it lies inside one chunk, otherwise chunk t holds its partial sum in slot 2t (the bucket is the
// first run of the chunk) or 2t + 1 (only possible for t0, when the bucket starts inside the chunk).  Used behind the
// binned sort, whose bins bound a bucket to a few hundred chunks; arbitrary distributions keep the level scheme.
template <int L>
__global__ void __launch_bounds__(128)
k_bucket_fixup(const uint32_t* __restrict__ offsets, uint32_t B, const xyzz_t* __restrict__ slot_pts, size_t slot_stride, xyzz_t* __restrict__ buckets) {
    const uint32_t g = blockIdx.y, b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t* off = offsets + (size_t)g * (B + 1);
    const uint32_t lo = off[b], hi = off[b + 1];
    if (hi <= lo) return;
    const uint32_t t0 = lo / L, t1 = (hi - 1) / L;
    if (t0 == t1) return;   // complete inside one chunk: stored by k_accum_entries
    const xyzz_t* sp = slot_pts + (size_t)g * slot_stride;
    xyzz_t acc = ld_xyzz(sp + 2 * (size_t)t0 + (lo == t0 * L ? 0 : 1));
#pragma unroll 1
    for (uint32_t t = t0 + 1; t <= t1; t++) {
        const xyzz_t q = ld_xyzz(sp + 2 * (size_t)t);
        xyzz_add(acc, q);
    }
    st_xyzz(buckets + (size_t)g * B + b, acc);
}

// ---- sum_b (b+1) * B_b -------------------------------------------------------------------
// Three levels of running sums, every thread busy in the two large ones (the first version combined 128 threads
// per CTA with shared-memory scans and a one-thread double-and-add tail: 1.9x the additions, most of them in
// nearly empty warps - profiles/r01_ncu_full_baseline.md).
//   level 1  thread = 32 consecutive buckets:  S1_s = sum B,  A1_s = sum (d + 1) B_{32 s + d}        (2 adds / bucket)
//   level 2  thread = 8 consecutive segments:  S2_u = sum S1, A2_u = sum d S1_{8 u + d}, P1_u = sum A1  (3 adds / segment)
//   level 3  CTA per vector over the nu = B / 256 level-2 outputs: W = sum_u u S2_u (suffix scan + tree), sums of A2, P1
//   result = P1 + 32 (A2 + 8 W), one inversion to affine.
static constexpr uint32_t BR_PER1 = 32, BR_PER2 = 8, BR_T3 = 128;
#ifndef B2R_BR1_MINB
#define B2R_BR1_MINB 3
#endif

__global__ void __launch_bounds__(128, B2R_BR1_MINB)
k_br_level1(const xyzz_t* __restrict__ buckets, uint32_t B, uint32_t nseg, xyzz_t* __restrict__ S1, xyzz_t* __restrict__ A1) {
    const uint32_t g = blockIdx.y, s = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = s < nseg;  // no early return: the warp votes below need every lane
    const xyzz_t* bk = buckets + (size_t)g * B + (size_t)(live ? s : 0) * BR_PER1;
    xyzz_t run = xyzz_identity(), acc = xyzz_identity();
#pragma unroll 1
    for (uint32_t d = BR_PER1; d-- > 0;) {
        xyzz_t q = xyzz_identity();
        if (live) q = ld_xyzz(bk + d);
        // sparse bucket sets (small scalars): skip the adds a whole warp does not need.  (Issuing acc += run and
        // run += q as two independent additions per step was measured slower: 61 vs 47 ms per 64-proof step.)
        if (!__all_sync(0xffffffffu, xyzz_is_identity(q))) xyzz_add_ls(run, q);
        if (!__all_sync(0xffffffffu, xyzz_is_identity(run))) xyzz_add_ls(acc, run);
    }
    if (live) {
        st_xyzz(S1 + (size_t)g * nseg + s, run);
        st_xyzz(A1 + (size_t)g * nseg + s, acc);
    }
}

// Level 1 for a single vector (or a handful): the per-thread running sums above are a chain of 64 dependent full
// additions with only nseg = B / 32 threads per vector (8 CTAs at c = 16: 0.48 ms of latency for a standalone 2^20 MSM,
// profiles/r01_ncu_config5.md).  Here a WARP owns the 32 buckets of a segment: inclusive suffix scan over the lanes
// (5 shuffle steps) gives T_l = sum_{j >= l} B_j, so S1 = T_0 and A1 = sum_l T_l (bucket d counted d + 1 times), a
// second 5-step tree sum.  10 dependent additions instead of 64, 5x the total work - used when the launch would
// otherwise leave most of the GPU idle.
__device__ __forceinline__ xyzz_t shfl_down_xyzz(const xyzz_t& v, uint32_t d) {
    xyzz_t r;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r.x.l[i] = __shfl_down_sync(0xffffffffu, v.x.l[i], d);
        r.y.l[i] = __shfl_down_sync(0xffffffffu, v.y.l[i], d);
        r.zz.l[i] = __shfl_down_sync(0xffffffffu, v.zz.l[i], d);
        r.zzz.l[i] = __shfl_down_sync(0xffffffffu, v.zzz.l[i], d);
    }
    return r;
}
__global__ void __launch_bounds__(128)
k_br_level1_warp(const xyzz_t* __restrict__ buckets, uint32_t B, uint32_t nseg, uint32_t total_seg, xyzz_t* __restrict__ S1, xyzz_t* __restrict__ A1) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= total_seg) return;   // whole warps leave together
    const uint32_t g = w / nseg, sgm = w % nseg;
    xyzz_t run = ld_xyzz(buckets + (size_t)g * B + (size_t)sgm * BR_PER1 + lane);
#pragma unroll 1
    for (uint32_t d = 1; d < 32; d <<= 1) {
        xyzz_t o = shfl_down_xyzz(run, d);
        if (lane + d >= 32) o = xyzz_identity();
        xyzz_add_ls(run, o);
    }
    xyzz_t acc = run;   // T_lane
#pragma unroll 1
    for (uint32_t d = 16; d > 0; d >>= 1) {
        xyzz_t o = shfl_down_xyzz(acc, d);
        if (lane >= d) o = xyzz_identity();   // only lanes < d keep a meaningful partial sum
        xyzz_add_ls(acc, o);
    }
    if (lane == 0) {
        st_xyzz(S1 + (size_t)g * nseg + sgm, run);
        st_xyzz(A1 + (size_t)g * nseg + sgm, acc);
    }
}

__global__ void __launch_bounds__(128)
k_br_level2(const xyzz_t* __restrict__ S1, const xyzz_t* __restrict__ A1, uint32_t nseg, uint32_t nu, uint32_t total, xyzz_t* __restrict__ L2) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < total;  // total = G * nu
    const uint32_t g = live ? t / nu : 0, u = live ? t % nu : 0;
    const xyzz_t* s1 = S1 + (size_t)g * nseg + (size_t)u * BR_PER2;
    const xyzz_t* a1 = A1 + (size_t)g * nseg + (size_t)u * BR_PER2;
    xyzz_t run = xyzz_identity(), acc = xyzz_identity(), plain = xyzz_identity();
#pragma unroll 1
    for (uint32_t d = BR_PER2; d-- > 0;) {
        xyzz_t q = xyzz_identity(), a = xyzz_identity();
        if (live) {
            q = ld_xyzz(s1 + d);
            a = ld_xyzz(a1 + d);
        }
        // weight d (zero based): the running sum is added before this element joins it
        if (!__all_sync(0xffffffffu, xyzz_is_identity(run))) xyzz_add_ls(acc, run);
        if (!__all_sync(0xffffffffu, xyzz_is_identity(q))) xyzz_add_ls(run, q);
        if (!__all_sync(0xffffffffu, xyzz_is_identity(a))) xyzz_add_ls(plain, a);
    }
    if (live) {
        xyzz_t* o = L2 + (size_t)t * 3;
        st_xyzz(o, run);
        st_xyzz(o + 1, acc);
        st_xyzz(o + 2, plain);
    }
}

__device__ __forceinline__ xyzz_t block_tree_sum(xyzz_t v, xyzz_t* sm, uint32_t tid) {
    __syncthreads();
    sm[tid] = v;
    __syncthreads();
    for (uint32_t s = BR_T3 / 2; s > 0; s >>= 1) {
        if (tid < s) {
            xyzz_t a = sm[tid], b = sm[tid + s];
            xyzz_add_ls(a, b);
            sm[tid] = a;
        }
        __syncthreads();
    }
    return sm[0];
}

__global__ void __launch_bounds__(BR_T3)
k_br_level3(const xyzz_t* __restrict__ L2, uint32_t nu, affine_t* __restrict__ out) {
    __shared__ uint4 smem_raw[BR_T3 * sizeof(xyzz_t) / sizeof(uint4)];
    xyzz_t* sm = reinterpret_cast<xyzz_t*>(smem_raw);
    const uint32_t g = blockIdx.x, tid = threadIdx.x;
    xyzz_t s2 = xyzz_identity(), a2 = xyzz_identity(), p1 = xyzz_identity();
    if (tid < nu) {
        const xyzz_t* in = L2 + ((size_t)g * nu + tid) * 3;
        s2 = ld_xyzz(in);
        a2 = ld_xyzz(in + 1);
        p1 = ld_xyzz(in + 2);
    }
    // inclusive suffix scan of S2 (Hillis-Steele): T_u = sum_{v >= u} S2_v ;  sum_u u S2_u = sum_{u >= 1} T_u
    xyzz_t run = s2;
    sm[tid] = run;
    __syncthreads();
#pragma unroll 1
    for (uint32_t d = 1; d < BR_T3; d <<= 1) {
        xyzz_t o = xyzz_identity();
        if (tid + d < BR_T3) o = sm[tid + d];
        __syncthreads();
        xyzz_add_ls(run, o);
        sm[tid] = run;
        __syncthreads();
    }
    if (tid == 0) run = xyzz_identity();
    const xyzz_t w = block_tree_sum(run, sm, tid);
    const xyzz_t sa = block_tree_sum(a2, sm, tid);
    const xyzz_t sp = block_tree_sum(p1, sm, tid);
    if (tid == 0) {
        xyzz_t r = w;
        for (int i = 0; i < 3; i++) r = xyzz_double(r);  // BR_PER2 = 8
        xyzz_add(r, sa);
        for (int i = 0; i < 5; i++) r = xyzz_double(r);  // BR_PER1 = 32
        xyzz_add(r, sp);
        out[g] = xyzz_to_affine(r);
    }
}

// Levels 2 and 3 for a single vector at c = 16 (nseg = 1024): a warp per 32 segments, then one CTA of three warps.
//   level 2  S2_u = sum_l S1,  A2_u = sum_l l * S1_{32u+l} = sum_{l >= 1} T_l,  P1_u = sum_l A1      (u < 32)
//   level 3  W = sum_u u * S2_u, SA = sum A2, SP = sum P1 by warps 0, 1, 2 side by side; result = SP + 32 (SA + 32 W)
__device__ __forceinline__ xyzz_t warp_tree_sum(xyzz_t v, uint32_t lane) {
#pragma unroll 1
    for (uint32_t d = 16; d > 0; d >>= 1) {
        xyzz_t o = shfl_down_xyzz(v, d);
        if (lane >= d) o = xyzz_identity();
        xyzz_add_ls(v, o);
    }
    return v;   // lane 0
}
// -> lane 0: sum_{l >= 1} T_l = sum_l l * v_l ; total = sum_l v_l
__device__ __forceinline__ xyzz_t warp_weighted_sum(xyzz_t v, uint32_t lane, xyzz_t& total) {
#pragma unroll 1
    for (uint32_t d = 1; d < 32; d <<= 1) {
        xyzz_t o = shfl_down_xyzz(v, d);
        if (lane + d >= 32) o = xyzz_identity();
        xyzz_add_ls(v, o);
    }
    total = v;   // lane 0: T_0
    if (lane == 0) v = xyzz_identity();
    return warp_tree_sum(v, lane);
}
// two warps per output u, side by side: the weighted sum of S1 (10 dependent additions) and the plain sum of A1 (5)
__global__ void __launch_bounds__(128)
k_br_level2_warp(const xyzz_t* __restrict__ S1, const xyzz_t* __restrict__ A1, uint32_t total_u, xyzz_t* __restrict__ L2) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t u = w >> 1;
    if (u >= total_u) return;
    xyzz_t* o = L2 + (size_t)u * 3;
    if (w & 1u) {
        const xyzz_t p1 = warp_tree_sum(ld_xyzz(A1 + (size_t)u * 32 + lane), lane);
        if (lane == 0) st_xyzz(o + 2, p1);
    } else {
        xyzz_t tot;
        const xyzz_t a2 = warp_weighted_sum(ld_xyzz(S1 + (size_t)u * 32 + lane), lane, tot);
        if (lane == 0) {
            st_xyzz(o, tot);
            st_xyzz(o + 1, a2);
        }
    }
}
__global__ void __launch_bounds__(96)
k_br_level3_warp(const xyzz_t* __restrict__ L2 /* [G][32][3] */, affine_t* __restrict__ out) {
    __shared__ uint4 smem_raw[3 * sizeof(xyzz_t) / sizeof(uint4)];
    xyzz_t* sm = reinterpret_cast<xyzz_t*>(smem_raw);
    const uint32_t g = blockIdx.x, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const xyzz_t v = ld_xyzz(L2 + ((size_t)g * 32 + lane) * 3 + wid);
    xyzz_t r, tot;
    if (wid == 0) r = warp_weighted_sum(v, lane, tot);   // W
    else r = warp_tree_sum(v, lane);                     // SA, SP
    if (lane == 0) sm[wid] = r;
    __syncthreads();
    if (threadIdx.x == 0) {
        xyzz_t acc = sm[0];
        for (int i = 0; i < 5; i++) acc = xyzz_double(acc);
        xyzz_add(acc, sm[1]);
        for (int i = 0; i < 5; i++) acc = xyzz_double(acc);
        xyzz_add(acc, sm[2]);
        out[g] = xyzz_to_affine(acc);
    }
}

static uint32_t pick_window(size_t n) {
    if (n <= ((size_t)1 << 10)) return 10;
    if (n <= ((size_t)1 << 14)) return 13;
    if (n <= ((size_t)1 << 18)) return 16;
    return 16;
}

static int32_t msm_group(b2r_ctx* ctx, const b2r_bases* bs, const fe_t* scalars_dev, size_t G, size_t n, affine_t* out_dev, bool uniform) {
    const uint32_t c = bs->c, W = bs->W, B = 1u << (c - 1);
    constexpr uint32_t L2 = 16;
    // entries per accumulation thread.  A batch fills the GPU many times over whatever the chunk length; a single vector
    // is a handful of waves of 4 CTAs x 148 SMs, and a last wave that is 46 % full (2^20 scalars at 64 entries: 3.46
    // waves) costs 13 % of the kernel - pick the chunk length whose CTA count comes closest below a whole number of waves
    uint32_t L1 = MSM_L1;
    const char* variant_env = getenv("B2R_MSM_VARIANT");  // tuning hook: 1 = register prefetch, 3 / 4 = shared-memory ring (fixed chunk length)
    const int variant = variant_env ? atoi(variant_env) : 0;
    if (G <= 4 && variant < 3) {
        const double per_wave = 4.0 * ctx->sm_count;
        double best_eff = 0;
        for (uint32_t cand : {56u, 64u, 74u}) {
            const double ctas = (double)G * (double)(((n * W + cand - 1) / cand + 127) / 128);
            const double waves = ctas / per_wave, eff = waves / (double)(uint64_t)(waves + 0.999999);
            if (eff > best_eff + 1e-9) { best_eff = eff; L1 = cand; }
        }
    }
    const size_t ent_cap = (size_t)n * W;
    const uint32_t nch1 = (uint32_t)((ent_cap + L1 - 1) / L1);
    const size_t slotsA = 2 * (size_t)nch1;
    const uint32_t nch2 = (uint32_t)((slotsA + L2 - 1) / L2);
    const size_t slotsB = 2 * (size_t)nch2;
    if (B % (BR_PER1 * BR_PER2) || B / (BR_PER1 * BR_PER2) > BR_T3) return fail(ctx, B2R_ERR_INVALID, "msm: bucket count does not fit the reduce tiles");
    const uint32_t nseg = B / BR_PER1, nu = nseg / BR_PER2;

    // carve the work arena
    size_t o = 0;
    auto carve = [&](size_t bytes) { size_t r = o; o += (bytes + 255) & ~(size_t)255; return r; };
    size_t o_cnt = carve(G * B * 4), o_off = carve(G * (B + 1) * 4), o_cur = carve(G * B * 4);
    size_t o_ent = carve(G * ent_cap * 4), o_key = carve(G * ent_cap * 4);
    size_t o_bk = carve(G * B * sizeof(xyzz_t));
    size_t o_ka = carve(G * slotsA * 4), o_pa = carve(G * slotsA * sizeof(xyzz_t));
    size_t o_kb = carve(G * slotsB * 4), o_pb = carve(G * slotsB * sizeof(xyzz_t));
    size_t o_s1 = carve(G * nseg * sizeof(xyzz_t)), o_a1 = carve(G * nseg * sizeof(xyzz_t)), o_l2 = carve(G * nu * 3 * sizeof(xyzz_t));
    char* base = nullptr;
    B2R_TRY(scratch_get(ctx, SC_MSM_A, o, (void**)&base));
    uint32_t* cnt = (uint32_t*)(base + o_cnt);
    uint32_t* off = (uint32_t*)(base + o_off);
    uint32_t* cur = (uint32_t*)(base + o_cur);
    uint32_t* ent = (uint32_t*)(base + o_ent);
    uint32_t* key = (uint32_t*)(base + o_key);
    xyzz_t* bk = (xyzz_t*)(base + o_bk);
    uint32_t* ka = (uint32_t*)(base + o_ka);
    xyzz_t* pa = (xyzz_t*)(base + o_pa);
    uint32_t* kb = (uint32_t*)(base + o_kb);
    xyzz_t* pb = (xyzz_t*)(base + o_pb);
    xyzz_t* s1 = (xyzz_t*)(base + o_s1);
    xyzz_t* a1 = (xyzz_t*)(base + o_a1);
    xyzz_t* l2 = (xyzz_t*)(base + o_l2);
    cudaStream_t st = ctx->stream;

    B2R_CUDA(ctx, cudaMemsetAsync(bk, 0, G * B * sizeof(xyzz_t), st));
    dim3 gd((unsigned)((n + 255) / 256), (unsigned)G);
    auto classic_sort = [&]() -> int32_t {
        B2R_CUDA(ctx, cudaMemsetAsync(cnt, 0, G * B * 4, st));
        { KTimer kt(ctx, "msm_count", (double)G * n);
        launch_digits<false>(c, !uniform, gd, st, scalars_dev, (uint32_t)n, (uint32_t)bs->n, B, cnt, nullptr, nullptr, 0); }
        B2R_LAUNCH_CHECK(ctx);
        { KTimer kt(ctx, "msm_scan");
        k_scan<<<(unsigned)G, 1024, 0, st>>>(cnt, off, cur, B); }
        B2R_LAUNCH_CHECK(ctx);
        { KTimer kt(ctx, "msm_scatter", (double)G * n);
        launch_digits<true>(c, !uniform, gd, st, scalars_dev, (uint32_t)n, (uint32_t)bs->n, B, cur, ent, key, ent_cap); }
        B2R_LAUNCH_CHECK(ctx);
        return 0;
    };
    bool sorted = false;
    uint32_t idx_bits = 0;
    while (((size_t)1 << idx_bits) < (size_t)bs->n * W) idx_bits++;
    const char* bin_env = getenv("B2R_MSM_BINSORT");   // "0" disables, "1" forces (tests); read per call so a test can toggle it
    const bool want_bins = bin_env ? bin_env[0] == '1' : uniform;
    // bins of ~n * W / HI entries (capacity 20480 per bin, 25 % slack): 128 bins up to 2^17 scalars ... 1024 bins at 2^20
    uint32_t HB = 7;
    while (HB < 10 && (size_t)n * W * 9 / 8 > ((size_t)BIN_CAP << HB)) HB++;
    const uint32_t HI = 1u << HB;
    if (want_bins && c == 16 && idx_bits + 1 + (c - 1 - HB) <= 32 && (size_t)n * W * 9 / 8 <= (size_t)HI * BIN_CAP) {
        uint32_t* bins = nullptr;
        const size_t fill_bytes = (G * HI * 4 + 255) & ~(size_t)255, base_bytes = (G * (HI + 1) * 4 + 255) & ~(size_t)255;
        B2R_TRY(scratch_get(ctx, SC_MSM_B, fill_bytes + base_bytes + 256 + G * (size_t)HI * BIN_CAP * 4 + G * 4, (void**)&bins));
        uint32_t* bin_fill = bins;
        uint32_t* bin_base = (uint32_t*)((char*)bins + fill_bytes);
        uint32_t* bflags = (uint32_t*)((char*)bins + fill_bytes + base_bytes);
        uint32_t* bin_data = (uint32_t*)((char*)bins + fill_bytes + base_bytes + 256 + ((G * 4 + 255) & ~(size_t)255));
        B2R_CUDA(ctx, cudaMemsetAsync(bins, 0, fill_bytes + base_bytes + 256 + ((G * 4 + 255) & ~(size_t)255), st));
        { KTimer kt(ctx, "msm_scatter", (double)G * n);
        switch (HB) {
            case 7: k_bin_scatter<16, 7><<<gd, 256, 0, st>>>(scalars_dev, (uint32_t)n, (uint32_t)bs->n, idx_bits, bin_fill, bin_data, bflags); break;
            case 8: k_bin_scatter<16, 8><<<gd, 256, 0, st>>>(scalars_dev, (uint32_t)n, (uint32_t)bs->n, idx_bits, bin_fill, bin_data, bflags); break;
            case 9: k_bin_scatter<16, 9><<<gd, 256, 0, st>>>(scalars_dev, (uint32_t)n, (uint32_t)bs->n, idx_bits, bin_fill, bin_data, bflags); break;
            default: k_bin_scatter<16, 10><<<gd, 256, 0, st>>>(scalars_dev, (uint32_t)n, (uint32_t)bs->n, idx_bits, bin_fill, bin_data, bflags); break;
        }
        B2R_LAUNCH_CHECK(ctx);
        k_bin_prefix<<<(unsigned)G, BIN_HI_MAX, 0, st>>>(bin_fill, bin_base, off, B, HI);
        B2R_LAUNCH_CHECK(ctx);
        // function attributes are per device: set on every call (a process may hold one context per GPU)
        B2R_CUDA(ctx, cudaFuncSetAttribute(k_bin_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BIN_CAP * 4)));
        k_bin_sort<<<dim3(HI, (unsigned)G), BIN_SORT_T, BIN_CAP * 4, st>>>(bin_data, bin_fill, bin_base, idx_bits, B, HI, off, ent, key, ent_cap);
        B2R_LAUNCH_CHECK(ctx); }
        std::vector<uint32_t> hf(G);
        B2R_CUDA(ctx, cudaMemcpyAsync(hf.data(), bflags, G * 4, cudaMemcpyDeviceToHost, st));
        B2R_CUDA(ctx, cudaStreamSynchronize(st));
        sorted = true;
        for (uint32_t f : hf) sorted = sorted && f == 0;   // a bin overflowed (the scalars were not uniform): general path
    }
    if (!sorted) B2R_TRY(classic_sort());
    double entries_total = (double)G * n;
    if (ctx->profile) {   // actual number of bucket entries of this group (the sum of the last offsets), for the roofline
        std::vector<uint32_t> last(G);
        B2R_CUDA(ctx, cudaMemcpy2DAsync(last.data(), 4, off + B, (size_t)(B + 1) * 4, 4, G, cudaMemcpyDeviceToHost, st));
        B2R_CUDA(ctx, cudaStreamSynchronize(st));
        entries_total = 0;
        for (uint32_t v : last) entries_total += v;
    }
    { KTimer kt(ctx, "msm_accum_entries", entries_total);
    {
        const dim3 ga((nch1 + 127) / 128, (unsigned)G);
#define B2R_ACC(MINB, PF)                                                                                                        \
    do {                                                                                                                     \
        if (L1 == 56) k_accum_entries<56, MINB, PF><<<ga, 128, 0, st>>>(bs->table, ent, key, ent_cap, off, B, nch1, bk, ka, pa, slotsA);      \
        else if (L1 == 74) k_accum_entries<74, MINB, PF><<<ga, 128, 0, st>>>(bs->table, ent, key, ent_cap, off, B, nch1, bk, ka, pa, slotsA); \
        else k_accum_entries<MSM_L1, MINB, PF><<<ga, 128, 0, st>>>(bs->table, ent, key, ent_cap, off, B, nch1, bk, ka, pa, slotsA);           \
    } while (0)
        // measured on B200 (64 x 2^17 uniform scalars, tools/microbench.py): 4 CTAs/SM without the point prefetch 22.7 ms,
        // with it 24.1 ms (spills); 5 or 6 CTAs/SM (96 / 80 registers, spills) 23.5 - 24.0 ms.  Re-measured on the in-place
        // 1160-MAD addition (bench.py, accumulation ms per step): 4 CTAs 220.1, 5 CTAs (96 registers, 40 bytes of spills) 220.1
        switch (variant) {
            case 1: B2R_ACC(4, true); break;
            case 3:   // table points staged through a shared-memory ring by per-lane cp.async.bulk + mbarrier
                k_accum_entries_ring<MSM_L1, 4, true><<<ga, 128, 0, st>>>(bs->table, ent, key, ent_cap, off, B, nch1, bk, ka, pa, slotsA);
                break;
            case 4:   // the same ring filled by LDGSTS (cp.async) groups
                k_accum_entries_ring<MSM_L1, 4, false><<<ga, 128, 0, st>>>(bs->table, ent, key, ent_cap, off, B, nch1, bk, ka, pa, slotsA);
                break;
            default: B2R_ACC(4, false); break;
        }
#undef B2R_ACC
    } }
    B2R_LAUNCH_CHECK(ctx);
    if (sorted && G <= 4) {   // binned sort succeeded: every bucket spans a bounded number of chunks - one fixup launch
        // (a handful of vectors only: across a large batch the per-bucket loops diverge and the level scheme below is faster)
        { KTimer kt(ctx, "msm_accum_slots");
        const dim3 gf((B + 127) / 128, (unsigned)G);
        if (L1 == 56) k_bucket_fixup<56><<<gf, 128, 0, st>>>(off, B, pa, slotsA, bk);
        else if (L1 == 74) k_bucket_fixup<74><<<gf, 128, 0, st>>>(off, B, pa, slotsA, bk);
        else k_bucket_fixup<MSM_L1><<<gf, 128, 0, st>>>(off, B, pa, slotsA, bk); }
        B2R_LAUNCH_CHECK(ctx);
    } else {
    // upper levels: ping-pong slot lists until one chunk remains
    const uint32_t* ik = ka;
    const xyzz_t* ip = pa;
    size_t in_stride = slotsA;
    uint32_t M = (uint32_t)slotsA;
    bool to_b = true;
    for (;;) {
        uint32_t nch = (M + L2 - 1) / L2;
        uint32_t* ok = to_b ? kb : ka;
        xyzz_t* op = to_b ? pb : pa;
        size_t out_stride = to_b ? slotsB : slotsA;
        { KTimer kt(ctx, "msm_accum_slots");
        k_accum_slots<L2><<<dim3((nch + 127) / 128, (unsigned)G), 128, 0, st>>>(ik, ip, in_stride, M, nch, B, bk, ok, op, out_stride); }
        B2R_LAUNCH_CHECK(ctx);
        if (nch == 1) break;
        ik = ok;
        ip = op;
        in_stride = out_stride;
        M = 2 * nch;
        to_b = !to_b;
    }
    }
    const bool warp_reduce = nseg == 1024 && G * nseg < (size_t)ctx->sm_count * 128;   // c = 16, a handful of vectors
    { KTimer kt(ctx, "msm_bucket_reduce");
    static_assert(BR_PER1 == 32, "k_br_level1_warp maps one lane to one bucket of a segment");
    if (G * nseg < (size_t)ctx->sm_count * 128) {   // fewer level-1 threads than one CTA per SM: a warp per segment instead
        const uint32_t total_seg = (uint32_t)(G * nseg);
        k_br_level1_warp<<<(total_seg + 3) / 4, 128, 0, st>>>(bk, B, nseg, total_seg, s1, a1);
    } else {
        k_br_level1<<<dim3((nseg + 127) / 128, (unsigned)G), 128, 0, st>>>(bk, B, nseg, s1, a1);
    }
    B2R_LAUNCH_CHECK(ctx);
    const uint32_t total = (uint32_t)(G * nu);
    if (warp_reduce) k_br_level2_warp<<<(unsigned)((G * 64 + 3) / 4), 128, 0, st>>>(s1, a1, (uint32_t)(G * 32), l2);
    else k_br_level2<<<(total + 127) / 128, 128, 0, st>>>(s1, a1, nseg, nu, total, l2); }
    B2R_LAUNCH_CHECK(ctx);
    { KTimer kt(ctx, "msm_final");
    if (warp_reduce) k_br_level3_warp<<<(unsigned)G, 96, 0, st>>>(l2, out_dev);
    else k_br_level3<<<(unsigned)G, BR_T3, 0, st>>>(l2, nu, out_dev); }
    B2R_LAUNCH_CHECK(ctx);
    return 0;
}

static size_t msm_group_bytes(const b2r_bases* bs, size_t n) {
    const uint32_t W = bs->W, B = 1u << (bs->c - 1);
    size_t ent_cap = n * W;
    size_t nch1 = (ent_cap + 55) / 56 /* the shortest chunk length msm_group may pick */, slotsA = 2 * nch1, slotsB = 2 * ((slotsA + 15) / 16);
    return 3 * (size_t)B * 4 + ent_cap * 8 + (size_t)B * 128 + (slotsA + slotsB) * 132 + ((size_t)B / 16 + (size_t)B / 64 + 8) * 128 + 4096 * 8;   // (+ 10 - 20 MiB per vector in the second arena for the binned sort)
}

// `uniform`: the caller knows the non-zero scalars to be uniformly random field elements (no repeated values)
int32_t msm_batch_dev(b2r_ctx* ctx, const b2r_bases* bs, const fe_t* scalars_dev, size_t m, size_t n, affine_t* out_dev, bool uniform) {
    if (n > bs->n) return fail(ctx, B2R_ERR_INVALID, "msm: more scalars than registered bases");
    if (n == 0) {
        B2R_CUDA(ctx, cudaMemsetAsync(out_dev, 0, m * sizeof(affine_t), ctx->stream));
        return 0;
    }
    size_t per_vec = msm_group_bytes(bs, n);
    size_t budget = (size_t)16 << 30;  // work arena per group of vectors (fewer, larger launches: 180 GB of HBM)
    size_t G = budget / per_vec;
    if (G < 1) G = 1;
    if (G > 1024) G = 1024;
    for (size_t v = 0; v < m; v += G) {
        size_t g = (m - v < G) ? (m - v) : G;
        B2R_TRY(msm_group(ctx, bs, scalars_dev + v * n, g, n, out_dev + v, uniform));
    }
    return 0;
}

// registration from device-resident affine points (d_in must not live in the SC_MSM_B arena)
int32_t bases_register_dev(b2r_ctx* ctx, const affine_t* d_in, size_t n, b2r_bases** out, uint32_t window) {
    *out = nullptr;
    b2r_bases* bs = new b2r_bases();
    bs->n = n;
    bs->c = (window == 10 || window == 13 || window == 16) ? window : pick_window(n);
    bs->W = (255 + bs->c - 1) / bs->c;
    cudaError_t e = cudaMalloc(&bs->table, (size_t)bs->W * n * sizeof(affine_t));
    if (e != cudaSuccess) {
        delete bs;
        return cuda_fail(ctx, e, "cudaMalloc(base table)");
    }
    const uint32_t SLICE = 1u << 17;
    xyzz_t* tmp = nullptr;
    int32_t rc = scratch_get(ctx, SC_MSM_B, (size_t)SLICE * (bs->W - 1) * sizeof(xyzz_t), (void**)&tmp);
    if (rc) { cudaFree(bs->table); delete bs; return rc; }
    for (size_t i0 = 0; i0 < n; i0 += SLICE) {
        uint32_t cnt = (uint32_t)((n - i0 < SLICE) ? (n - i0) : SLICE);
        k_precompute<<<(cnt + 127) / 128, 128, 0, ctx->stream>>>(d_in, bs->table, tmp, (uint32_t)n, (uint32_t)i0, cnt, bs->c, bs->W);
        ctx->launches++;
    }
    e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { cudaFree(bs->table); delete bs; return cuda_fail(ctx, e, "k_precompute"); }
    *out = bs;
    return 0;
}

size_t bases_count(const b2r_bases* bs) { return bs ? bs->n : 0; }

void bases_destroy(b2r_bases* bs) {
    if (!bs) return;
    cudaFree(bs->table);
    delete bs;
}

// the same points again with another window width (table row 0 is the points themselves): sparse scalar vectors
// want few buckets (their bucket reduction costs more than their additions), dense ones few windows
int32_t bases_register_rewindowed(b2r_ctx* ctx, const b2r_bases* src, uint32_t window, b2r_bases** out) {
    return bases_register_dev(ctx, src->table, src->n, out, window);
}

// new base set S_j = sum_{i >= j} P_i of a registered set (see k_sfx_local)
int32_t bases_register_suffix_sums(b2r_ctx* ctx, const b2r_bases* src, b2r_bases** out) {
    *out = nullptr;
    const uint32_t n = (uint32_t)src->n, nchunks = (n + SFX_CHUNK - 1) / SFX_CHUNK;
    char* buf = nullptr;
    const size_t o_x = ((size_t)nchunks * sizeof(xyzz_t) + 255) & ~(size_t)255, o_aff = o_x + (size_t)n * sizeof(xyzz_t);
    B2R_CUDA(ctx, cudaMalloc(&buf, o_aff + (size_t)n * sizeof(affine_t)));
    xyzz_t* tot = (xyzz_t*)buf;
    xyzz_t* sx = (xyzz_t*)(buf + o_x);
    affine_t* sa = (affine_t*)(buf + o_aff);
    cudaStream_t st = ctx->stream;
    const uint32_t gb = (nchunks + 127) / 128;
    k_sfx_local<<<gb, 128, 0, st>>>(src->table, tot, n);
    k_sfx_scan<<<1, 32, 0, st>>>(tot, nchunks);
    k_sfx_write<<<gb, 128, 0, st>>>(src->table, tot, sx, n);
    k_normalize<<<(n + 127) / 128, 128, 0, st>>>(sx, sa, n);
    ctx->launches += 4;
    int32_t rc = bases_register_dev(ctx, sa, n, out, 0);  // synchronises the stream
    cudaFree(buf);
    return rc;
}

}  // namespace b2r

using namespace b2r;

extern "C" {

int32_t b2r_bases_register(b2r_ctx* ctx, const b2r_g1_affine* bases_host, size_t n, b2r_bases** out) try {
    B2R_ENTER(ctx);
    if (!bases_host || !out || n == 0) return fail(ctx, B2R_ERR_INVALID, "bases_register: bad argument");
    if (n > ((size_t)1 << 26)) return fail(ctx, B2R_ERR_INVALID, "bases_register: n > 2^26");
    *out = nullptr;
    affine_t* d_in = nullptr;
    B2R_TRY(scratch_get(ctx, SC_STAGE, n * sizeof(affine_t), (void**)&d_in));
    B2R_CUDA(ctx, cudaMemcpyAsync(d_in, bases_host, n * sizeof(affine_t), cudaMemcpyHostToDevice, ctx->stream));
    return bases_register_dev(ctx, d_in, n, out, 0);
} B2R_ABI_CATCH(ctx)

int32_t b2r_bases_download(b2r_ctx* ctx, const b2r_bases* bases, b2r_g1_affine* out_host, size_t n) try {
    B2R_ENTER(ctx);
    if (!bases || !out_host || n > bases->n) return fail(ctx, B2R_ERR_INVALID, "bases_download: bad argument");
    B2R_CUDA(ctx, cudaMemcpyAsync(out_host, bases->table, n * sizeof(affine_t), cudaMemcpyDeviceToHost, ctx->stream));
    B2R_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
} B2R_ABI_CATCH(ctx)

int32_t b2r_bases_free(b2r_ctx* ctx, b2r_bases* bases) try {
    B2R_ENTER(ctx);
    if (!ctx || !bases) return B2R_ERR_INVALID;
    cudaStreamSynchronize(ctx->stream);
    cudaFree(bases->table);
    delete bases;
    return 0;
} B2R_ABI_CATCH(ctx)

int32_t b2r_msm_g1_batch_dev(b2r_ctx* ctx, const b2r_bases* bases, const b2r_fr* scalars_dev, size_t m, size_t n,
                             b2r_g1_affine* out_dev) try {
    B2R_ENTER(ctx);
    if (!bases || !out_dev || (!scalars_dev && n)) return fail(ctx, B2R_ERR_INVALID, "msm: null pointer");
    if (n > bases->n) return fail(ctx, B2R_ERR_INVALID, "msm: more scalars than registered bases");
    if (m == 0) return 0;
    return msm_batch_dev(ctx, bases, (const fe_t*)scalars_dev, m, n, (affine_t*)out_dev, false);
} B2R_ABI_CATCH(ctx)

int32_t b2r_msm_g1_batch_dev_ex(b2r_ctx* ctx, const b2r_bases* bases, const b2r_fr* scalars_dev, size_t m, size_t n, uint32_t flags,
                                b2r_g1_affine* out_dev) try {
    B2R_ENTER(ctx);
    if (!bases || !out_dev || (!scalars_dev && n)) return fail(ctx, B2R_ERR_INVALID, "msm: null pointer");
    if (flags & ~B2R_MSM_UNIFORM) return fail(ctx, B2R_ERR_INVALID, "msm: unknown flag");
    if (n > bases->n) return fail(ctx, B2R_ERR_INVALID, "msm: more scalars than registered bases");
    if (m == 0) return 0;
    return msm_batch_dev(ctx, bases, (const fe_t*)scalars_dev, m, n, (affine_t*)out_dev, (flags & B2R_MSM_UNIFORM) != 0);
} B2R_ABI_CATCH(ctx)

int32_t b2r_msm_g1_batch(b2r_ctx* ctx, const b2r_bases* bases, const b2r_fr* scalars, size_t m, size_t n,
                         b2r_g1_affine* out) try {
    B2R_ENTER(ctx);
    if (!bases || !out || (!scalars && n)) return fail(ctx, B2R_ERR_INVALID, "msm: null pointer");
    if (n > bases->n) return fail(ctx, B2R_ERR_INVALID, "msm: more scalars than registered bases");
    if (m == 0) return 0;
    char* d = nullptr;
    size_t sb = m * n * sizeof(fe_t);
    size_t sb_al = (sb + 255) & ~(size_t)255;
    B2R_TRY(scratch_get(ctx, SC_STAGE, sb_al + m * sizeof(affine_t), (void**)&d));
    if (sb) B2R_CUDA(ctx, cudaMemcpyAsync(d, scalars, sb, cudaMemcpyHostToDevice, ctx->stream));
    B2R_TRY(msm_batch_dev(ctx, bases, (const fe_t*)d, m, n, (affine_t*)(d + sb_al), false));
    B2R_CUDA(ctx, cudaMemcpyAsync(out, d + sb_al, m * sizeof(affine_t), cudaMemcpyDeviceToHost, ctx->stream));
    B2R_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
} B2R_ABI_CATCH(ctx)

int32_t b2r_msm_g1(b2r_ctx* ctx, const b2r_bases* bases, const b2r_fr* scalars, size_t n, b2r_g1* out) try {
    B2R_ENTER(ctx);
    if (!out) return fail(ctx, B2R_ERR_INVALID, "msm: null pointer");
    b2r_g1_affine a;
    B2R_TRY(b2r_msm_g1_batch(ctx, bases, scalars, 1, n, &a));
    bool ident = true;
    for (int i = 0; i < 4; i++) ident = ident && a.x.l[i] == 0 && a.y.l[i] == 0;
    fe_t one = Fq::one();
    b2r_fq one64;
    for (int i = 0; i < 4; i++) one64.l[i] = (uint64_t)one.l[2 * i] | ((uint64_t)one.l[2 * i + 1] << 32);
    if (ident) {  // halo2curves G1::identity() = (0, 1, 0)
        for (int i = 0; i < 4; i++) out->x.l[i] = 0, out->z.l[i] = 0;
        out->y = one64;
    } else {
        out->x = a.x;
        out->y = a.y;
        out->z = one64;
    }
    return 0;
} B2R_ABI_CATCH(ctx)

}  // extern "C"
