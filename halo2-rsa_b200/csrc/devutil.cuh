// Small device helpers shared by the translation units: 128-bit field-element loads/stores and
// the seeded blinding stream that stands in for halo2's OsRng on both the product and the oracle
// (oracle/plonk.py: blind_fe restates it).
#pragma once
#include "field.cuh"

namespace b2r {

__device__ __forceinline__ fe_t ldv(const fe_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    fe_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ fe_t ldv_nc(const fe_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    fe_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void stv(fe_t* p, const fe_t& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

// blinding streams: one per polynomial the prover blinds (ids shared with oracle/plonk.py)
enum BlindStream : uint32_t {
    ST_ADVICE = 0,       // + advice column
    ST_LOOKUP_A = 8,     // + lookup index
    ST_LOOKUP_S = 16,
    ST_LOOKUP_Z = 24,
    ST_PERM_Z = 32,      // + permutation set
    ST_RANDOM_POLY = 40
};

B2R_HD uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// splitmix64 over (seed, proof, stream, row), 254 bits, one conditional subtraction of r; the limbs
// are taken as a Montgomery representation (a uniform field element either way)
B2R_HD fe_t blind_value(uint64_t seed, uint32_t proof, uint32_t stream, uint32_t row) {
    fe_t r;
    uint64_t base = splitmix64(seed ^ splitmix64(((uint64_t)proof << 40) | ((uint64_t)stream << 28) | row));
    for (int j = 0; j < 4; j++) {
        uint64_t w = splitmix64(base + j);
        r.l[2 * j] = (uint32_t)w;
        r.l[2 * j + 1] = (uint32_t)(w >> 32);
    }
    r.l[7] &= 0x3fffffffu;
    Fr::final_sub(r.l);
    return r;
}

}  // namespace b2r
