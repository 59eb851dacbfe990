/*
 * gm_rng.cuh -- counter-based Philox4x32-10 streams keyed by photon identity.
 *
 * Replaces the reference's process-global mt19937 (cuda_grmonty/monty_rand.cpp:19-31) and the slot-keyed
 * cuRAND states of the reference GPU build (super_photon.cu:1039-1043): the stream of a photon depends only
 * on (seed, photon identity), never on launch geometry or GPU count.
 *   counter = (draw index, id0, id1, id2), key = 64-bit seed
 *   primary photon i : id = (i_lo, i_hi, 0)
 *   zone z rounding  : id = (z_lo, z_hi, 0x40000000)
 *   scattered photon : id = 3 words drawn from the parent's stream, top bit of id2 set
 */
#pragma once
#include <cstdint>
#include "gm_params.h"

namespace gm {

struct Rng {
    uint32_t id0, id1, id2, ctr;
};

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

__device__ __forceinline__ Rng rng_primary(uint64_t idx) {
    Rng r;
    r.id0 = (uint32_t)idx;
    r.id1 = (uint32_t)(idx >> 32);
    r.id2 = 0u;
    r.ctr = 0u;
    return r;
}

__device__ __forceinline__ Rng rng_zone(uint64_t z) {
    Rng r;
    r.id0 = (uint32_t)z;
    r.id1 = (uint32_t)(z >> 32);
    r.id2 = 0x40000000u;
    r.ctr = 0u;
    return r;
}

/* uniform in the open interval (0,1): (53-bit integer + 0.5) * 2^-53 */
__device__ __forceinline__ double rng_uniform(const GmParams &P, Rng &r) {
    uint32_t o[4];
    philox4x32_10(r.ctr, r.id0, r.id1, r.id2, P.seed_lo, P.seed_hi, o);
    r.ctr += 1u;
    const uint64_t bits = ((uint64_t)o[1] << 32) | o[0];
    return ((double)(bits >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ Rng rng_child(const GmParams &P, Rng &parent) {
    uint32_t o[4];
    philox4x32_10(parent.ctr, parent.id0, parent.id1, parent.id2, P.seed_lo, P.seed_hi, o);
    parent.ctr += 1u;
    Rng c;
    c.id0 = o[0];
    c.id1 = o[1];
    c.id2 = o[2] | 0x80000000u;
    c.ctr = 0u;
    return c;
}

/* chi-square with 3..6 degrees of freedom from uniforms (exact; replaces reference monty_rand.cpp:28-31):
 * a sum of dof/2 exponentials (-2 ln U) plus, for odd dof, one squared normal by Box-Muller. */
__device__ __forceinline__ double rng_chi_sq(const GmParams &P, Rng &r, int dof) {
    double s = 0.0;
    const int n_exp = dof >> 1;
    for (int i = 0; i < n_exp; ++i)
        s += -2.0 * log(rng_uniform(P, r));
    if (dof & 1) {
        const double ua = rng_uniform(P, r);
        const double ub = rng_uniform(P, r);
        const double c = cospi(2.0 * ub);
        s += -2.0 * log(ua) * c * c;
    }
    return s;
}

} /* namespace gm */
