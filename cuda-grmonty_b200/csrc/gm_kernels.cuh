/*
 * gm_kernels.cuh -- __global__ entry points of the B200 transport path.
 *
 *   zone_kernel       per-zone emission data: photon count, dn_max, fluid state, tetrad   (once per context)
 *   birth_kernel      make_super_photon for one generation of primaries -> photon queue   (per generation)
 *   transport_kernel  persistent loop: pop / track / scatter stage / deferred record stage (per generation)
 *                     (default geometry one warp per block, eight blocks per SM; gm_wavefront.cuh and gm_pipeline.cuh
 *                     hold the two optional kernels built on the same device functions)
 *
 * Reference functions restated: init_zone harm_model.cpp:1337-1389, get_zone :673-704,
 * sample_zone_photon :706-782, linear_interp_weight :784-792, run_simulation CPU loop :366-404.
 */
#pragma once
#include "gm_transport.cuh"

namespace gm {

/* per-zone record used by the birth kernel (AoS: all photons of a zone read the same record) */
struct ZoneData {
    double n_e, theta_e, b, dn_max;
    double e_con[4][4];
    double e_cov_t[4]; /* e_cov[a][0], a = 0..3: gives e = -k_t */
    double e_cov_p[4]; /* e_cov[a][3]: gives l = k_phi */
};

/* reference init_zone (+ get_zone's stochastic rounding, keyed by zone id instead of the global stream) */
__global__ void zone_kernel(GmParams P, ZoneData *zones, double *nz_out, long long *num_to_gen) {
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    if (z >= P.n0 * P.n1)
        return;
    const int i = z / P.n1, j = z % P.n1;
    double x[4];
    MetricCov g;
    Fluid f;
    fluid_zone(P, i, j, x, g, f);
    ZoneData zd;
    zd.n_e = f.n_e;
    zd.theta_e = f.theta_e;
    zd.b = f.b;
    zd.dn_max = 0.0;
    double nz = 0.0;
    bool emits = !(f.n_e == 0.0 || f.theta_e < kThetaEMin);
    if (emits) {
        const double l_bth = log(f.b * f.theta_e * f.theta_e);
        double d_l = (l_bth - P.l_b_min) / P.d_l_b;
        const int l = (int)d_l;
        d_l -= l;
        double ninterp = 0.0, dn_max = 0.0;
        if (l < 0) {
            emits = false;
        } else if (l >= kNint) {
            /* out-of-table branch incl. the reference's index slip (:1358-1369, Appendix A.9) */
            for (int s = 0; s <= kNESamp; ++s) {
                const double dn =
                    f_eval(P, f.theta_e, f.b, exp(j * P.d_l_nu + P.l_nu_min)) / (exp(P.weight[s]) + 1.0e-100);
                dn_max = fmax(dn_max, dn);
                ninterp += P.d_l_nu * dn;
            }
        } else if (!isinf(P.nint[l]) && !isinf(P.nint[l + 1])) {
            ninterp = exp((1.0 - d_l) * P.nint[l] + d_l * P.nint[l + 1]);
            dn_max = exp((1.0 - d_l) * P.dndlnu_max[l] + d_l * P.dndlnu_max[l + 1]);
        }
        if (emits) {
            const double k2 = k2_eval(P, f.theta_e);
            if (k2 == 0.0) {
                emits = false;
            } else {
                nz = P.geom_det[z] * f.n_e * f.b * f.theta_e * f.theta_e * ninterp / k2;
                if (nz > P.nz_max) {
                    nz = 0.0;
                    emits = false;
                }
                zd.dn_max = emits ? dn_max : 0.0;
            }
        }
    }
    if (!emits)
        nz = 0.0;
    Rng r = rng_zone((uint64_t)z);
    const double u = rng_uniform(P, r);
    long long n = (long long)nz;
    if (fmod(nz, 1.0) > u)
        n += 1;
    num_to_gen[z] = n;
    if (nz_out)
        nz_out[z] = nz;
    /* tetrad of the zone (reference sample_zone_photon :717-731) */
    double b_hat[4];
    if (f.b > 0.0) {
#pragma unroll
        for (int d = 0; d < 4; ++d)
            b_hat[d] = f.b_con[d] * P.b_unit / f.b;
    } else {
        b_hat[0] = 1.0;
        b_hat[1] = b_hat[2] = b_hat[3] = 0.0;
    }
    double e_cov[4][4];
    make_tetrad(g, f.u_con, b_hat, zd.e_con, e_cov);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        zd.e_cov_t[a] = e_cov[a][0];
        zd.e_cov_p[a] = e_cov[a][3];
    }
    zones[z] = zd;
}

/* reference linear_interp_weight, harm_model.cpp:784-792 */
__device__ __forceinline__ double linear_interp_weight(const GmParams &P, double nu) {
    return interp_exp_table(P.weight, log(nu), P.l_nu_min, P.d_l_nu);
}

/* birth state of the primary with global index idx (reference sample_zone_photon :733-781) */
struct Birth {
    double x[4], k[4], w, e, l, n_e, theta_e, b;
    Rng rng;
};

__device__ __forceinline__ void make_primary(const GmParams &P, const ZoneData *zones, const long long *prefix,
                                             long long idx, Birth &B) {
    /* zone z with prefix[z] <= idx < prefix[z+1] */
    long long lo = 0, hi = (long long)P.n0 * P.n1;
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (__ldg(prefix + mid) <= idx)
            lo = mid;
        else
            hi = mid;
    }
    const int i = (int)(lo / P.n1), j = (int)(lo % P.n1);
    const ZoneData *zd = zones + lo;
    const double n_e = __ldg(&zd->n_e), theta_e = __ldg(&zd->theta_e), b = __ldg(&zd->b);
    const double dn_max = __ldg(&zd->dn_max);
    Rng r = rng_primary((uint64_t)idx);
    double nu, weight;
    do {
        nu = exp(rng_uniform(P, r) * P.n_l_n + P.l_nu_min);
        weight = linear_interp_weight(P, nu);
    } while (rng_uniform(P, r) > (f_eval(P, theta_e, b, nu) / (weight + 1.0e-100)) / dn_max);
    const double j_max = synch_sin(P, nu, n_e, theta_e, b, 1.0);
    double cos_th, sin_th;
    do {
        cos_th = 2.0 * rng_uniform(P, r) - 1.0;
        sin_th = sqrt(1.0 - cos_th * cos_th);
    } while (rng_uniform(P, r) > (synch_sin(P, nu, n_e, theta_e, b, sin_th) / j_max));
    double sin_phi, cos_phi;
    sincospi(2.0 * rng_uniform(P, r), &sin_phi, &cos_phi);
    const double e = nu * kHPL / (kME * kCL * kCL);
    double kt[4] = {e, e * cos_th, e * sin_th * cos_phi, e * sin_th * sin_phi};
#pragma unroll
    for (int m = 0; m < 4; ++m)
        B.k[m] = __ldg(&zd->e_con[0][m]) * kt[0] + __ldg(&zd->e_con[1][m]) * kt[1] +
                 __ldg(&zd->e_con[2][m]) * kt[2] + __ldg(&zd->e_con[3][m]) * kt[3];
    kt[0] = -kt[0];
    B.e = -(__ldg(&zd->e_cov_t[0]) * kt[0] + __ldg(&zd->e_cov_t[1]) * kt[1] + __ldg(&zd->e_cov_t[2]) * kt[2] +
            __ldg(&zd->e_cov_t[3]) * kt[3]);
    B.l = __ldg(&zd->e_cov_p[0]) * kt[0] + __ldg(&zd->e_cov_p[1]) * kt[1] + __ldg(&zd->e_cov_p[2]) * kt[2] +
          __ldg(&zd->e_cov_p[3]) * kt[3];
    B.x[0] = 0.0;
    B.x[1] = P.x_start1 + (i + 0.5) * P.dx1;
    B.x[2] = P.x_start2 + (j + 0.5) * P.dx2;
    B.x[3] = 0.0;
    B.w = weight;
    B.n_e = n_e;
    B.theta_e = theta_e;
    B.b = b;
    B.rng = r;
}

/* Processing order: position j of a run handles primary (j * mult) mod total, a Weyl sequence with
 * mult ~ total / golden ratio coprime to total, so that every contiguous range of positions samples all
 * emission zones evenly and the bias statistics frozen at a generation start are representative.  The photon
 * itself (Philox stream, birth zone) depends only on its primary index. */
__host__ __device__ __forceinline__ long long permute_position(long long j, long long mult, long long total) {
    return (long long)(((unsigned __int128)(unsigned long long)j * (unsigned long long)mult) %
                       (unsigned long long)total);
}

/* Complete start-of-track record of a photon born at x with wave-vector k (reference track_super_photon
 * :902-915 -- done here, by the producer, with all lanes active, instead of by the consumer). */
__device__ __forceinline__ void store_new_photon(const TransportArgs &A, unsigned int slot, const double x[4],
                                                 const double k[4], double w, double e, double x1i, double x2i,
                                                 double n_e_0, double theta_e_0, double b_0, double e_0,
                                                 int n_scatt, const Rng &rng, int clock0, int ns_bits = -1) {
    /* ns_bits >= 0 (pipelined generations): generation tag bits of the record; the bias is left to the pick-up
     * (kFreshBit), because the generation's statistics are not frozen yet when its primaries are born */
    const GmParams &P = A.P;
    const GeoPoint q = geo_point(P, x[1], x[2]);
    const MetricCov g = metric_cov(P, q);
    Fluid f;
    fluid_params(P, x[1], x[2], g, q, f);
    const bool lazy = ns_bits >= 0;
    const TrackInit t = track_init(P, A.bias.bias_den, k, w, f, lazy);
    Connection c;
    connection_eval(P, q, c);
    double dk[4];
    geodesic_rhs(c, k, dk);
    pool_store_hot(A.pool, slot, x, k, dk, w, e, 0.0, 0.0, t, rng, lazy ? (ns_bits | (t.ne_pos ? kFreshBit : 0)) : 0);
    pstore(A.pool, P_E, slot, e);
    pstore(A.pool, P_X1I, slot, x1i);
    pstore(A.pool, P_X2I, slot, x2i);
    pstore(A.pool, P_NE0, slot, n_e_0);
    pstore(A.pool, P_TE0, slot, theta_e_0);
    pstore(A.pool, P_B0, slot, b_0);
    pstore(A.pool, P_E0, slot, e_0);
    __stcg(A.pool.n_scatt + slot, n_scatt);
    __stcg(A.pool.gclock + slot, clock0);
}

/* positions first, first + stride, ... (count of them) -> pool slots / ready-queue entries 0..count-1 */
/* `spread` > 0: the t-th of the batch's primaries starts its lineage clock at -(count - 1 - t) / spread, i.e. may
 * make that many attempts on top of the generation's budget -- lanes pick the primaries up in order, so a lineage that
 * starts early has the rest of the generation to run without delaying its end */
/* `order` (optional): the batch's primary indices sorted by the expected lifetime of their birth zone, longest first.
 * Lanes take pool slots in order, so the long-lived lineages start first and the generation's drain is short; the
 * results do not depend on the order (statistics are frozen within a generation). */
/* `slot0`, `ns_bits` >= 0 (pipelined generations): the records go to pool slots slot0 .. slot0 + count - 1 with the
 * generation tag in their n_step word, and there are no queue entries (lanes take them by index) */
__global__ void birth_kernel(TransportArgs A, const ZoneData *zones, const long long *prefix, long long first,
                             long long stride, long long count, long long mult, long long total, long long spread,
                             const long long *order, unsigned int slot0 = 0u, int ns_bits = -1, int clock_base = 0) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < count;
         t += (long long)gridDim.x * blockDim.x) {
        Birth B;
        make_primary(A.P, zones, prefix, order ? order[t] : permute_position(first + t * stride, mult, total), B);
        const int clock0 = clock_base + (spread > 0 ? -(int)((count - 1 - t) / spread) : 0);
        store_new_photon(A, slot0 + (unsigned int)t, B.x, B.k, B.w, B.e, B.x[1], B.x[2], B.n_e, B.theta_e, B.b, B.e, 0,
                         B.rng, clock0, ns_bits);
        if (ns_bits < 0)
            A.ready.entries[t] = (unsigned int)t + 1u;
    }
}

/* sort keys of a batch: key[t] = rank of the birth zone's radial bin (0 = longest-lived), val[t] = primary index */
__global__ void order_key_kernel(const long long *prefix, int n0, int n1, const unsigned char *bin_rank,
                                 long long first, long long stride, long long count, long long mult, long long total,
                                 unsigned char *keys, long long *vals) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= count)
        return;
    const long long idx = permute_position(first + t * stride, mult, total);
    long long lo = 0, hi = (long long)n0 * n1;
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (__ldg(prefix + mid) <= idx)
            lo = mid;
        else
            hi = mid;
    }
    keys[t] = bin_rank[(int)(lo / n1)];
    vals[t] = idx;
}

/* copy carried records between pools: dst slot (dst0 + i) <- src slot (list ? list[i] - 1 : src0 + i);
 * when `ready` is given the destination slots are also published on that queue at position dst0 + i */
/* keep_clock (pipelined generations): the record's own clock is copied (suspend_photon_pipe set it) and the queue
 * entries are written from index 0 (the limbo queue of the next launch's first generation) */
__global__ void carry_copy_kernel(PhotonPool dst, unsigned int dst0, PhotonPool src, unsigned int src0,
                                  const unsigned int *list, unsigned int n, unsigned int *ready_entries, int clock0,
                                  bool keep_clock = false) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const unsigned int s = list ? list[i] - 1u : src0 + i, d = dst0 + i;
    for (int f = 0; f < P_NFIELDS; ++f)
        dst.f[(size_t)f * dst.capacity + d] = src.f[(size_t)f * src.capacity + s];
    dst.rng[d] = src.rng[s];
    dst.n_scatt[d] = src.n_scatt[s];
    dst.n_step[d] = src.n_step[s];
    dst.gclock[d] = keep_clock ? src.gclock[s] : clock0;
    if (ready_entries)
        ready_entries[keep_clock ? i : d] = d + 1u;
}

/* counters[0] += created, in stream order (the host counts primaries; harm_model.cpp:395) */
__global__ void add_u64_kernel(unsigned long long *p, unsigned long long v) { *p += v; }

/* ---- the persistent transport kernel ------------------------------------------------------------------- */
extern __shared__ double gm_smem[];

__device__ __noinline__ void record_call(const TransportArgs *Ag, unsigned int slot, double x2, double x3, double w,
                                         double tau_abs, double tau_scatt) {
    record_super_photon(*Ag, slot, x2, x3, w, tau_abs, tau_scatt);
}

/* lifetime statistics of primaries by birth radius (cold: once per photon) */
__device__ __noinline__ void cost_call(const TransportArgs *Ag, unsigned int slot, int n_step) {
    const TransportArgs &A = *Ag;
    if (!A.zone_cost || __ldcg(A.pool.n_scatt + slot) != 0)
        return;
    int i = (int)((pload(A.pool, P_X1I, slot) - A.P.x_start1) * A.P.inv_dx1);
    i = max(0, min(i, A.P.n0 - 1));
    atomicAdd(A.zone_cost + i, (unsigned long long)n_step);
    atomicAdd(A.zone_cost + A.P.n0 + i, 1ull);
}

/* iterations between two block-wide synchronisations (measured, ms per step at configs[1]: 8 -> 673, 16 -> 660,
 * 24 -> 658, 32 -> 657; with the later kernel 16 -> 640, 32 -> 636; profiles/r1_ab_microopts.txt) */
#ifndef GM_ITERS_PER_SYNC
#define GM_ITERS_PER_SYNC 32
#endif
constexpr int kItersPerSync = GM_ITERS_PER_SYNC;

/* Record stage compaction.  A photon that ends is recorded (12 RED.F64 + counters, after 7 loads of its cold record
 * fields and a logarithm) and its lifetime is booked for the issue order: ~3 us of a warp's time when ONE lane does
 * it while 31 wait -- 5.7 % of the transport kernel's stall samples (ncu, round 2).  So the lane only leaves what the
 * record needs in shared memory (5 doubles, slot, n_step) and takes its next photon; when kRecordBatch lanes of the
 * warp have one pending -- or the warp has nothing else to do, or a lane ends a second photon first -- they record
 * together: one call, the loads in parallel, and __match_any_sync combines the lanes that hit the same bin. */
#ifndef GM_RECORD_BATCH
#define GM_RECORD_BATCH 12
#endif
constexpr int kRecordBatch = GM_RECORD_BATCH; /* 0: record at once (round 1) */
/* (Measured and rejected, profiles/r2_ab_microopts.txt: the same deferral for the publication of a photon parked for
 * the scattering stage -- one fence and one fetch-add per warp for several lanes -- costs 7 % at configs[1] and 20 % at
 * configs[0]: scattered photons are on the critical path of their lineage, recorded ones are not.) */
constexpr int kTransportSmemRows = 13 + 6; /* snapshot rows + pending-record rows per thread */
/* Per-photon debug output (end state and status bits of the test exports' batches, TransportArgs::D) is compiled into
 * the TEST library only: in the product it is ~60 instructions of the loop body that never run -- and the loop body
 * sits at the edge of the instruction cache, where every KB shows (profiles/r2_ab_microopts.txt). */
/* The per-lane loop-iteration counters behind grmonty_b200_stats::n_live_iterations / n_slot_iterations (lane
 * occupancy) likewise: two registers and two instructions per iteration, 1 % of the run time; the test library and
 * the tools that read them (GRMONTY_B200_LIB=...libgrmonty_b200_test.so) have them, the product reports 0. */
#ifndef GM_OCC_COUNTERS
#ifdef GRMONTY_B200_TEST_EXPORTS
#define GM_OCC_COUNTERS 1
#else
#define GM_OCC_COUNTERS 0
#endif
#endif
#ifndef GM_DEBUG_OUT
#ifdef GRMONTY_B200_TEST_EXPORTS
#define GM_DEBUG_OUT 1
#else
#define GM_DEBUG_OUT 0
#endif
#endif

template <int BLOCK, int MIN_BLOCKS>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS) transport_kernel(const TransportArgs A) {
    /* Control words shared by the block.  (Default geometry since round 2: one warp per block, eight blocks per SM;
     * with larger blocks the warps of a block re-align at the barrier of every outer iteration -- see DESIGN.md,
     * "Block lockstep".) */
    __shared__ unsigned long long s_base; /* first scatter-queue position claimed for the block */
    __shared__ int s_count;               /* number of scatter-queue entries claimed (0: no service) */
    __shared__ int s_quit;
    const int lane = threadIdx.x & 31;
    double *snap = gm_smem + threadIdx.x; /* 13 rows of BLOCK doubles */
    double *pend = gm_smem + (size_t)13 * BLOCK + threadIdx.x; /* 6 rows: x2 x3 w tau_abs tau_scatt (slot, n_step) */
    int pending = 0; /* bit 0: lifetime to be booked, bit 1: record to be made (data in `pend`) */
    /* the lanes that have something pending do it together */
    auto flush_pending = [&]() {
        if (pending) {
            const int2 sn = *reinterpret_cast<const int2 *>(pend + 5 * BLOCK);
            if (pending & 1)
                cost_call(A.self, (unsigned int)sn.x, sn.y);
            if (pending & 2)
                record_call(A.self, (unsigned int)sn.x, pend[0], pend[BLOCK], pend[2 * BLOCK], pend[3 * BLOCK],
                            pend[4 * BLOCK]);
            pending = 0;
        }
    };
    Live L;
    bool has = false;
    long long ticket = -1; /* position in the ready queue this lane is entitled to (monotone queue, no wrap) */
    Work wk = {0u, 0u, 0u, 0u, 0u, 0u, 0u};
    unsigned int idle_spins = 0;
    bool was_idle = true;

    for (;;) {
        int n_done = 0;
        /* ---- block control (one thread): claim parked photons for the scattering stage ---- */
        if (threadIdx.x == 0) {
            int cnt = 0;
            const unsigned long long h = ld_volatile_u64(A.scatter.head);
            unsigned long long t = ld_volatile_u64(A.scatter.tail);
            t = t < A.scatter.capacity ? t : A.scatter.capacity;
            const unsigned long long avail = t > h ? t - h : 0ull;
            /* a full block-load of parked photons, or -- when the block has nothing else to do -- any */
            if (avail >= (unsigned long long)BLOCK || (was_idle && avail > 0)) {
                cnt = avail < (unsigned long long)BLOCK ? (int)avail : BLOCK;
                if (atomicCAS(A.scatter.head, h, h + cnt) == h)
                    s_base = h;
                else
                    cnt = 0;
            }
            s_count = cnt;
        }
        __syncthreads();
        /* ---- scattering stage, all lanes of the block that got a parked photon ---- */
        {
            const int cnt = s_count;
            if ((int)threadIdx.x < cnt) {
                const unsigned long long pos = s_base + threadIdx.x;
                unsigned int v, spins = 0;
                while ((v = ld_volatile_u32(A.scatter.entries + pos)) == 0u) {
                    if (++spins > (1u << 26)) {
                        atomicOr(A.A.error, 2u);
                        break;
                    }
                }
                __threadfence();
                if (v) {
                    const ScatterStageResult sr = scatter_stage(A.self, v - 1u);
                    n_done += sr.done;
                    wk.attempts += sr.attempts;
                    wk.scatters += sr.scatters;
                    wk.tracked += sr.children;
                }
            }
        }
        /* ---- kItersPerSync loop iterations per warp between two block barriers.  The barrier keeps the warps
         *      of the block in the same code region (instruction cache); syncing only every few iterations
         *      keeps one warp's rare slow paths (record, park, suspend) from stalling the other warps every
         *      time ---- */
#pragma unroll 1
        for (int sub = 0; sub < kItersPerSync; ++sub) {
            /* refill empty lanes from the ready queue: a ticket per empty lane (one atomicAdd per warp, never
             * fails), then loads only */
            {
                const bool want = !has && ticket < 0;
                const unsigned int need = __ballot_sync(0xffffffffu, want);
                if (need) {
                    unsigned long long base = 0;
                    if (lane == 0)
                        base = atomicAdd(A.ready.head, (unsigned long long)__popc(need));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (want)
                        ticket = (long long)(base + __popc(need & ((1u << lane) - 1u)));
                }
                if (!has && ticket >= 0 && ticket < (long long)A.ready.capacity) {
                    /* acquire: pairs with the producer's fence + store in queue_push (no separate membar here) */
                    const unsigned int v = ld_acquire_u32(A.ready.entries + ticket);
                    if (v) {
                        const unsigned int slot = v - 1u;
                        ticket = -1;
                        live_load(A, slot, L);
                        bool bad = (L.w == 0.0);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            bad = bad || isnan(L.x[i]) || isnan(L.k[i]);
                        if (bad) {
                            ++n_done; /* invalid photon (reference :895-900): dropped */
#if GM_DEBUG_OUT
                            if (A.D.status && slot < A.D.n)
                                atomicOr(A.D.status + slot, 4);
#endif
                        } else {
                            has = true;
                        }
                    }
                }
            }
            const unsigned int live_mask = __ballot_sync(0xffffffffu, has);
            if (!live_mask)
                break; /* nothing to do in this warp: go to the barrier */
            /* one flattened iteration for every live lane */
#if GM_OCC_COUNTERS
            ++wk.slot_iters;
#endif
            if (has) {
#if GM_OCC_COUNTERS
                ++wk.live_iters;
#endif
                bool record;
                const StepResult r = advance(A, L, live_mask, snap, BLOCK, wk, record);
                if (r == STEP_FINISHED) {
                    if (kRecordBatch > 0) {
                        if (pending)
                            flush_pending(); /* rare: a second photon ended before the first was recorded */
                        pend[0] = L.x[2];
                        pend[BLOCK] = L.x[3];
                        pend[2 * BLOCK] = L.w;
                        pend[3 * BLOCK] = L.tau_abs;
                        pend[4 * BLOCK] = L.tau_scatt;
                        *reinterpret_cast<int2 *>(pend + 5 * BLOCK) = make_int2((int)L.slot, L.n_step);
                        pending = 1 | (record ? 2 : 0);
                    } else {
                        cost_call(A.self, L.slot, L.n_step);
                        if (record)
                            record_call(A.self, L.slot, L.x[2], L.x[3], L.w, L.tau_abs, L.tau_scatt);
                    }
                    if (record)
                        L.status |= 1;
#if GM_DEBUG_OUT
                    if (A.D.final_state && L.slot < A.D.n) {
                        double *o = A.D.final_state + (size_t)L.slot * 12;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            o[i] = L.x[i];
                            o[4 + i] = L.k[i];
                        }
                        o[8] = L.w;
                        o[9] = L.tau_abs;
                        o[10] = L.tau_scatt;
                        o[11] = L.e_0_s;
                        atomicOr(A.D.status + L.slot, L.status);
                        A.pool.rng[L.slot] = make_uint4(L.rng.id0, L.rng.id1, L.rng.id2, L.rng.ctr);
                    }
#endif
                    has = false;
                    ++n_done;
                } else if (r == STEP_SCATTER) {
                    has = false; /* parked for the scattering stage */
                } else if (r == STEP_SUSPEND) {
                    suspend_photon(A.self, L.slot, L.x[0], L.x[1], L.x[2], L.x[3], L.k[0], L.k[1], L.k[2], L.k[3],
                                   L.dk[0], L.dk[1], L.dk[2], L.dk[3], L.w, L.e_0_s, L.tau_abs, L.tau_scatt,
                                   L.alpha_scatt, L.alpha_abs, L.bi, L.ne_pos, L.rng.id0, L.rng.id1, L.rng.id2,
                                   L.rng.ctr, L.n_step);
                    has = false;
                    ++n_done; /* done as far as this generation is concerned */
                }
            }
            if (kRecordBatch > 0 && __popc(__ballot_sync(0xffffffffu, pending != 0)) >= kRecordBatch)
                flush_pending();
        }
        /* ---- publish finished counts; is the whole block out of work? ---- */
        if (kRecordBatch > 0 && !__ballot_sync(0xffffffffu, has))
            flush_pending(); /* nothing else to do in this warp */
        {
            int s = n_done;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0 && s)
                atomicAdd(A.pool.finished, (unsigned long long)s);
        }
        const int block_live = __syncthreads_or(has ? 1 : 0);
        if (!block_live) {
            if (threadIdx.x == 0) {
                const unsigned long long fin = ld_volatile_u64(A.pool.finished);
                __threadfence();
                const unsigned long long na = ld_volatile_u64(A.pool.n_alloc);
                s_quit = (fin >= na) || (ld_volatile_u32(A.A.error) & 2u);
            }
            __syncthreads();
            if (s_quit)
                break;
            was_idle = true;
            if (++idle_spins > 2)
                __nanosleep(1000);
        } else {
            was_idle = false;
            idle_spins = 0;
        }
    }
    /* flush work counters: warp-reduce, one atomic per warp and counter */
    unsigned int c[7] = {wk.tracked, wk.steps, wk.attempts, wk.interactions, wk.scatters, wk.live_iters, wk.slot_iters};
#pragma unroll
    for (int q = 0; q < 7; ++q) {
        unsigned long long v = c[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v)
            atomicAdd(A.A.work + q, v);
    }
}

/* Hot Compton cross-section table on the device (reference hotcross.cpp:60-79 with the numeric
 * integral :108-142): one thread per (w, theta_e) cell runs the reference's 40 x 240-point midpoint sum in the
 * reference's order, so the entries agree with the CPU table to rounding.  17 901 cells x 9 600 points. */
__global__ void hotcross_table_kernel(double *table, const double *w_axis, const double *theta_axis) {
    /* the axes 10^(l_min + i d_l) are evaluated on the host with std::pow, like the reference: the first theta_e
     * is exactly the table edge 1e-4, where a last-bit difference would select the analytic Klein-Nishina branch */
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= (kHcNW + 1) * (kHcNT + 1))
        return;
    const int i = cell / (kHcNT + 1), j = cell % (kHcNT + 1);
    table[cell] = log10(hotcross_num(w_axis[i], theta_axis[j]));
}

/* FP64 FMA peak probe: 8 independent dependent-FMA chains per thread */
__global__ void fp64_peak_kernel(double *out, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 0.999999, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c);
        a1 = fma(a1, m, c);
        a2 = fma(a2, m, c);
        a3 = fma(a3, m, c);
        a4 = fma(a4, m, c);
        a5 = fma(a5, m, c);
        a6 = fma(a6, m, c);
        a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

} /* namespace gm */
