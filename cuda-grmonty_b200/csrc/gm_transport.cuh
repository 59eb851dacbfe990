/*
 * gm_transport.cuh -- the superphoton life cycle on the device: photon pool, stage queues, the flattened
 * geodesic/interaction step, the warp-coherent scattering stage, and spectrum recording.
 *
 * Reference (CPU path, cuda_grmonty/harm_model.cpp): track_super_photon :894-1069, stop_criterion :1589-1616,
 * scatter_super_photon :1071-1145, record_super_photon :1291-1335.
 *
 * Design (B200-first):
 *   - ONE persistent kernel per generation of primaries.  A thread owns one live photon whose hot state
 *     (x, k, dk/dlambda, weight, optical depths, previous-point opacities, RNG counter) stays in registers
 *     while it is being stepped; nothing is written back per step.
 *   - Photons live in a POOL of records in HBM (field-major SoA).  Two monotone multi-producer /
 *     multi-consumer index queues connect the stages:
 *         ready queue    photons that can be stepped (primaries from the birth kernel, scattered children,
 *                        parents coming back from a scattering), with ALL derived start-of-track quantities
 *                        (dk/dlambda, opacities, bias) already computed by the producer, so that taking a
 *                        photon is a handful of coalesced loads and never a divergent computation;
 *         scatter queue  photons that decided to scatter in this step.
 *   - Stage compaction: the rare, expensive scattering stage (back-up push, tetrad, electron and
 *     Klein-Nishina rejection sampling, child creation: ~15x a step) is NOT executed inline by the one
 *     lane that needs it.  The lane parks the photon in its pool record and takes new work; any warp that
 *     sees >= 32 parked photons processes them with all 32 lanes active, keeping its own live photons in
 *     registers meanwhile.  (ncu on the inline version: 10.8 of 32 threads active per instruction.)
 *   - push_photon's recursive halving is flattened: one loop iteration = one push attempt for every lane.
 *   - Record stage compaction (round 2): a photon's end is not handled inline by its one lane either.  The
 *     lane leaves what the record needs in shared memory and takes its next photon; the warp makes the
 *     records of 12 lanes together (gm_kernels.cuh, kRecordBatch).  Recorded photons are off the critical
 *     path of their lineage, so the deferral costs nothing; parked (scattering) photons are on it, and the
 *     same deferral for their publication was measured to lose.
 *   - Spectrum bins / counters use global atomics (RED.F64) aggregated per warp by bin.
 *   - Scattering-bias statistics are frozen per generation (GmBiasStats): results do not depend on the
 *     order in which the hardware happens to finish photons.
 *   - The device functions here are shared by the three kernels (fused loop, wavefront, pipelined
 *     scheduler) through template switches (DEFER_PARK, PIPE): per-photon results are identical by
 *     construction, and tests/test_gpu_invariance.py checks it.
 */
#pragma once
#include "gm_geometry.cuh"
#include "gm_fluid.cuh"
#include "gm_params.h"
#include "gm_radiation.cuh"
#include "gm_rng.cuh"
#include "gm_scatter.cuh"

namespace gm {

/* ---- photon pool ------------------------------------------------------------------------------------------ */
enum PField {
    /* hot state */
    P_X0 = 0, P_X1, P_X2, P_X3, P_K0, P_K1, P_K2, P_K3, P_DK0, P_DK1, P_DK2, P_DK3,
    P_W, P_E0S, P_TAU_ABS, P_TAU_SCATT,
    P_ALPHA_SCATT, /* while parked for scattering: dl * frac */
    P_ALPHA_ABS,   /* while parked for scattering: weight of the child */
    P_BI,
    /* cold data: written at birth, read at record / scatter time only */
    P_E, P_X1I, P_X2I, P_NE0, P_TE0, P_B0, P_E0,
    P_NFIELDS
};

struct PhotonPool {
    double *f;       /* [P_NFIELDS][capacity] */
    uint4 *rng;      /* [capacity] id0 id1 id2 ctr */
    uint4 *crng;     /* [capacity] identity of the child to be created (while parked for scattering) */
    int *n_scatt;    /* [capacity] */
    int *n_step;     /* [capacity] bit 30: fluid n_e > 0 at the previous evaluation */
    int *gclock;     /* [capacity] push attempts made by the photon's lineage in the current generation */
    unsigned long long *n_alloc;  /* bump allocator */
    unsigned long long *finished; /* photons whose life is over */
    unsigned int capacity;
};

/* monotone MPMC queue of pool slots; an entry holds slot + 1 once the record is completely written */
struct SlotQueue {
    unsigned int *entries;
    unsigned long long *head, *tail;
    unsigned int capacity;
};

/* optional per-slot debug output for the test exports */
struct DebugOut {
    double *final_state; /* [n][12]: x[4] k[4] w tau_abs tau_scatt e_0_s, or nullptr */
    int *status;         /* [n]: bit0 recorded, bit1 scattered, bit2 absorbed/dropped */
    unsigned int n;
};

struct Accumulators {
    double *spectrum;                 /* [6][200][13] */
    unsigned long long *counters;     /* [0] created [1] scattered [2] recorded */
    unsigned long long *max_tau_bits; /* max_tau_scatt as the bit pattern of a non-negative double */
    unsigned long long *work;         /* [0] tracked [1] steps [2] attempts [3] interactions [4] scatter events */
    unsigned int *error;              /* bit0 pool/queue overflow, bit1 queue entry timeout */
};

/* ---- generation pipeline (gm_pipeline.cuh): device-side generation clock ---------------------------------- */
/* the hot control words of the pipelined kernel, one 128-byte line per queue (the fetch-adds and polls of different
 * queues then go to different L2 slices instead of serialising on one line) */
enum CtlWord {
    CW_R0_HEAD = 0, CW_R0_TAIL = 1,                   /* runnable queues by generation parity: lines 0 and 1 */
    CW_R1_HEAD = 16, CW_R1_TAIL = 17,
    CW_L0_HEAD = 32, CW_L0_TAIL = 33, CW_L0_LIM = 34, /* limbo queues (suspended lineages, next generation): 2 and 3 */
    CW_L1_HEAD = 48, CW_L1_TAIL = 49, CW_L1_LIM = 50,
    CW_PRIM_CUR = 64, CW_PRIM_LIM = 65, /* next primary record to issue / end of the open generations' records */
    CW_COMPLETE = 80,                   /* number of complete generations = index of the oldest incomplete one */
    CW_SC_HEAD = 96, CW_SC_TAIL = 97,   /* scatter queue */
    CW_WORDS = 112
};
constexpr int kCwRStride = 16, kCwLStride = 16; /* word distance between the two parities of a queue kind */
constexpr unsigned long long kGateLive = ~0ull; /* CW_L*_LIM: every published entry may be taken */
constexpr int kClockForever = -(1 << 30);       /* lineage clock of the final drain generation: no budget */
constexpr int kTagShift = 28;                   /* bits 28-29 of a record's n_step word: generation & 3 */
constexpr int kFreshBit = 1 << 27;              /* n_step word: P_BI holds theta_e, bias to be taken at pick-up */
constexpr int kNStepMask = (1 << 24) - 1;

struct GenDesc {
    unsigned long long count;    /* primaries of this rank in the generation */
    unsigned long long prim_end; /* end (exclusive) of the generation's records in the current launch's pool */
    int carry_clock0;            /* lineage clock a photon suspended into this generation starts with */
    int lag;                     /* 1: the generation may start one generation early (it opens when g - 2 is complete
                                  * and uses the statistics up to there); 0: it opens when g - 1 is complete */
    unsigned long long t_open, t_done; /* %globaltimer when the generation was opened / completed (diagnostics) */
};

struct GenCtl {
    unsigned long long line[CW_WORDS];
    unsigned long long alloc[4], done[4]; /* ring by generation & 3: lineage records entered / left the generation */
    unsigned long long acc_scatt[4], acc_rec[4], acc_maxtau[4]; /* statistics recorded in the generation */
    double bias_den[4];                   /* frozen bias denominator of the generation (see GmBiasStats) */
    unsigned int lock;
    int g_end;  /* the launch ends when line[CW_COMPLETE] reaches this */
    int n_desc; /* generations of the run (incl. the final drain generation) */
    int pad;
    double bias_norm;
    const GenDesc *desc;
    unsigned long long t_start, t_limit; /* watchdog: %globaltimer at launch, nanoseconds allowed */
};

struct TransportArgs {
    GmParams P;
    GmBiasStats bias;
    PhotonPool pool;
    SlotQueue ready, scatter;
    SlotQueue carry; /* photons suspended at their attempt budget: they continue in the next generation */
    int budget;      /* attempts a lineage may make per generation ("generation clock") */
    Accumulators A;
    DebugOut D;
    /* [2][n0]: sum of steps and number of finished PRIMARIES per radial bin of their birth zone: what the host
     * sorts the next generations' issue order by (long-lived zones first; see gm_api.cu run_batch) */
    unsigned long long *zone_cost;
    /* wavefront kernel (gm_wavefront.cuh): pre-step snapshots [13][snap_stride] of its resident photons, and its
     * phase thresholds in 1/256: the interaction phase runs when that share of the lanes holding work have a step
     * pending, the service phase when that share of all lanes has a finished photon or an empty slot to refill */
    double *snap;
    unsigned int snap_stride;
    int wf_thr_interact, wf_thr_service;
    /* pipelined kernel: control block, and the four queues R0 R1 L0 L1 (qcap entries each, contiguous) */
    GenCtl *ctl;
    unsigned int *qent;
    unsigned int qcap;
    /* device-global copy of this very struct: out-of-line (cold) stages take it by pointer so that the kernel
     * parameter itself never has its address taken and stays in the constant bank for the hot loop */
    const TransportArgs *self;
};

constexpr int kNeposBit = 1 << 30;

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}
__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int *p) {
    return *reinterpret_cast<const volatile unsigned int *>(p);
}
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
/* Bounds checks of our own (test library, -DGRMONTY_B200_TEST_EXPORTS; compute-sanitizer is not available on the GPU
 * pool): every pool access through pload / pstore and every slot number taken from a queue entry or a ticket is
 * checked against the pool capacity; a violation is counted (grmonty_b200_test_bounds_violations) and redirected to
 * record 0 instead of touching memory outside the pool.  tests/test_gpu_invariance.py runs both schedulers and both
 * kernels through the checked build and asserts a count of zero. */
#ifdef GRMONTY_B200_TEST_EXPORTS
__device__ unsigned int g_bounds_violations = 0u;
__device__ __forceinline__ unsigned int chk_slot(const PhotonPool &pool, unsigned int slot) {
    if (slot >= pool.capacity) {
        atomicAdd(&g_bounds_violations, 1u);
        return 0u;
    }
    return slot;
}
#else
__device__ __forceinline__ unsigned int chk_slot(const PhotonPool &, unsigned int slot) { return slot; }
#endif
__device__ __forceinline__ double pload(const PhotonPool &pool, int field, unsigned int slot) {
    return __ldcg(pool.f + (size_t)field * pool.capacity + chk_slot(pool, slot));
}
__device__ __forceinline__ void pstore(const PhotonPool &pool, int field, unsigned int slot, double v) {
    __stcg(pool.f + (size_t)field * pool.capacity + chk_slot(pool, slot), v);
}

/* allocate a pool record; returns false (and flags the overflow) when the pool is full */
__device__ __forceinline__ bool pool_alloc(const TransportArgs &A, unsigned int &slot) {
    const unsigned long long s = atomicAdd(A.pool.n_alloc, 1ull);
    if (s >= A.pool.capacity) {
        atomicOr(A.A.error, 1u);
        atomicAdd(A.pool.finished, 1ull); /* keep finished == n_alloc reachable */
        return false;
    }
    slot = (unsigned int)s;
    return true;
}

/* publish a completely written record on a queue */
__device__ __forceinline__ void queue_push(const TransportArgs &A, const SlotQueue &q, unsigned int slot) {
    __threadfence();
    const unsigned long long pos = atomicAdd(q.tail, 1ull);
    if (pos >= q.capacity) {
        atomicOr(A.A.error, 1u);
        atomicAdd(A.pool.finished, 1ull);
        return;
    }
    *reinterpret_cast<volatile unsigned int *>(q.entries + pos) = slot + 1u;
}

/* Ring form of the same queue (pipelined kernel): head and tail count positions without bound, position p lives in
 * entry p & (capacity - 1) (capacity is a power of two), and the consumer of a position writes 0 back.  The live
 * entries of a queue never exceed the number of pool records, so with capacity >= pool capacity a producer finds its
 * entry free; it waits for the 0 all the same (the consumer of the position one lap earlier may still be reading). */
__device__ __forceinline__ void queue_push_ring(const TransportArgs &A, const SlotQueue &q, unsigned int slot) {
    __threadfence();
    const unsigned long long pos = atomicAdd(q.tail, 1ull);
    unsigned int *e = q.entries + (pos & (unsigned long long)(q.capacity - 1u));
    unsigned int spins = 0;
    while (ld_volatile_u32(e) != 0u) {
        if (++spins > (1u << 24)) {
            atomicOr(A.A.error, 1u);
            return;
        }
    }
    *reinterpret_cast<volatile unsigned int *>(e) = slot + 1u;
}

/* Warp-collective pop: lanes with `want` set receive a slot (returns true) if the queue has one.
 * `min_batch`: take nothing unless at least that many entries are available. */
__device__ __forceinline__ bool queue_pop_warp(const TransportArgs &A, const SlotQueue &q, bool want, int min_batch,
                                               unsigned int &slot) {
    const unsigned int need = __ballot_sync(0xffffffffu, want);
    if (!need)
        return false;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    int n_got = 0;
    if (lane == 0) {
        const unsigned long long h = ld_volatile_u64(q.head);
        unsigned long long t = ld_volatile_u64(q.tail);
        t = t < q.capacity ? t : q.capacity;
        if (t > h && t - h >= (unsigned long long)min_batch) {
            const unsigned long long avail = t - h;
            const int n_want = __popc(need);
            n_got = avail < (unsigned long long)n_want ? (int)avail : n_want;
            if (atomicCAS(q.head, h, h + n_got) == h)
                base = h;
            else
                n_got = 0;
        }
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    n_got = __shfl_sync(0xffffffffu, n_got, 0);
    bool got = false;
    if (want) {
        const int my = __popc(need & ((1u << lane) - 1u));
        if (my < n_got) {
            /* the producer reserved the position before publishing the slot: wait for it */
            unsigned int v, spins = 0;
            while ((v = ld_volatile_u32(q.entries + base + my)) == 0u) {
                if (++spins > (1u << 26)) {
                    atomicOr(A.A.error, 2u);
                    break;
                }
            }
            if (v) {
                slot = v - 1u;
                got = true;
            }
            __threadfence();
        }
    }
    return got;
}

/* ---- small physics helpers ----------------------------------------------------------------------------------- */
/* reference stop_criterion, harm_model.cpp:1589-1616 */
/* cold part of stop_criterion: Russian roulette for light photons (draws from the photon's stream).  Everything is
 * passed and returned BY VALUE: a reference parameter of a non-inlined function would force the caller's live
 * photon state (weight, RNG counter) into local memory for the whole loop. */
struct RouletteResult {
    double w;
    uint32_t ctr;
    int stop;
};
__device__ __noinline__ RouletteResult stop_criterion_roulette(const GmParams *Pg, double x1, double w, uint32_t id0,
                                                               uint32_t id1, uint32_t id2, uint32_t ctr) {
    const GmParams &P = *Pg;
    Rng rng = {id0, id1, id2, ctr};
    RouletteResult r;
    const double u = rng_uniform(P, rng);
    r.ctr = rng.ctr;
    if (x1 > P.x1_max) {
        r.w = (u <= 1.0 / kRoulette) ? w * kRoulette : 0.0;
        r.stop = 1;
    } else if (u <= 1.0 / kRoulette) {
        r.w = w * kRoulette;
        r.stop = 0;
    } else {
        r.w = 0.0;
        r.stop = 1;
    }
    return r;
}

__device__ __forceinline__ bool stop_criterion_fast(const TransportArgs &A, double x1, double &w, Rng &rng) {
    if (x1 < A.P.x1_min)
        return true;
    if (w < kWeightMin) {
        const RouletteResult r = stop_criterion_roulette(&A.self->P, x1, w, rng.id0, rng.id1, rng.id2, rng.ctr);
        w = r.w;
        rng.ctr = r.ctr;
        return r.stop != 0;
    }
    return x1 > A.P.x1_max;
}

__device__ __forceinline__ bool stop_criterion(const GmParams &P, double x1, double &w, Rng &rng) {
    if (x1 < P.x1_min)
        return true;
    if (x1 > P.x1_max) {
        if (w < kWeightMin) {
            if (rng_uniform(P, rng) <= 1.0 / kRoulette)
                w *= kRoulette;
            else
                w = 0.0;
        }
        return true;
    }
    if (w < kWeightMin) {
        if (rng_uniform(P, rng) <= 1.0 / kRoulette) {
            w *= kRoulette;
        } else {
            w = 0.0;
            return true;
        }
    }
    return false;
}

/* exp(-dtau) with the reference's 4th-order series below 1e-3 (harm_model.cpp:998-1002, :1047-1051) */
__device__ __forceinline__ double attenuation(double d_tau, bool use_series) {
    double att = 1.0 - d_tau * (1.0 / 24.0) * (24.0 - d_tau * (12.0 - d_tau * (4.0 - d_tau)));
    if (!use_series) /* optical depth >= 1e-3 over one step: rare */
        att = fm::exp_(-d_tau);
    return att;
}

/* start-of-track quantities at a position where geometry and fluid are known
 * (reference track_super_photon :902-915): opacities and bias for wave-vector k and weight w */
struct TrackInit {
    double alpha_scatt, alpha_abs, bi;
    bool ne_pos;
};

/* `lazy`: the bias is taken when the photon is picked up (pipelined generations: a primary is born before its
 * generation's statistics are frozen); bi then holds theta_e */
__device__ __forceinline__ TrackInit track_init(const GmParams &P, double bias_den, const double k[4],
                                                double w, const Fluid &f, bool lazy = false) {
    TrackInit t;
    t.ne_pos = f.n_e > 0.0;
    if (t.ne_pos) {
        double nu;
        opacities(P, k, f, nu, t.alpha_scatt, t.alpha_abs);
        t.bi = lazy ? f.theta_e : bias_func_den(f.theta_e, w, bias_den);
    } else {
        t.alpha_scatt = 0.0;
        t.alpha_abs = 0.0;
        t.bi = 0.0;
    }
    return t;
}

/* reference record_super_photon, harm_model.cpp:1291-1335.  Lanes of the calling (possibly partial) warp
 * that hit the same spectrum bin are combined before the global atomics. */
/* PIPE: the bias statistics (maximum of tau_scatt, scatterings and records) go to the accumulators of the photon's
 * generation `tag`; the generation clock folds them into the run totals when the generation is complete. */
template <bool PIPE = false>
__device__ __forceinline__ void record_super_photon(const TransportArgs &A, unsigned int slot, double x2, double x3,
                                                    double w, double tau_abs, double tau_scatt, int tag = 0) {
    const GmParams &P = A.P;
    const double e = pload(A.pool, P_E, slot);
    bool ok = !(isnan(w) || isnan(e));
    int bin = -1;
    int n_scatt = 0;
    if (ok) {
        atomicMax(PIPE ? A.ctl->acc_maxtau + tag : A.A.max_tau_bits,
                  (unsigned long long)__double_as_longlong(fmax(tau_scatt, 0.0)));
        const double dx2 = (P.x_stop2 - P.x_start2) / (2.0 * kNThBins);
        int ix2;
        if (x2 < 0.5 * (P.x_start2 + P.x_stop2))
            ix2 = (int)(x2 / dx2);
        else
            ix2 = (int)((P.x_stop2 - x2) / dx2);
        const double l_e = log(e);
        const int i_e = (int)((l_e - P.spec_l_e_0) / kSpecDLE + 2.5) - 2;
        ok = !(ix2 < 0 || ix2 >= kNThBins || i_e < 0 || i_e >= kNEBins);
        if (ok) {
            bin = ix2 * kNEBins + i_e;
            n_scatt = __ldcg(A.pool.n_scatt + slot);
        }
    }
    double v[12];
    if (ok) {
        const double x1i = pload(A.pool, P_X1I, slot), x2i = pload(A.pool, P_X2I, slot);
        v[0] = w;                              /* dn_dle    */
        v[1] = w * e;                          /* de_dle    */
        v[2] = 1.0;                            /* nph       */
        v[3] = (double)n_scatt;                /* nscatt    */
        v[4] = w * x1i;                        /* x1i_av    */
        v[5] = w * (x2i * x2i);                /* x2i_sq    */
        v[6] = w * (x3 * x3);                  /* x3f_sq    */
        v[7] = w * tau_abs;                    /* tau_abs   */
        v[8] = w * tau_scatt;                  /* tau_scatt */
        v[9] = w * pload(A.pool, P_NE0, slot);  /* ne_0      */
        v[10] = w * pload(A.pool, P_TE0, slot); /* theta_e_0 */
        v[11] = w * pload(A.pool, P_B0, slot);  /* b_0       */
    }
    /* warp-aggregate by bin among the lanes that are here together */
    const unsigned int active = __activemask();
    /* (pipelined: lanes of different generations in one bin are not combined -- the counters are per generation) */
    const unsigned int peers = __match_any_sync(active, PIPE && bin >= 0 ? bin * 4 + tag : bin);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    unsigned long long cnt_scatt = (unsigned long long)n_scatt;
    if (bin >= 0 && peers != (1u << lane)) {
        unsigned int rest = peers & ~(1u << leader);
        while (rest) {
            const int src = __ffs(rest) - 1;
            rest &= rest - 1;
#pragma unroll
            for (int q = 0; q < 12; ++q) {
                const double o = __shfl_sync(peers, v[q], src);
                if (lane == leader)
                    v[q] += o;
            }
            const unsigned long long oc = __shfl_sync(peers, cnt_scatt, src);
            if (lane == leader)
                cnt_scatt += oc;
        }
    }
    if (bin >= 0 && lane == leader) {
        double *s = A.A.spectrum + (size_t)bin * kSpecFields;
#pragma unroll
        for (int q = 0; q < 12; ++q)
            atomicAdd(s + q, v[q]);
        atomicAdd(PIPE ? A.ctl->acc_rec + tag : A.A.counters + 2, (unsigned long long)__popc(peers));
        if (cnt_scatt)
            atomicAdd(PIPE ? A.ctl->acc_scatt + tag : A.A.counters + 1, cnt_scatt);
    }
}

/* per-lane live photon (registers) */
struct Live {
    double x[4], k[4], dk[4];
    double w, e_0_s, tau_abs, tau_scatt;
    double alpha_scatt, alpha_abs, bi; /* values at the previous point (alpha_scatti, alpha_absi, bi) */
    double dl;                         /* size of the step in progress */
    Rng rng;
    unsigned int slot;
    int n_step;
    int clock; /* generation clock, see PhotonPool::gclock */
    int pos, level; /* halving state of the step in progress; pos == 0 && level == 0: at a step start */
    bool ne_pos;    /* fluid n_e > 0 at the previous evaluation */
    int status;
};

/* take a photon from its pool record: loads only (the producer computed every derived quantity); returns the
 * record's n_step word (the pipelined kernel keeps its generation tag there) */
__device__ __forceinline__ int live_load(const TransportArgs &A, unsigned int slot, Live &L) {
    const PhotonPool &pool = A.pool;
    slot = chk_slot(pool, slot);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        L.x[i] = pload(pool, P_X0 + i, slot);
        L.k[i] = pload(pool, P_K0 + i, slot);
        L.dk[i] = pload(pool, P_DK0 + i, slot);
    }
    L.w = pload(pool, P_W, slot);
    L.e_0_s = pload(pool, P_E0S, slot);
    L.tau_abs = pload(pool, P_TAU_ABS, slot);
    L.tau_scatt = pload(pool, P_TAU_SCATT, slot);
    L.alpha_scatt = pload(pool, P_ALPHA_SCATT, slot);
    L.alpha_abs = pload(pool, P_ALPHA_ABS, slot);
    L.bi = pload(pool, P_BI, slot);
    const uint4 r = __ldcg(pool.rng + slot);
    L.rng.id0 = r.x;
    L.rng.id1 = r.y;
    L.rng.id2 = r.z;
    L.rng.ctr = r.w;
    const int ns = __ldcg(pool.n_step + slot);
    L.ne_pos = (ns & kNeposBit) != 0;
    L.n_step = ns & (kNeposBit - 1);
    L.clock = __ldcg(pool.gclock + slot);
    L.slot = slot;
    L.pos = 0;
    L.level = 0;
    L.dl = 0.0;
    L.status = 0;
    return ns;
}

/* work counters kept per thread and flushed once */
struct Work {
    unsigned int tracked, steps, attempts, interactions, scatters;
    unsigned int live_iters, slot_iters; /* loop iterations with / regardless of a live photon in the lane */
};

enum StepResult { STEP_CONTINUE = 0, STEP_FINISHED = 1, STEP_SCATTER = 2, STEP_SUSPEND = 3 };

/* Interaction with the fluid after an accepted step (reference harm_model.cpp:936-1056).
 * If the photon scatters in this step it is parked for the scattering stage (STEP_SCATTER).
 * Single exit, no early returns: the lanes of a warp must leave this function together (see advance). */
/* DEFER_PARK (wavefront kernel): a photon that scatters is not written to its pool record here -- the lane returns
 * STEP_SCATTER with the parked record's values in L (w attenuated, tau_abs / tau_scatt at the scattering point,
 * alpha_scatt = dl * frac, alpha_abs = weight of the child, rng before the child-identity draw) and the record is
 * written by the service phase (wf_park), where all lanes that have something rare to do are together. */
/* PIPE (pipelined generations): the bias denominator is the one of the photon's generation (sbias[tag], a copy of
 * GenCtl::bias_den in shared memory; tag = bits 4-5 of L.status), and a parked record keeps the tag. */
template <bool DEFER_PARK = false, bool PIPE = false>
__device__ __forceinline__ StepResult interact(const TransportArgs &A, Live &L, const GeoPoint &q,
                                               const double *snap, int snap_stride, Work &wk,
                                               const double *sbias = nullptr) {
    const GmParams &P = A.P;
    ++wk.interactions;
    /* q = geometry at the new position, already evaluated by the accepted push attempt */
    const MetricCov g = metric_cov(P, q);
    Fluid f;
    fluid_params(P, L.x[1], L.x[2], g, q, f);
    const bool bound = (f.n_e == 0.0);
    double e_fluid = 0.0, mu = 0.0;
    if (!bound)
        fluid_frame(P, L.k, f, e_fluid, mu);
    const double nu = bound ? 0.0 : e_fluid * (kME * kCL * kCL / kHPL);
    const bool outside = bound || nu < 0.0;
    /* evaluate the opacities for every lane (harmless dummy arguments outside the fluid) so that the warp
     * stays converged through the expensive part; the three logarithms (nu, theta_e, and the interaction
     * draw below) are independent and interleave */
    const double nu_e = (outside || !(nu > 0.0)) ? 1.0e12 : nu;
    const double te_e = outside ? 1.0 : f.theta_e, ne_e = outside ? 1.0 : f.n_e;
    const double b_e = outside ? 1.0 : f.b;
    const double l_nu = fm::log_(nu_e), l_theta = fm::log_(te_e);
    const double a_sf = alpha_inv_scatt_l(P, nu_e, te_e, ne_e, l_nu, l_theta);
    const double a_af = alpha_inv_abs_sin_l(P, nu_e, te_e, ne_e, b_e, fm::sqrt_(1.0 - mu * mu), l_theta);
    const double bf = bias_func_den(te_e, L.w, PIPE ? sbias[(L.status >> 4) & 3] : A.bias.bias_den);
    double d_tau_scatt, d_tau_abs, bias;
    if (outside) {
        d_tau_scatt = 0.5 * L.alpha_scatt * P.d_tau_k * L.dl;
        d_tau_abs = 0.5 * L.alpha_abs * P.d_tau_k * L.dl;
        L.alpha_scatt = 0.0;
        L.alpha_abs = 0.0;
        bias = 0.0;
        L.bi = 0.0;
    } else {
        d_tau_scatt = 0.5 * (L.alpha_scatt + a_sf) * P.d_tau_k * L.dl;
        L.alpha_scatt = a_sf;
        d_tau_abs = 0.5 * (L.alpha_abs + a_af) * P.d_tau_k * L.dl;
        L.alpha_abs = a_af;
        bias = 0.5 * (L.bi + bf);
        L.bi = bf;
    }
    L.ne_pos = f.n_e > 0.0;
    /* Scatter test of the reference (:980-985): x1 = -ln U, scatter iff bias * d_tau_scatt > x1 and w / bias > w_min.
     * -ln U >= 1 - U, so when 1 - U exceeds bias * d_tau_scatt (by a margin far above the rounding of the logarithm)
     * the test fails whatever x1 is: the logarithm and the division run only for the ~0.5 % of steps that can
     * scatter, and every decision is the one the reference's form gives. */
    const double u_scatt = rng_uniform(P, L.rng);
    const double bd = bias * d_tau_scatt;
    double x1r = 0.0, w_child = 0.0;
    bool scatters = false;
    if ((1.0 - u_scatt) <= bd * 1.0000001) {
        x1r = -fm::log_(u_scatt);
        /* w / bias; bias == 0 gives +inf in the reference (Appendix A.3), which the test rejects through bd = 0 */
        w_child = bias > 0.0 ? fm::div(L.w, bias) : fm::from_bits(0x7ff0000000000000ull);
        scatters = bd > x1r && w_child > kWeightMin;
    }
    StepResult res = STEP_CONTINUE;
    if (DEFER_PARK && scatters) {
        const double frac = fm::div(x1r, bd);
        d_tau_abs *= frac;
        if (d_tau_abs > 100) {
            L.rng.ctr += 1u; /* the child-identity draw of the fused form (rng_child), never used */
            L.status |= 4;
            res = STEP_FINISHED; /* absorbed before scattering */
        } else {
            d_tau_scatt *= frac;
            L.w *= attenuation(d_tau_abs + d_tau_scatt, d_tau_abs < 1.0e-3);
            L.tau_abs += d_tau_abs;
            L.tau_scatt += d_tau_scatt;
            L.alpha_scatt = L.dl * frac;
            L.alpha_abs = w_child;
            res = STEP_SCATTER;
        }
    } else if (scatters) {
        /* ---- the photon scatters in this step (reference :985-1005): park it ---- */
        const Rng crng = rng_child(P, L.rng);
        const double frac = fm::div(x1r, bd);
        d_tau_abs *= frac;
        if (d_tau_abs > 100) {
            L.status |= 4;
            res = STEP_FINISHED; /* absorbed before scattering */
        } else {
            d_tau_scatt *= frac;
            L.w *= attenuation(d_tau_abs + d_tau_scatt, d_tau_abs < 1.0e-3);
            const PhotonPool &pool = A.pool;
            const unsigned int s = L.slot;
            /* the scattering stage restarts from the pre-step snapshot and pushes it by dl * frac */
#pragma unroll
            for (int i = 0; i < 12; ++i)
                pstore(pool, P_X0 + i, s, snap[i * snap_stride]);
            pstore(pool, P_E0S, s, snap[12 * snap_stride]);
            pstore(pool, P_W, s, L.w);
            pstore(pool, P_TAU_ABS, s, L.tau_abs + d_tau_abs);
            pstore(pool, P_TAU_SCATT, s, L.tau_scatt + d_tau_scatt);
            pstore(pool, P_ALPHA_SCATT, s, L.dl * frac);
            pstore(pool, P_ALPHA_ABS, s, w_child);
            __stcg(pool.rng + s, make_uint4(L.rng.id0, L.rng.id1, L.rng.id2, L.rng.ctr));
            __stcg(pool.crng + s, make_uint4(crng.id0, crng.id1, crng.id2, crng.ctr));
            __stcg(pool.n_step + s, PIPE ? (L.n_step | (((L.status >> 4) & 3) << kTagShift)) : L.n_step);
            __stcg(pool.gclock + s, L.clock);
            if (PIPE)
                queue_push_ring(A, A.scatter, s);
            else
                queue_push(A, A.scatter, s);
            res = STEP_SCATTER;
        }
    } else if (d_tau_abs > 100) {
        L.status |= 4;
        res = STEP_FINISHED; /* absorbed */
    } else {
        const double d_tau = d_tau_abs + d_tau_scatt;
        L.w *= attenuation(d_tau, d_tau < 1.0e-3);
        L.tau_abs += d_tau_abs;
        L.tau_scatt += d_tau_scatt;
    }
    return res;
}

/* One iteration of the flattened per-photon loop, in three phases that every live lane of the warp walks
 * through together:  A  step-start bookkeeping (suspend / stop test, snapshot, step size);
 *                    B  one push attempt;
 *                    C  step end (stop test, interaction with the fluid).
 * There are no early returns: each phase ends with __syncwarp(live) so that the lanes RECONVERGE before the
 * next one.  (With early returns the compiler let the lanes that came through phase A and those that were in
 * the middle of a halved step run phase B separately: ncu showed 15.6 of ~27 live threads per instruction.)
 * `record` tells whether a finished photon escaped through r > r_max (reference :1066-1068). */
template <bool PIPE = false>
__device__ __forceinline__ StepResult advance(const TransportArgs &A, Live &L, unsigned int live, double *snap,
                                              int snap_stride, Work &wk, bool &record, const double *sbias = nullptr) {
    const GmParams &P = A.P;
    record = false;
    StepResult st = STEP_CONTINUE;
    /* ---- phase A ---- */
    if (L.pos == 0 && L.level == 0) {
        /* out of attempts for this generation: continue in the next one.  Checked BEFORE the stop test so
         * that the test (and its roulette draw) runs exactly once per loop iteration, on resumption. */
        if (L.clock >= A.budget) {
            st = STEP_SUSPEND;
        } else if (stop_criterion_fast(A, L.x[1], L.w, L.rng)) { /* top of the while loop (:919) */
            record = L.x[1] > P.x1_max;
            st = STEP_FINISHED;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                snap[(0 + i) * snap_stride] = L.x[i];
                snap[(4 + i) * snap_stride] = L.k[i];
                snap[(8 + i) * snap_stride] = L.dk[i];
            }
            snap[12 * snap_stride] = L.e_0_s;
            L.dl = step_size(P, L.x, L.k);
        }
    }
    __syncwarp(live);
    /* ---- phase B ---- */
    bool step_done = false;
    GeoPoint q;
    if (st == STEP_CONTINUE) {
        double xn[4], kn[4], dkn[4], e1;
        /* below the grid's inner edge push_photon is a silent no-op (:1218-1220): the attempt is computed but
         * discarded, so that the warp does not diverge (it happens only inside the horizon) */
        const bool noop = L.x[1] < P.x_start1;
        /* dl / 2^level: an exact scaling */
        const double dl_now = L.dl * fm::from_bits((uint64_t)(1023 - L.level) << 52);
        const bool fail = push_attempt(P, L.x, L.k, L.dk, dl_now, L.e_0_s, xn, kn, dkn, e1, q);
        bool accept = true;
        if (!noop) {
            ++wk.attempts;
            ++L.clock;
            accept = !(fail && L.level < kMaxHalvings);
            if (accept) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    L.x[i] = xn[i];
                    L.k[i] = kn[i];
                    L.dk[i] = dkn[i];
                }
                L.e_0_s = e1;
            }
        }
        if (!accept) {
            ++L.level;
        } else {
            L.pos += 128 >> L.level;
            if (L.pos < 128) {
                L.level = halving_next_level(L.pos);
            } else {
                L.pos = 0;
                L.level = 0;
                step_done = true;
                ++wk.steps;
            }
        }
    }
    __syncwarp(live);
    /* ---- phase C ---- */
    if (step_done) {
        if (stop_criterion_fast(A, L.x[1], L.w, L.rng)) {
            record = L.x[1] > P.x1_max;
            st = STEP_FINISHED;
        } else {
            if (L.alpha_abs > 0.0 || L.alpha_scatt > 0.0 || L.ne_pos)
                st = interact<false, PIPE>(A, L, q, snap, snap_stride, wk, sbias);
            if (st == STEP_CONTINUE) {
                ++L.n_step;
                if (L.n_step > kMaxNStep)
                    st = STEP_FINISHED; /* step cap: not recorded (:1060-1066) */
            }
        }
    }
    __syncwarp(live);
    return st;
}

/* write the start-of-track record of a photon (hot part); cold fields are written by the caller */
__device__ __forceinline__ void pool_store_hot(const PhotonPool &pool, unsigned int s, const double x[4],
                                               const double k[4], const double dk[4], double w, double e_0_s,
                                               double tau_abs, double tau_scatt, const TrackInit &t, const Rng &rng,
                                               int n_step) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        pstore(pool, P_X0 + i, s, x[i]);
        pstore(pool, P_K0 + i, s, k[i]);
        pstore(pool, P_DK0 + i, s, dk[i]);
    }
    pstore(pool, P_W, s, w);
    pstore(pool, P_E0S, s, e_0_s);
    pstore(pool, P_TAU_ABS, s, tau_abs);
    pstore(pool, P_TAU_SCATT, s, tau_scatt);
    pstore(pool, P_ALPHA_SCATT, s, t.alpha_scatt);
    pstore(pool, P_ALPHA_ABS, s, t.alpha_abs);
    pstore(pool, P_BI, s, t.bi);
    __stcg(pool.rng + s, make_uint4(rng.id0, rng.id1, rng.id2, rng.ctr));
    __stcg(pool.n_step + s, n_step | (t.ne_pos ? kNeposBit : 0));
}

/* Suspend a live photon at a step boundary: its complete hot state goes back to its pool record and the record
 * is put on the carry queue; the host moves carried records into the next generation's batch. */
__device__ __noinline__ void suspend_photon(const TransportArgs *Ag, unsigned int slot, double x0, double x1,
                                            double x2, double x3, double k0, double k1, double k2, double k3,
                                            double dk0, double dk1, double dk2, double dk3, double w, double e_0_s,
                                            double tau_abs, double tau_scatt, double alpha_scatt, double alpha_abs,
                                            double bi, bool ne_pos, uint32_t id0, uint32_t id1, uint32_t id2,
                                            uint32_t ctr, int n_step) {
    /* all arguments by value (see stop_criterion_roulette) */
    const TransportArgs &A = *Ag;
    TrackInit t;
    t.alpha_scatt = alpha_scatt;
    t.alpha_abs = alpha_abs;
    t.bi = bi;
    t.ne_pos = ne_pos;
    const double x[4] = {x0, x1, x2, x3}, k[4] = {k0, k1, k2, k3}, dk[4] = {dk0, dk1, dk2, dk3};
    const Rng rng = {id0, id1, id2, ctr};
    pool_store_hot(A.pool, slot, x, k, dk, w, e_0_s, tau_abs, tau_scatt, t, rng, n_step);
    __stcg(A.pool.gclock + slot, 0);
    queue_push(A, A.carry, slot);
}

/* The scattering stage for one parked photon (reference harm_model.cpp:1005-1039 + scatter_super_photon).
 * Called with all lanes of a warp holding a parked photon (or idle).  The parent and the child that continue
 * are pushed on the ready queue.
 * PIPE (pipelined generations): the record's n_step word carries the generation tag; statistics are those of that
 * generation, the child is counted in it (GenCtl::alloc) and both go to the runnable queue of its parity. */
struct ScatterStageResult {
    int done;                                  /* photons whose life ended here (0 or 1) */
    unsigned int attempts, scatters, children; /* work counters, returned by value (see stop_criterion_roulette) */
    int tag;                                   /* PIPE: generation tag of the lineage */
};
template <bool PIPE>
__device__ __noinline__ ScatterStageResult scatter_stage_t(const TransportArgs *Ag, unsigned int slot) {
    unsigned int n_attempts = 0, n_scatters = 0, n_children = 0;
    const TransportArgs &A = *Ag;
    const GmParams &P = A.P;
    const PhotonPool &pool = A.pool;
    slot = chk_slot(pool, slot);
    double x[4], k[4], dk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        x[i] = pload(pool, P_X0 + i, slot);
        k[i] = pload(pool, P_K0 + i, slot);
        dk[i] = pload(pool, P_DK0 + i, slot);
    }
    double e_0_s = pload(pool, P_E0S, slot);
    double w = pload(pool, P_W, slot);
    const double tau_abs = pload(pool, P_TAU_ABS, slot), tau_scatt = pload(pool, P_TAU_SCATT, slot);
    const double dl_frac = pload(pool, P_ALPHA_SCATT, slot);
    const double w_child = pload(pool, P_ALPHA_ABS, slot);
    const uint4 r4 = __ldcg(pool.rng + slot);
    Rng rng = {r4.x, r4.y, r4.z, r4.w};
    const int ns_word = __ldcg(pool.n_step + slot);
    const int tag = PIPE ? (ns_word >> kTagShift) & 3 : 0;
    const int tag_bits = PIPE ? tag << kTagShift : 0;
    int n_step = ns_word & (PIPE ? kNStepMask : kNeposBit - 1);
    int clock = __ldcg(pool.gclock + slot);
    const double bias_den = PIPE ? __ldcg(A.ctl->bias_den + tag) : A.bias.bias_den;
    const SlotQueue ready_q = PIPE ? SlotQueue{A.qent + (size_t)(tag & 1) * A.qcap, A.ctl->line + CW_R0_HEAD + kCwRStride * (tag & 1),
                                               A.ctl->line + CW_R0_TAIL + kCwRStride * (tag & 1), A.qcap}
                                   : A.ready;

    /* back up to the scattering point */
    {
        const int att = push_photon_full(P, x, k, dk, e_0_s, dl_frac);
        n_attempts += att;
        clock += att;
    }
    const GeoPoint q = geo_point(P, x[1], x[2]);
    const MetricCov g = metric_cov(P, q);
    Fluid f;
    fluid_params(P, x[1], x[2], g, q, f);
    TrackInit ti;
    if (f.n_e > 0.0) {
        ++n_scatters;
        ScatterChild ch;
        const bool child_ok = scatter_super_photon(P, rng, k, w, f, g, ch);
        if (w < 1.0e-100) {
            if (A.D.status && slot < A.D.n)
                A.D.status[slot] = 4 | 2;
            return ScatterStageResult{1, n_attempts, n_scatters, n_children, tag}; /* k could not be put back on the
                                                                                     * light cone (:1018-1021): dropped */
        }
        if (child_ok) {
            unsigned int cs;
            if (pool_alloc(A, cs)) {
                const uint4 c4 = __ldcg(pool.crng + slot);
                const Rng crng = {c4.x, c4.y, c4.z, c4.w};
                Connection c;
                connection_eval(P, q, c);
                double dkc[4];
                geodesic_rhs(c, ch.k, dkc);
                const TrackInit tc = track_init(P, bias_den, ch.k, w_child, f);
                pool_store_hot(pool, cs, x, ch.k, dkc, w_child, ch.e, 0.0, 0.0, tc, crng, tag_bits);
                pstore(pool, P_E, cs, ch.e);
                pstore(pool, P_X1I, cs, x[1]);
                pstore(pool, P_X2I, cs, x[2]);
                pstore(pool, P_NE0, cs, pload(pool, P_NE0, slot));
                pstore(pool, P_TE0, cs, pload(pool, P_TE0, slot));
                pstore(pool, P_B0, cs, f.b);
                pstore(pool, P_E0, cs, pload(pool, P_E0, slot));
                __stcg(pool.n_scatt + cs, __ldcg(pool.n_scatt + slot) + 1);
                __stcg(pool.gclock + cs, clock); /* the child inherits its lineage's clock */
                ++n_children;
                if (PIPE) {
                    atomicAdd(A.ctl->alloc + tag, 1ull); /* before the record is published */
                    queue_push_ring(A, ready_q, cs);
                } else {
                    queue_push(A, ready_q, cs);
                }
            }
        }
        ti = track_init(P, bias_den, k, w, f);
    } else {
        /* left the grid while backing up (the reference reads uninitialised data here, Appendix A.15) */
        ti.alpha_scatt = 0.0;
        ti.alpha_abs = 0.0;
        ti.bi = 0.0;
        ti.ne_pos = false;
    }
    if (A.D.status && slot < A.D.n)
        atomicOr(A.D.status + slot, 2);
    /* end of the loop body (:1054-1063) */
    ++n_step;
    if (n_step > kMaxNStep) {
        if (A.D.status && slot < A.D.n)
            atomicOr(A.D.status + slot, 4);
        return ScatterStageResult{1, n_attempts, n_scatters, n_children, tag};
    }
    pool_store_hot(pool, slot, x, k, dk, w, e_0_s, tau_abs, tau_scatt, ti, rng, n_step | tag_bits);
    __stcg(pool.gclock + slot, clock);
    if (PIPE)
        queue_push_ring(A, ready_q, slot);
    else
        queue_push(A, ready_q, slot);
    return ScatterStageResult{0, n_attempts, n_scatters, n_children, tag};
}
__device__ __forceinline__ ScatterStageResult scatter_stage(const TransportArgs *Ag, unsigned int slot) {
    return scatter_stage_t<false>(Ag, slot);
}

/* Pipelined generations: suspend a photon of generation `g` (tag g & 3) at its attempt budget.  It continues in
 * generation g + 1: the record is rewritten in place with that generation's tag and starting clock, counted in its
 * alloc counter and put on the limbo queue of its parity, where it waits until the generation is open. */
__device__ __noinline__ void suspend_photon_pipe(const TransportArgs *Ag, unsigned int slot, double x0, double x1,
                                                 double x2, double x3, double k0, double k1, double k2, double k3,
                                                 double dk0, double dk1, double dk2, double dk3, double w, double e_0_s,
                                                 double tau_abs, double tau_scatt, double alpha_scatt, double alpha_abs,
                                                 double bi, bool ne_pos, uint32_t id0, uint32_t id1, uint32_t id2,
                                                 uint32_t ctr, int n_step, int tag) {
    const TransportArgs &A = *Ag;
    GenCtl *C = A.ctl;
    TrackInit t;
    t.alpha_scatt = alpha_scatt;
    t.alpha_abs = alpha_abs;
    t.bi = bi;
    t.ne_pos = ne_pos;
    const double x[4] = {x0, x1, x2, x3}, k[4] = {k0, k1, k2, k3}, dk[4] = {dk0, dk1, dk2, dk3};
    const Rng rng = {id0, id1, id2, ctr};
    /* the photon's generation: the one of [nc, nc + 3) with this tag (nc: oldest incomplete generation) */
    const long long nc = (long long)ld_volatile_u64(C->line + CW_COMPLETE);
    const long long g = nc + ((tag - (int)(nc & 3)) & 3);
    const long long gn = g + 1 < C->n_desc ? g + 1 : C->n_desc - 1;
    const int nt = (tag + 1) & 3;
    pool_store_hot(A.pool, slot, x, k, dk, w, e_0_s, tau_abs, tau_scatt, t, rng, n_step | (nt << kTagShift));
    __stcg(A.pool.gclock + slot, C->desc[gn].carry_clock0);
    atomicAdd(C->alloc + nt, 1ull); /* before this photon is counted as done in its old generation */
    const SlotQueue lq = {A.qent + (size_t)(2 + (nt & 1)) * A.qcap, C->line + CW_L0_HEAD + kCwLStride * (nt & 1),
                          C->line + CW_L0_TAIL + kCwLStride * (nt & 1), A.qcap};
    queue_push_ring(A, lq, slot);
}

} /* namespace gm */
