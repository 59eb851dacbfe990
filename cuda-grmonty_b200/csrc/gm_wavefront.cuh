/*
 * gm_wavefront.cuh -- the state-compacting wavefront kernel (round 2): the transport loop of
 * track_super_photon (reference harm_model.cpp:894-1069) as block-synchronous PHASES over photons that live in
 * shared memory.
 *
 * Why (ncu on the round-1 kernel, profiles/r2_why_more_warps_are_slower.txt): the fused per-lane loop is ~2000
 * straight-line instructions (32 KB) per iteration, every warp streams all of it every iteration and the warps of an
 * SM drift apart, so the 32 KB instruction cache serves 8 - 12 different streams: `no_instruction` stalls were 0.87
 * per issued instruction with 8 warps/SM and 1.74 with 12 -- 50 % more warps issued not one instruction more.  And
 * the interaction half of the loop ran for the 62 % of lanes whose step had just been accepted inside the fluid.
 *
 * How:
 *   - A thread owns R photon slots in SHARED memory (field-major, [field][slot][thread]: a lane only ever touches
 *     its own column, so every access is conflict-free).  A slot is EMPTY, waiting for a PUSH attempt, waiting for
 *     its INTERACTion, or FINished.  Registers hold a photon only inside a phase.
 *   - Every iteration the block counts, with one ballot per warp and one shared-memory atomic per warp, how many
 *     lanes hold a slot in each state, and ALL its warps then run the same phase:
 *         push      one push_photon attempt (step start bookkeeping, halving state machine, step-end stop test)
 *         interact  the fluid interaction of an accepted in-fluid step (opacities, bias, scatter decision)
 *         service   finished photons are recorded / suspended and empty slots refilled from the ready queue
 *         scatter   the scattering stage for photons parked on the global scatter queue (as in round 1)
 *     A lane picks, among its own slots, one that needs the phase: with R >= 2 nearly every lane has one, so the
 *     interaction runs when ~3/4 of the lanes have a step pending instead of after every attempt (compaction by
 *     state without moving a byte), and the warps of an SM execute the same ~15 KB of code at the same time (one
 *     instruction stream per SM instead of one per warp).
 *   - The pre-step snapshot (needed only by the 0.4 % of steps that scatter) goes to a coalesced global scratch
 *     row instead of shared memory: 13 fire-and-forget stores per step.
 *   - Per-photon results are bit-identical to the round-1 kernel: same device functions, same Philox streams, same
 *     frozen statistics and attempt budget; only the order in which photons are advanced differs.
 */
#pragma once
#include "gm_kernels.cuh"

namespace gm {

/* per-slot doubles in shared memory */
enum WfField {
    WF_X0 = 0, WF_X1, WF_X2, WF_X3, WF_K0, WF_K1, WF_K2, WF_K3, WF_DK0, WF_DK1, WF_DK2, WF_DK3,
    WF_E0S, WF_DL, WF_W, WF_TAU_ABS, WF_TAU_SCATT, WF_ALPHA_SCATT, WF_ALPHA_ABS, WF_BI,
    WF_QR, WF_QSTH, WF_QCTH, WF_QHFAC, /* geometry at the new position: handed from the push to the interaction */
    WF_ND
};
/* per-slot words */
enum WfWord { WI_ID0 = 0, WI_ID1, WI_ID2, WI_CTR, WI_SLOT, WI_NSTEP, WI_CLOCK, WI_POSLEV, WI_STATE, WI_NI };

enum WfState {
    WS_EMPTY = 0,
    WS_PUSH,         /* needs a push attempt (pos == level == 0: at a step start) */
    WS_INTERACT,     /* an accepted step ended inside the fluid: interaction pending */
    WS_FIN_RECORD,   /* escaped through r > r_max: to be recorded */
    WS_FIN_DROP,     /* ended without a record (horizon, roulette, step cap) */
    WS_FIN_ABSORBED, /* absorbed (status bit 2) */
    WS_SUSPEND,      /* attempt budget of the generation used up: to the carry queue */
    WS_PARK          /* scatters in this step: to its pool record and the scatter queue */
};

enum WfPhase { WP_PUSH = 0, WP_INTERACT, WP_SERVICE, WP_SCATTER, WP_IDLE };

template <int BLOCK, int R> constexpr size_t wavefront_smem_bytes() {
    return ((size_t)WF_ND * sizeof(double) + (size_t)WI_NI * sizeof(int)) * R * BLOCK;
}

/* ---- phase: one push attempt for the photon in slot (sd, si) ---------------------------------------------------- */
template <int BLOCK, int R>
__device__ __forceinline__ void wf_push(const TransportArgs &A, double *sd, int *si, double *snap, Work &wk) {
    const GmParams &P = A.P;
    constexpr int SD = R * BLOCK, SI = R * BLOCK; /* field strides */
    double x[4], k[4], dk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        x[i] = sd[(WF_X0 + i) * SD];
        k[i] = sd[(WF_K0 + i) * SD];
        dk[i] = sd[(WF_DK0 + i) * SD];
    }
    double e_0_s = sd[WF_E0S * SD];
    double dl = sd[WF_DL * SD];
    const int poslev = si[WI_POSLEV * SI];
    int pos = poslev & 0xff, level = poslev >> 8;
    int clock = si[WI_CLOCK * SI];
    if (poslev == 0) {
        /* ---- step start (reference :919-930).  Out of attempts for this generation: continue in the next one;
         *      checked BEFORE the stop test so that the test (and its roulette draw) runs exactly once per loop
         *      iteration of the reference, on resumption. ---- */
        if (clock >= A.budget) {
            si[WI_STATE * SI] = WS_SUSPEND;
            return;
        }
        double w = sd[WF_W * SD];
        Rng rng = {(uint32_t)si[WI_ID0 * SI], (uint32_t)si[WI_ID1 * SI], (uint32_t)si[WI_ID2 * SI],
                   (uint32_t)si[WI_CTR * SI]};
        if (stop_criterion_fast(A, x[1], w, rng)) {
            sd[WF_W * SD] = w;
            si[WI_CTR * SI] = (int)rng.ctr;
            si[WI_STATE * SI] = x[1] > P.x1_max ? WS_FIN_RECORD : WS_FIN_DROP;
            return;
        }
        sd[WF_W * SD] = w;
        si[WI_CTR * SI] = (int)rng.ctr;
        /* pre-step snapshot: only a step that scatters reads it back (interact -> park) */
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __stcg(snap + (size_t)(0 + i) * A.snap_stride, x[i]);
            __stcg(snap + (size_t)(4 + i) * A.snap_stride, k[i]);
            __stcg(snap + (size_t)(8 + i) * A.snap_stride, dk[i]);
        }
        __stcg(snap + (size_t)12 * A.snap_stride, e_0_s);
        dl = step_size(P, x, k);
        sd[WF_DL * SD] = dl;
    }
    /* ---- one attempt (reference push_photon :1217-1289, recursion flattened as in round 1) ---- */
    double xn[4], kn[4], dkn[4], e1;
    GeoPoint q;
    /* below the grid's inner edge push_photon is a silent no-op (:1218-1220) */
    const bool noop = x[1] < P.x_start1;
    const double dl_now = dl * fm::from_bits((uint64_t)(1023 - level) << 52); /* dl / 2^level, exact */
    const bool fail = push_attempt(P, x, k, dk, dl_now, e_0_s, xn, kn, dkn, e1, q);
    bool accept = true;
    if (!noop) {
        ++wk.attempts;
        ++clock;
        si[WI_CLOCK * SI] = clock;
        accept = !(fail && level < kMaxHalvings);
        if (accept) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                sd[(WF_X0 + i) * SD] = xn[i];
                sd[(WF_K0 + i) * SD] = kn[i];
                sd[(WF_DK0 + i) * SD] = dkn[i];
            }
            sd[WF_E0S * SD] = e1;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            xn[i] = x[i];
    }
    if (!accept) {
        si[WI_POSLEV * SI] = pos | ((level + 1) << 8);
        return;
    }
    pos += 128 >> level;
    if (pos < 128) {
        si[WI_POSLEV * SI] = pos | (halving_next_level(pos) << 8);
        return;
    }
    /* ---- the step is complete (reference :932-936): stop test, then the gate of the interaction ---- */
    si[WI_POSLEV * SI] = 0;
    ++wk.steps;
    double w = sd[WF_W * SD];
    Rng rng = {(uint32_t)si[WI_ID0 * SI], (uint32_t)si[WI_ID1 * SI], (uint32_t)si[WI_ID2 * SI],
               (uint32_t)si[WI_CTR * SI]};
    if (stop_criterion_fast(A, xn[1], w, rng)) {
        sd[WF_W * SD] = w;
        si[WI_CTR * SI] = (int)rng.ctr;
        si[WI_STATE * SI] = xn[1] > P.x1_max ? WS_FIN_RECORD : WS_FIN_DROP;
        return;
    }
    sd[WF_W * SD] = w;
    si[WI_CTR * SI] = (int)rng.ctr;
    const int ns = si[WI_NSTEP * SI];
    if (sd[WF_ALPHA_ABS * SD] > 0.0 || sd[WF_ALPHA_SCATT * SD] > 0.0 || (ns & kNeposBit)) {
        /* hand the geometry of the new point to the interaction phase */
        sd[WF_QR * SD] = q.r;
        sd[WF_QSTH * SD] = q.sth;
        sd[WF_QCTH * SD] = q.cth;
        sd[WF_QHFAC * SD] = q.hfac;
        si[WI_STATE * SI] = WS_INTERACT;
        return;
    }
    /* vacuum step: end of the loop body (:1054-1063) */
    const int n_step = (ns & (kNeposBit - 1)) + 1;
    si[WI_NSTEP * SI] = n_step | (ns & kNeposBit);
    if (n_step > kMaxNStep)
        si[WI_STATE * SI] = WS_FIN_DROP; /* step cap: not recorded (:1060-1066) */
}

/* ---- phase: interaction of the photon in slot (sd, si) after an accepted in-fluid step ---------------------- */
template <int BLOCK, int R>
__device__ __forceinline__ void wf_interact(const TransportArgs &A, double *sd, int *si, const double *snap, Work &wk) {
    constexpr int SD = R * BLOCK, SI = R * BLOCK;
    Live L;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        L.x[i] = sd[(WF_X0 + i) * SD];
        L.k[i] = sd[(WF_K0 + i) * SD];
    }
    L.w = sd[WF_W * SD];
    L.tau_abs = sd[WF_TAU_ABS * SD];
    L.tau_scatt = sd[WF_TAU_SCATT * SD];
    L.alpha_scatt = sd[WF_ALPHA_SCATT * SD];
    L.alpha_abs = sd[WF_ALPHA_ABS * SD];
    L.bi = sd[WF_BI * SD];
    L.dl = sd[WF_DL * SD];
    L.rng.id0 = (uint32_t)si[WI_ID0 * SI];
    L.rng.id1 = (uint32_t)si[WI_ID1 * SI];
    L.rng.id2 = (uint32_t)si[WI_ID2 * SI];
    L.rng.ctr = (uint32_t)si[WI_CTR * SI];
    L.slot = (unsigned int)si[WI_SLOT * SI];
    L.n_step = si[WI_NSTEP * SI] & (kNeposBit - 1);
    L.clock = si[WI_CLOCK * SI];
    L.status = 0;
    L.ne_pos = false;
    GeoPoint q;
    q.r = sd[WF_QR * SD];
    q.rm = q.r + A.P.r_0;
    q.sth = sd[WF_QSTH * SD];
    q.cth = sd[WF_QCTH * SD];
    q.hfac = sd[WF_QHFAC * SD];
    q.sx = 0.0; /* not read by the metric */
    q.cx = 0.0;
    (void)snap;
    const StepResult res = interact<true>(A, L, q, nullptr, 0, wk);
    sd[WF_W * SD] = L.w;
    sd[WF_TAU_ABS * SD] = L.tau_abs;
    sd[WF_TAU_SCATT * SD] = L.tau_scatt;
    sd[WF_ALPHA_SCATT * SD] = L.alpha_scatt;
    sd[WF_ALPHA_ABS * SD] = L.alpha_abs;
    sd[WF_BI * SD] = L.bi;
    si[WI_CTR * SI] = (int)L.rng.ctr;
    if (res == STEP_SCATTER) {
        si[WI_STATE * SI] = WS_PARK; /* the service phase writes the pool record (wf_park) */
        return;
    }
    if (res == STEP_FINISHED) {
        si[WI_NSTEP * SI] = L.n_step | (L.ne_pos ? kNeposBit : 0);
        si[WI_STATE * SI] = WS_FIN_ABSORBED;
        return;
    }
    const int n_step = L.n_step + 1; /* end of the loop body (:1054-1063) */
    si[WI_NSTEP * SI] = n_step | (L.ne_pos ? kNeposBit : 0);
    si[WI_STATE * SI] = n_step > kMaxNStep ? WS_FIN_DROP : WS_PUSH;
}

/* ---- service: finish the photon in slot (sd, si) (cold; once per photon life) ---------------------------- */
template <int BLOCK, int R>
__device__ __noinline__ void wf_finish(const TransportArgs *Ag, double *sd, int *si, int state) {
    const TransportArgs &A = *Ag;
    constexpr int SD = R * BLOCK, SI = R * BLOCK;
    const unsigned int slot = (unsigned int)si[WI_SLOT * SI];
    const int n_step = si[WI_NSTEP * SI] & (kNeposBit - 1);
    if (state == WS_SUSPEND) {
        suspend_photon(Ag, slot, sd[WF_X0 * SD], sd[WF_X1 * SD], sd[WF_X2 * SD], sd[WF_X3 * SD], sd[WF_K0 * SD],
                       sd[WF_K1 * SD], sd[WF_K2 * SD], sd[WF_K3 * SD], sd[WF_DK0 * SD], sd[WF_DK1 * SD],
                       sd[WF_DK2 * SD], sd[WF_DK3 * SD], sd[WF_W * SD], sd[WF_E0S * SD], sd[WF_TAU_ABS * SD],
                       sd[WF_TAU_SCATT * SD], sd[WF_ALPHA_SCATT * SD], sd[WF_ALPHA_ABS * SD], sd[WF_BI * SD],
                       (si[WI_NSTEP * SI] & kNeposBit) != 0, (uint32_t)si[WI_ID0 * SI], (uint32_t)si[WI_ID1 * SI],
                       (uint32_t)si[WI_ID2 * SI], (uint32_t)si[WI_CTR * SI], n_step);
        return;
    }
    cost_call(Ag, slot, n_step);
    int status = state == WS_FIN_ABSORBED ? 4 : 0;
    if (state == WS_FIN_RECORD) {
        record_call(Ag, slot, sd[WF_X2 * SD], sd[WF_X3 * SD], sd[WF_W * SD], sd[WF_TAU_ABS * SD],
                    sd[WF_TAU_SCATT * SD]);
        status |= 1;
    }
    if (A.D.final_state && slot < A.D.n) {
        double *o = A.D.final_state + (size_t)slot * 12;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            o[i] = sd[(WF_X0 + i) * SD];
            o[4 + i] = sd[(WF_K0 + i) * SD];
        }
        o[8] = sd[WF_W * SD];
        o[9] = sd[WF_TAU_ABS * SD];
        o[10] = sd[WF_TAU_SCATT * SD];
        o[11] = sd[WF_E0S * SD];
        atomicOr(A.D.status + slot, status);
        A.pool.rng[slot] = make_uint4((uint32_t)si[WI_ID0 * SI], (uint32_t)si[WI_ID1 * SI], (uint32_t)si[WI_ID2 * SI],
                                      (uint32_t)si[WI_CTR * SI]);
    }
}

/* ---- service: park a photon that scatters in the step it has just made (reference :985-1005; the record layout is
 *      the one scatter_stage reads, see interact<false>) ---------------------------------------------------------- */
template <int BLOCK, int R>
__device__ __noinline__ void wf_park(const TransportArgs *Ag, const double *sd, const int *si, const double *snap) {
    const TransportArgs &A = *Ag;
    constexpr int SD = R * BLOCK, SI = R * BLOCK;
    const PhotonPool &pool = A.pool;
    const unsigned int s = (unsigned int)si[WI_SLOT * SI];
    Rng rng = {(uint32_t)si[WI_ID0 * SI], (uint32_t)si[WI_ID1 * SI], (uint32_t)si[WI_ID2 * SI],
               (uint32_t)si[WI_CTR * SI]};
    const Rng crng = rng_child(A.P, rng);
    /* the scattering stage restarts from the pre-step snapshot and pushes it by dl * frac */
#pragma unroll
    for (int i = 0; i < 12; ++i)
        pstore(pool, P_X0 + i, s, __ldcg(snap + (size_t)i * A.snap_stride));
    pstore(pool, P_E0S, s, __ldcg(snap + (size_t)12 * A.snap_stride));
    pstore(pool, P_W, s, sd[WF_W * SD]);
    pstore(pool, P_TAU_ABS, s, sd[WF_TAU_ABS * SD]);
    pstore(pool, P_TAU_SCATT, s, sd[WF_TAU_SCATT * SD]);
    pstore(pool, P_ALPHA_SCATT, s, sd[WF_ALPHA_SCATT * SD]); /* dl * frac */
    pstore(pool, P_ALPHA_ABS, s, sd[WF_ALPHA_ABS * SD]);     /* weight of the child */
    __stcg(pool.rng + s, make_uint4(rng.id0, rng.id1, rng.id2, rng.ctr));
    __stcg(pool.crng + s, make_uint4(crng.id0, crng.id1, crng.id2, crng.ctr));
    __stcg(pool.n_step + s, si[WI_NSTEP * SI] & (kNeposBit - 1));
    __stcg(pool.gclock + s, si[WI_CLOCK * SI]);
    queue_push(A, A.scatter, s);
}

/* ---- service: take the photon in pool record `slot` into the shared-memory slot (sd, si); false: invalid ------ */
template <int BLOCK, int R>
__device__ __forceinline__ bool wf_load(const TransportArgs &A, unsigned int slot, double *sd, int *si) {
    constexpr int SD = R * BLOCK, SI = R * BLOCK;
    const PhotonPool &pool = A.pool;
    bool bad = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double x = pload(pool, P_X0 + i, slot), k = pload(pool, P_K0 + i, slot);
        bad = bad || isnan(x) || isnan(k);
        sd[(WF_X0 + i) * SD] = x;
        sd[(WF_K0 + i) * SD] = k;
        sd[(WF_DK0 + i) * SD] = pload(pool, P_DK0 + i, slot);
    }
    const double w = pload(pool, P_W, slot);
    bad = bad || (w == 0.0);
    sd[WF_W * SD] = w;
    sd[WF_E0S * SD] = pload(pool, P_E0S, slot);
    sd[WF_TAU_ABS * SD] = pload(pool, P_TAU_ABS, slot);
    sd[WF_TAU_SCATT * SD] = pload(pool, P_TAU_SCATT, slot);
    sd[WF_ALPHA_SCATT * SD] = pload(pool, P_ALPHA_SCATT, slot);
    sd[WF_ALPHA_ABS * SD] = pload(pool, P_ALPHA_ABS, slot);
    sd[WF_BI * SD] = pload(pool, P_BI, slot);
    sd[WF_DL * SD] = 0.0;
    const uint4 r = __ldcg(pool.rng + slot);
    si[WI_ID0 * SI] = (int)r.x;
    si[WI_ID1 * SI] = (int)r.y;
    si[WI_ID2 * SI] = (int)r.z;
    si[WI_CTR * SI] = (int)r.w;
    si[WI_SLOT * SI] = (int)slot;
    si[WI_NSTEP * SI] = __ldcg(pool.n_step + slot);
    si[WI_CLOCK * SI] = __ldcg(pool.gclock + slot);
    si[WI_POSLEV * SI] = 0;
    return !bad;
}

/* ---- the kernel --------------------------------------------------------------------------------------------- */
template <int BLOCK, int R, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) wavefront_kernel(const TransportArgs A) {
    constexpr int SD = R * BLOCK, SI = R * BLOCK;
    constexpr int NWARP = BLOCK / 32;
    double *sd0 = gm_smem + threadIdx.x;                                              /* [WF_ND][R][BLOCK] */
    int *si0 = reinterpret_cast<int *>(gm_smem + (size_t)WF_ND * R * BLOCK) + threadIdx.x; /* [WI_NI][R][BLOCK] */
    /* block control words.  s_cnt: per-iteration lane counts, three buffers in rotation (the one two iterations ahead
     * is cleared after this iteration's barrier, which every thread passes before it can add to that buffer). */
    __shared__ unsigned int s_cnt[3][4];
    /* queue counters seen by thread 0 (ready tail, ready head, scatter entries available), two buffers: written before
     * the barrier of iteration `it` into [it & 1], read by everybody after that barrier ([it & 1]: block-uniform
     * decisions) and before the next one ([(it - 1) & 1] at that point: per-lane hints only) */
    __shared__ unsigned long long s_q[2][4];
    __shared__ unsigned long long s_time[5]; /* clock cycles of thread 0 per phase kind (diagnostics) */
    __shared__ int s_quit;
    const int lane = threadIdx.x & 31;
    const size_t gslot0 = ((size_t)blockIdx.x * R) * BLOCK + threadIdx.x; /* snapshot column of slot 0 */
    long long ticket = -1; /* position in the ready queue this lane is entitled to */
    Work wk = {0u, 0u, 0u, 0u, 0u, 0u, 0u};
    int n_done = 0;
    unsigned int idle_spins = 0;
    /* thread 0: counters loaded (and scatter-queue entries claimed) at the end of an iteration */
    unsigned long long q_tail = 0, q_head = 0, q_scat = 0, q_scat_base = 0;
    unsigned int n_phase[5] = {0u, 0u, 0u, 0u, 0u};        /* thread 0: phases executed by the block, by kind */
    unsigned int live_push = 0, live_int = 0;              /* hot phases in which this lane had work */
    long long t_phase = clock64();
    int last_hot = WP_INTERACT;

#pragma unroll
    for (int r = 0; r < R; ++r)
        si0[WI_STATE * SI + r * BLOCK] = WS_EMPTY;
    if (threadIdx.x == 0) {
        for (int b = 0; b < 3; ++b)
            for (int c = 0; c < 4; ++c)
                s_cnt[b][c] = 0u;
        for (int b = 0; b < 2; ++b) {
            s_q[b][0] = ld_volatile_u64(A.ready.tail);
            s_q[b][1] = ld_volatile_u64(A.ready.head);
            s_q[b][2] = 0ull;
            s_q[b][3] = 0ull;
        }
        s_quit = 0;
        for (int q = 0; q < 5; ++q)
            s_time[q] = 0ull;
        q_tail = s_q[0][0];
        q_head = s_q[0][1];
    }
    __syncthreads();

    for (unsigned int it = 0, i3 = 0;; ++it, i3 = (i3 == 2u ? 0u : i3 + 1u)) {
        if (threadIdx.x == 0) {
            unsigned long long *qw = s_q[it & 1u];
            qw[0] = q_tail;
            qw[1] = q_head;
            qw[2] = q_scat;
            qw[3] = q_scat_base;
        }
        /* ---- what does this lane hold? ---- */
        int r_push = -1, r_int = -1, r_fin = -1, r_empty = -1, n_active = 0;
#pragma unroll
        for (int r = R - 1; r >= 0; --r) {
            const int st = si0[WI_STATE * SI + r * BLOCK];
            if (st == WS_PUSH)
                r_push = r;
            else if (st == WS_INTERACT)
                r_int = r;
            else if (st == WS_EMPTY)
                r_empty = r;
            else
                r_fin = r;
            n_active += (st == WS_PUSH || st == WS_INTERACT) ? 1 : 0;
        }
        const unsigned long long tail_seen = s_q[(it & 1u) ^ 1u][0], head_seen = s_q[(it & 1u) ^ 1u][1];
        /* refill: an empty slot can be filled if the lane's ticket has been (or is about to be) published, or if it
         * has no ticket and there are unclaimed entries.  A lane that already holds a photon only takes another one
         * while the backlog exceeds one photon for every lane of the grid: in the tail of a generation, where the
         * longest-lived lineage sets the run time, every photon should have a lane of its own. */
        const bool backlog = n_active == 0 ? tail_seen > head_seen
                                           : tail_seen > head_seen + (unsigned long long)gridDim.x * BLOCK;
        const bool want_load = r_empty >= 0 && (ticket >= 0 ? (unsigned long long)ticket < tail_seen : backlog);
        {
            const unsigned int bB = __ballot_sync(0xffffffffu, r_push >= 0), bC = __ballot_sync(0xffffffffu, r_int >= 0);
            const unsigned int bF = __ballot_sync(0xffffffffu, r_fin >= 0), bL = __ballot_sync(0xffffffffu, want_load);
            const unsigned int bA = __ballot_sync(0xffffffffu, n_active > 0);
            if (lane == 0) {
                unsigned int *c = s_cnt[i3];
                atomicAdd(c + 0, (unsigned int)__popc(bB) | ((unsigned int)__popc(bC) << 16));
                atomicAdd(c + 1, (unsigned int)__popc(bF) | ((unsigned int)__popc(bL) << 16));
                atomicAdd(c + 2, (unsigned int)__popc(bA));
            }
        }
        __syncthreads();
        const unsigned int c0 = s_cnt[i3][0], c1 = s_cnt[i3][1];
        const int nB = (int)(c0 & 0xffffu), nC = (int)(c0 >> 16), nF = (int)(c1 & 0xffffu), nL = (int)(c1 >> 16);
        const int nA = (int)s_cnt[i3][2];
        const unsigned long long scat_claimed = s_q[it & 1u][2]; /* parked photons claimed for this block */
        if (threadIdx.x == 0) {
            unsigned int *c = s_cnt[i3 == 0u ? 2u : i3 - 1u]; /* (it + 2) mod 3 */
            c[0] = c[1] = c[2] = 0u;
        }
        /* ---- the block's phase for this iteration (every thread computes the same decision) ---- */
        int phase;
        const bool hot = nB > 0 || nC > 0;
        if (scat_claimed > 0ull)
            phase = WP_SCATTER;
        else if ((nF + nL) * 256 >= A.wf_thr_service * (nA > 8 ? nA : 8))
            phase = WP_SERVICE; /* relative to the lanes at work: a sparse block serves its few photons promptly */
        else if (hot && nA * 4 <= BLOCK)
            /* sparse block (the latency-bound head and tail of a generation): strict alternation, no photon waits
             * for a quorum */
            phase = (nB == 0 || (nC > 0 && last_hot == WP_PUSH)) ? WP_INTERACT : WP_PUSH;
        else if (hot)
            phase = (nB == 0 || nC * 256 >= A.wf_thr_interact * nA) ? WP_INTERACT : WP_PUSH;
        else if (nF + nL > 0)
            phase = WP_SERVICE;
        else
            phase = WP_IDLE;

        ++n_phase[phase];
        if (phase <= WP_INTERACT)
            last_hot = phase;
        if (phase == WP_PUSH) {
            ++wk.slot_iters;
            if (r_push >= 0) {
                ++wk.live_iters;
                ++live_push;
                wf_push<BLOCK, R>(A, sd0 + r_push * BLOCK, si0 + r_push * BLOCK,
                                  A.snap + gslot0 + (size_t)r_push * BLOCK, wk);
            }
        } else if (phase == WP_INTERACT) {
            ++wk.slot_iters;
            if (r_int >= 0) {
                ++wk.live_iters;
                ++live_int;
                wf_interact<BLOCK, R>(A, sd0 + r_int * BLOCK, si0 + r_int * BLOCK,
                                      A.snap + gslot0 + (size_t)r_int * BLOCK, wk);
            }
        } else if (phase == WP_SERVICE) {
            /* finished photons: record / suspend, slot becomes empty */
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int st = si0[WI_STATE * SI + r * BLOCK];
                if (st >= WS_FIN_RECORD) {
                    if (st == WS_PARK) {
                        wf_park<BLOCK, R>(A.self, sd0 + r * BLOCK, si0 + r * BLOCK,
                                          A.snap + gslot0 + (size_t)r * BLOCK);
                    } else {
                        wf_finish<BLOCK, R>(A.self, sd0 + r * BLOCK, si0 + r * BLOCK, st);
                        ++n_done;
                    }
                    si0[WI_STATE * SI + r * BLOCK] = WS_EMPTY;
                    if (r_empty < 0 || r < r_empty)
                        r_empty = r;
                }
            }
            /* refill: a ticket per wanting lane (one atomicAdd per warp, never fails), then loads only */
            {
                const bool can = r_empty >= 0;
                const bool take = can && ticket < 0 && backlog;
                const unsigned int need = __ballot_sync(0xffffffffu, take);
                if (need) {
                    unsigned long long base = 0;
                    if (lane == 0)
                        base = atomicAdd(A.ready.head, (unsigned long long)__popc(need));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (take)
                        ticket = (long long)(base + __popc(need & ((1u << lane) - 1u)));
                }
                if (can && ticket >= 0 && ticket < (long long)A.ready.capacity) {
                    /* acquire: pairs with the producer's fence + store in queue_push */
                    const unsigned int v = ld_acquire_u32(A.ready.entries + ticket);
                    if (v) {
                        ticket = -1;
                        double *sd = sd0 + r_empty * BLOCK;
                        int *si = si0 + r_empty * BLOCK;
                        if (wf_load<BLOCK, R>(A, v - 1u, sd, si)) {
                            si[WI_STATE * SI] = WS_PUSH;
                        } else {
                            ++n_done; /* invalid photon (reference :895-900): dropped */
                            if (A.D.status && v - 1u < A.D.n)
                                atomicOr(A.D.status + (v - 1u), 4);
                        }
                    }
                }
            }
            /* publish the finished count (other blocks' quit test reads it) */
            {
                int s = n_done;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                    s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0 && s)
                    atomicAdd(A.pool.finished, (unsigned long long)s);
                n_done = 0;
            }
        } else if (phase == WP_SCATTER) {
            /* ---- scattering stage for the queue entries thread 0 claimed at the end of the last iteration ---- */
            const int cnt = (int)scat_claimed;
            if ((int)threadIdx.x < cnt) {
                const unsigned long long pos = s_q[it & 1u][3] + threadIdx.x;
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
                    wk.attempts += sr.attempts;
                    wk.scatters += sr.scatters;
                    wk.tracked += sr.children;
                    if (sr.done) {
                        int s = sr.done; /* rare: a parent dropped in the scattering stage */
                        atomicAdd(A.pool.finished, (unsigned long long)s);
                    }
                }
            }
        } else {
            /* ---- nothing to do in this block: is the whole generation finished? ---- */
            if (threadIdx.x == 0) {
                const unsigned long long fin = ld_volatile_u64(A.pool.finished);
                __threadfence();
                const unsigned long long na = ld_volatile_u64(A.pool.n_alloc);
                s_quit = (fin >= na) || (ld_volatile_u32(A.A.error) & 2u);
            }
            __syncthreads();
            if (s_quit)
                break;
            if (++idle_spins > 2)
                __nanosleep(1000);
        }
        if (phase != WP_IDLE)
            idle_spins = 0;
        /* ---- queue counters for the next decision (one thread; the loads overlap the next barrier wait) ---- */
        if (threadIdx.x == 0) {
            const long long t_now = clock64();
            s_time[phase] += (unsigned long long)(t_now - t_phase);
            t_phase = t_now;
            const unsigned long long rt = ld_volatile_u64(A.ready.tail), rh = ld_volatile_u64(A.ready.head);
            const unsigned long long sh = ld_volatile_u64(A.scatter.head);
            unsigned long long st = ld_volatile_u64(A.scatter.tail);
            st = st < A.scatter.capacity ? st : A.scatter.capacity;
            q_tail = rt < A.ready.capacity ? rt : A.ready.capacity;
            q_head = rh;
            /* claim parked photons for the next iteration: a full block-load, or -- when this block has nothing
             * else to do -- any */
            const unsigned long long avail = st > sh ? st - sh : 0ull;
            q_scat = 0ull;
            if (avail >= (unsigned long long)BLOCK || (!hot && nF + nL == 0 && avail > 0ull)) {
                const unsigned long long cnt = avail < (unsigned long long)BLOCK ? avail : (unsigned long long)BLOCK;
                if (atomicCAS(A.scatter.head, sh, sh + cnt) == sh) {
                    q_scat = cnt;
                    q_scat_base = sh;
                }
            }
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
    /* phase statistics (A.A.work[8..15]): block-phases executed by kind, lane-phases with work for push / interact */
    if (threadIdx.x == 0)
        for (int q = 0; q < 5; ++q) {
            if (n_phase[q])
                atomicAdd(A.A.work + 8 + q, (unsigned long long)n_phase[q]);
            if (s_time[q])
                atomicAdd(A.A.work + 16 + q, s_time[q]);
        }
    {
        unsigned long long v0 = live_push, v1 = live_int;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v0 += __shfl_xor_sync(0xffffffffu, v0, o);
            v1 += __shfl_xor_sync(0xffffffffu, v1, o);
        }
        if (lane == 0) {
            atomicAdd(A.A.work + 13, v0);
            atomicAdd(A.A.work + 14, v1);
        }
    }
    (void)NWARP;
}

} /* namespace gm */
