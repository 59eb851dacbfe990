/*
 * gm_transport.cuh -- the superphoton life cycle on the device: birth, persistent transport loop, record.
 *
 * Reference (CPU path, cuda_grmonty/harm_model.cpp): make_super_photon :794-811, get_zone :673-704,
 * init_zone :1337-1389, sample_zone_photon :706-782, track_super_photon :894-1069, stop_criterion :1589-1616,
 * record_super_photon :1291-1335.
 *
 * Design (B200-first):
 *   - ONE persistent kernel per generation of primaries.  Every thread owns one live photon whose hot state
 *     (x, k, dk/dlambda, weight, optical depths, opacities at the previous point, RNG counter) stays in
 *     registers for the photon's whole flight; nothing is written back per step.
 *   - push_photon's recursive halving is flattened: one loop iteration = one push attempt for every lane, so
 *     lanes that are re-taking a halved sub-step and lanes starting a new step execute the same code.
 *   - photons waiting to be tracked (this generation's primaries, written by the birth kernel, and scattered
 *     children, appended by the transport kernel itself) live in a monotone multi-producer/multi-consumer
 *     queue in HBM (field-major SoA, coalesced warp-aggregated pops and pushes).  Cold per-photon data
 *     (emission-point diagnostics, e, l, n_scatt) is never carried in registers: it stays in the photon's
 *     queue slot and is read back at record / scatter time.
 *   - spectrum bins and counters are accumulated with atomics (RED.F64) aggregated per warp by bin.
 *   - scattering-bias statistics are frozen per generation (GmBiasStats), so results do not depend on the
 *     order in which the hardware happens to finish photons.
 */
#pragma once
#include "gm_geometry.cuh"
#include "gm_fluid.cuh"
#include "gm_params.h"
#include "gm_radiation.cuh"
#include "gm_rng.cuh"
#include "gm_scatter.cuh"

namespace gm {

/* ---- photon queue -------------------------------------------------------------------------------------- */
enum QField {
    Q_X0 = 0, Q_X1, Q_X2, Q_X3, Q_K0, Q_K1, Q_K2, Q_K3, Q_W, Q_E, Q_L, Q_X1I, Q_X2I, Q_NE0, Q_TE0, Q_B0, Q_E0,
    Q_NFIELDS
};

struct PhotonQueue {
    double *f;          /* [Q_NFIELDS][capacity] */
    uint4 *rng;         /* [capacity] id0 id1 id2 ctr */
    int *n_scatt;       /* [capacity] */
    unsigned int *ready; /* [capacity] generation tag once the slot is fully written */
    unsigned long long *head, *tail, *finished;
    unsigned long long capacity;
};

/* optional per-slot debug output for the test exports */
struct DebugOut {
    double *final_state; /* [n][12]: x[4] k[4] w tau_abs tau_scatt e_0_s, or nullptr */
    int *status;         /* [n]: bit0 recorded, bit1 scattered, bit2 absorbed/dropped */
    unsigned long long n;
};

struct Accumulators {
    double *spectrum;               /* [6][200][13] */
    unsigned long long *counters;   /* [0] created [1] scattered [2] recorded */
    unsigned long long *max_tau_bits; /* max_tau_scatt as the bit pattern of a non-negative double */
    unsigned long long *work;       /* [0] tracked [1] steps [2] attempts [3] interactions [4] scatter events */
    unsigned int *error;            /* bit0 queue overflow, bit1 ready-flag timeout */
};

struct TransportArgs {
    GmParams P;
    GmBiasStats bias;
    PhotonQueue Q;
    Accumulators A;
    DebugOut D;
    unsigned int gen_tag;
};

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}
__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int *p) {
    return *reinterpret_cast<const volatile unsigned int *>(p);
}
__device__ __forceinline__ double qload(const PhotonQueue &Q, int field, unsigned int slot) {
    return __ldcg(Q.f + (size_t)field * Q.capacity + slot);
}
__device__ __forceinline__ void qstore(const PhotonQueue &Q, int field, unsigned long long slot, double v) {
    Q.f[(size_t)field * Q.capacity + slot] = v;
}

/* reference stop_criterion, harm_model.cpp:1589-1616 */
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
    if (use_series)
        return 1.0 - d_tau / 24.0 * (24.0 - d_tau * (12.0 - d_tau * (4.0 - d_tau)));
    return exp(-d_tau);
}

/* reference record_super_photon, harm_model.cpp:1291-1335.  Lanes of the calling (possibly partial) warp
 * that hit the same spectrum bin are combined before the global atomics. */
__device__ __forceinline__ void record_super_photon(const TransportArgs &A, unsigned int slot, double x2, double x3,
                                                    double w, double tau_abs, double tau_scatt) {
    const GmParams &P = A.P;
    const double e = qload(A.Q, Q_E, slot);
    bool ok = !(isnan(w) || isnan(e));
    int bin = -1;
    int n_scatt = 0;
    if (ok) {
        atomicMax(A.A.max_tau_bits, (unsigned long long)__double_as_longlong(fmax(tau_scatt, 0.0)));
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
            n_scatt = __ldcg(A.Q.n_scatt + slot);
        }
    }
    double v[12];
    if (ok) {
        const double x1i = qload(A.Q, Q_X1I, slot), x2i = qload(A.Q, Q_X2I, slot);
        v[0] = w;                               /* dn_dle   */
        v[1] = w * e;                           /* de_dle   */
        v[2] = 1.0;                             /* nph      */
        v[3] = (double)n_scatt;                 /* nscatt   */
        v[4] = w * x1i;                         /* x1i_av   */
        v[5] = w * (x2i * x2i);                 /* x2i_sq   */
        v[6] = w * (x3 * x3);                   /* x3f_sq   */
        v[7] = w * tau_abs;                     /* tau_abs  */
        v[8] = w * tau_scatt;                   /* tau_scatt*/
        v[9] = w * qload(A.Q, Q_NE0, slot);     /* ne_0     */
        v[10] = w * qload(A.Q, Q_TE0, slot);    /* theta_e_0*/
        v[11] = w * qload(A.Q, Q_B0, slot);     /* b_0      */
    }
    /* warp-aggregate by bin among the lanes that are here together */
    const unsigned int active = __activemask();
    const unsigned int peers = __match_any_sync(active, bin);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    unsigned long long cnt_scatt = (unsigned long long)n_scatt;
    if (bin >= 0 && peers != (1u << lane)) {
        /* serial reduction over the peer set (peer sets are tiny: records are rare events) */
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
        atomicAdd(A.A.counters + 2, (unsigned long long)__popc(peers));
        if (cnt_scatt)
            atomicAdd(A.A.counters + 1, cnt_scatt);
    }
}

/* append a scattered photon to the queue; returns false on overflow */
__device__ __forceinline__ bool enqueue_child(const TransportArgs &A, const double x[4], const ScatterChild &ch,
                                              double w, double b0, unsigned int parent_slot, const Rng &crng) {
    const PhotonQueue &Q = A.Q;
    const unsigned long long slot = atomicAdd(Q.tail, 1ull);
    if (slot >= Q.capacity) {
        atomicOr(A.A.error, 1u);
        /* the slot index is beyond the arrays: count it as finished so the kernel still terminates */
        atomicAdd(Q.finished, 1ull);
        return false;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        qstore(Q, Q_X0 + i, slot, x[i]);
        qstore(Q, Q_K0 + i, slot, ch.k[i]);
    }
    qstore(Q, Q_W, slot, w);
    qstore(Q, Q_E, slot, ch.e);
    qstore(Q, Q_L, slot, ch.l);
    qstore(Q, Q_X1I, slot, x[1]);
    qstore(Q, Q_X2I, slot, x[2]);
    qstore(Q, Q_NE0, slot, qload(Q, Q_NE0, parent_slot));
    qstore(Q, Q_TE0, slot, qload(Q, Q_TE0, parent_slot));
    qstore(Q, Q_B0, slot, b0);
    qstore(Q, Q_E0, slot, qload(Q, Q_E0, parent_slot));
    Q.rng[slot] = make_uint4(crng.id0, crng.id1, crng.id2, crng.ctr);
    Q.n_scatt[slot] = __ldcg(Q.n_scatt + parent_slot) + 1;
    __threadfence();
    *reinterpret_cast<volatile unsigned int *>(Q.ready + slot) = A.gen_tag;
    return true;
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
    int pos, level;  /* halving state of the step in progress; pos == 0 && level == 0: at a step start */
    bool ne_pos;     /* fluid n_e > 0 at the previous evaluation */
    int status;
};

/* start of track_super_photon (reference harm_model.cpp:894-917): validate, initial opacities, dk/dlambda */
__device__ __forceinline__ bool begin_track(const TransportArgs &A, unsigned int slot, Live &L) {
    const GmParams &P = A.P;
    const PhotonQueue &Q = A.Q;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        L.x[i] = qload(Q, Q_X0 + i, slot);
        L.k[i] = qload(Q, Q_K0 + i, slot);
    }
    L.w = qload(Q, Q_W, slot);
    L.e_0_s = qload(Q, Q_E, slot);
    const uint4 r = __ldcg(Q.rng + slot);
    L.rng.id0 = r.x;
    L.rng.id1 = r.y;
    L.rng.id2 = r.z;
    L.rng.ctr = r.w;
    L.slot = slot;
    L.tau_abs = 0.0;
    L.tau_scatt = 0.0;
    L.n_step = 0;
    L.pos = 0;
    L.level = 0;
    L.status = 0;
    L.dl = 0.0;
    bool bad = (L.w == 0.0);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        bad = bad || isnan(L.x[i]) || isnan(L.k[i]);
    if (bad)
        return false;
    const GeoPoint q = geo_point(P, L.x[1], L.x[2]);
    const MetricCov g = metric_cov(P, q);
    Fluid f;
    fluid_params(P, L.x[1], L.x[2], g, q, f);
    L.ne_pos = f.n_e > 0.0;
    if (L.ne_pos) {
        double nu;
        opacities(P, L.k, f, nu, L.alpha_scatt, L.alpha_abs);
        L.bi = bias_func(P, A.bias, f.theta_e, L.w);
    } else {
        L.alpha_scatt = 0.0;
        L.alpha_abs = 0.0;
        L.bi = 0.0;
    }
    Connection c;
    connection_eval(P, q, c);
    geodesic_rhs(c, L.k, L.dk);
    return true;
}

/* work counters kept per thread and flushed once */
struct Work {
    unsigned int tracked, steps, attempts, interactions, scatters;
};

/* Interaction with the fluid after an accepted step (reference harm_model.cpp:936-1056), scattering inline.
 * `snap` points at this thread's column of the shared-memory snapshot (stride = blockDim.x).
 * Returns true if the photon is finished (absorbed or dropped). */
__device__ __forceinline__ bool interact(const TransportArgs &A, Live &L, const double *snap, int snap_stride,
                                         Work &wk) {
    const GmParams &P = A.P;
    ++wk.interactions;
    const GeoPoint q = geo_point(P, L.x[1], L.x[2]);
    const MetricCov g = metric_cov(P, q);
    Fluid f;
    fluid_params(P, L.x[1], L.x[2], g, q, f);
    const bool bound = (f.n_e == 0.0);
    double nu = 0.0, mu = 0.0;
    if (!bound) {
        double e_fluid;
        fluid_frame(P, L.k, f, e_fluid, mu);
        nu = e_fluid * kME * kCL * kCL / kHPL;
    }
    double d_tau_scatt, d_tau_abs, bias;
    if (bound || nu < 0.0) {
        d_tau_scatt = 0.5 * L.alpha_scatt * P.d_tau_k * L.dl;
        d_tau_abs = 0.5 * L.alpha_abs * P.d_tau_k * L.dl;
        L.alpha_scatt = 0.0;
        L.alpha_abs = 0.0;
        bias = 0.0;
        L.bi = 0.0;
    } else {
        const double a_sf = alpha_inv_scatt(P, nu, f.theta_e, f.n_e);
        d_tau_scatt = 0.5 * (L.alpha_scatt + a_sf) * P.d_tau_k * L.dl;
        L.alpha_scatt = a_sf;
        const double a_af = alpha_inv_abs_sin(P, nu, f.theta_e, f.n_e, f.b, sqrt(1.0 - mu * mu));
        d_tau_abs = 0.5 * (L.alpha_abs + a_af) * P.d_tau_k * L.dl;
        L.alpha_abs = a_af;
        const double bf = bias_func(P, A.bias, f.theta_e, L.w);
        bias = 0.5 * (L.bi + bf);
        L.bi = bf;
    }
    L.ne_pos = f.n_e > 0.0;
    const double x1r = -log(rng_uniform(P, L.rng));
    const double w_child = L.w / bias;
    if (bias * d_tau_scatt > x1r && w_child > kWeightMin) {
        /* ---- scattering (reference :985-1039) ---- */
        Rng crng = rng_child(P, L.rng);
        const double frac = x1r / (bias * d_tau_scatt);
        d_tau_abs *= frac;
        if (d_tau_abs > 100) {
            L.status |= 4;
            return true; /* absorbed before scattering */
        }
        d_tau_scatt *= frac;
        L.w *= attenuation(d_tau_abs + d_tau_scatt, d_tau_abs < 1.0e-3);
        /* back up: re-push the pre-step snapshot by dl * frac */
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            L.x[i] = snap[(0 + i) * snap_stride];
            L.k[i] = snap[(4 + i) * snap_stride];
            L.dk[i] = snap[(8 + i) * snap_stride];
        }
        L.e_0_s = snap[12 * snap_stride];
        wk.attempts += push_photon_full(P, L.x, L.k, L.dk, L.e_0_s, L.dl * frac);
        const GeoPoint q2 = geo_point(P, L.x[1], L.x[2]);
        const MetricCov g2 = metric_cov(P, q2);
        fluid_params(P, L.x[1], L.x[2], g2, q2, f);
        L.ne_pos = f.n_e > 0.0;
        if (L.ne_pos) {
            ++wk.scatters;
            L.status |= 2;
            ScatterChild ch;
            const bool child_ok = scatter_super_photon(P, L.rng, L.k, L.w, f, g2, ch);
            if (L.w < 1.0e-100) {
                L.status |= 4;
                return true; /* k could not be put back on the light cone (:1018-1021) */
            }
            if (child_ok)
                enqueue_child(A, L.x, ch, w_child, f.b, L.slot, crng);
            double nu2;
            opacities(P, L.k, f, nu2, L.alpha_scatt, L.alpha_abs);
            L.bi = bias_func(P, A.bias, f.theta_e, L.w);
        } else {
            /* left the grid while backing up (reference reads uninitialised data here, Appendix A.15) */
            L.alpha_scatt = 0.0;
            L.alpha_abs = 0.0;
            L.bi = 0.0;
        }
    } else {
        if (d_tau_abs > 100) {
            L.status |= 4;
            return true; /* absorbed */
        }
        const double d_tau = d_tau_abs + d_tau_scatt;
        L.w *= attenuation(d_tau, d_tau < 1.0e-3);
    }
    L.tau_abs += d_tau_abs;
    L.tau_scatt += d_tau_scatt;
    return false;
}

/* One iteration of the flattened per-photon loop: (step start bookkeeping) + one push attempt +
 * (step end: stop test, interaction).  Returns true when the photon's life is over; `record` tells whether
 * it escaped through r > r_max (reference :1066-1068). */
__device__ __forceinline__ bool advance(const TransportArgs &A, Live &L, double *snap, int snap_stride, Work &wk,
                                        bool &record) {
    const GmParams &P = A.P;
    record = false;
    if (L.pos == 0 && L.level == 0) {
        /* top of the while loop (:919) */
        if (stop_criterion(P, L.x[1], L.w, L.rng)) {
            record = L.x[1] > P.x1_max;
            return true;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            snap[(0 + i) * snap_stride] = L.x[i];
            snap[(4 + i) * snap_stride] = L.k[i];
            snap[(8 + i) * snap_stride] = L.dk[i];
        }
        snap[12 * snap_stride] = L.e_0_s;
        L.dl = step_size(P, L.x, L.k);
    }
    bool accept;
    if (L.x[1] < P.x_start1) {
        accept = true; /* push_photon is a silent no-op below the grid's inner edge (:1218-1220) */
    } else {
        double xn[4], kn[4], dkn[4], e1;
        const bool fail = push_attempt(P, L.x, L.k, L.dk, ldexp(L.dl, -L.level), L.e_0_s, xn, kn, dkn, e1);
        ++wk.attempts;
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
        return false;
    }
    L.pos += 128 >> L.level;
    if (L.pos < 128) {
        L.level = halving_next_level(L.pos);
        return false;
    }
    /* the step is complete */
    L.pos = 0;
    L.level = 0;
    ++wk.steps;
    if (stop_criterion(P, L.x[1], L.w, L.rng)) {
        record = L.x[1] > P.x1_max;
        return true;
    }
    if (L.alpha_abs > 0.0 || L.alpha_scatt > 0.0 || L.ne_pos) {
        if (interact(A, L, snap, snap_stride, wk))
            return true;
    }
    ++L.n_step;
    if (L.n_step > kMaxNStep)
        return true; /* step cap: not recorded (:1060-1066) */
    return false;
}

} /* namespace gm */
