/*
 * gm_pipeline.cuh -- the transport loop with OVERLAPPING generations: one persistent launch runs many generations
 * of primaries, the generation clock lives on the device.
 *
 * Why (round-1 profile, profiles/r1_launches_bench.csv): with one launch per generation every generation ends in a
 * drain of up to `gen_budget` loop iterations at falling occupancy, and the ~27 small generations of the start-up
 * ramp sit on a 2 - 3 ms floor each: 14 % of the configs[1] step, half of a configs[0] run.
 *
 * How.  The reference's bias_func reads running statistics (harm_model.cpp:1391-1404); this path freezes them per
 * generation.  Here a generation g marked "lag 1" (GenDesc::lag; by default the generations at the size
 * cap, i.e. the bulk of a large run, where the statistics are settled) uses the statistics of all generations
 * <= g - 2: it can start as soon as generation g - 2 is complete, so two generations are in flight and the drain of
 * one is covered by the bulk of the next.  The other generations (the start-up ramp, where the statistics still move
 * fast and a lag shows in the scattering counts: +1.3 % at configs[0], profiles/r2_pipeline_parity.txt) start when
 * their predecessor is complete, exactly as with one launch per generation -- but without a host round trip.  Everything a photon does still depends only on its identity and on its generations'
 * statistics -- results do not depend on scheduling, launch geometry or how many generations share a launch:
 *   - a record carries its generation tag (g & 3; bits 28-29 of the n_step word); at most generations nc, nc + 1
 *     (open) and nc + 2 (being filled by suspensions) exist at a time, nc = number of complete generations;
 *   - per-generation counters GenCtl::alloc / done (records that entered / left the generation: primaries, children,
 *     carried-in lineages; finished or suspended) tell when a generation is complete; the thread that sees it folds
 *     the generation's statistics into the run totals, freezes the bias denominator of generation nc + 2, opens it
 *     and publishes nc + 1 (gen_try_complete, serialised by a try-lock);
 *   - a lineage that uses up its attempt budget in generation g is rewritten with tag g + 1 and waits on the LIMBO
 *     queue of that parity until g + 1 is open (it is once g - 1 is complete);
 *   - work sources of a lane, oldest generation first: limbo and runnable queue of the older open generation, then
 *     of the newer one, then the primaries (born by birth_kernel before the launch, taken by index in generation
 *     order up to the end of the newest open generation).  All claims are fetch-adds (never fail); a lane whose
 *     claim ran ahead of what is published keeps its ticket and polls it.
 * The oracle follows the same rule (orc_run with stats_lag = 1): tests/test_gpu_parity.py compares photon by photon.
 */
#pragma once
#include "gm_kernels.cuh"

namespace gm {

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long *p, unsigned long long v) {
    *reinterpret_cast<volatile unsigned long long *>(p) = v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

/* The generation clock.  Called by one thread per block and outer iteration: if the oldest incomplete generation has
 * no record left, complete it (and any empty generation behind it). */
__device__ __noinline__ void gen_try_complete(const TransportArgs *Ag) {
    const TransportArgs &A = *Ag;
    GenCtl *C = A.ctl;
    for (int rounds = 0; rounds < 8; ++rounds) {
        const unsigned long long nc = ld_acquire_u64(C->line + CW_COMPLETE);
        if ((long long)nc >= (long long)C->g_end)
            return;
        const int t = (int)(nc & 3ull);
        const unsigned long long d = ld_volatile_u64(C->done + t);
        __threadfence();
        const unsigned long long a = ld_volatile_u64(C->alloc + t);
        if (d != a)
            return; /* records of the generation are still alive (every increment of alloc precedes the done that could
                     * make the two equal: children are counted before their parent can finish, suspended lineages in
                     * the next generation before they leave this one) */
        if (atomicCAS(&C->lock, 0u, 1u) != 0u)
            return; /* another block is completing it */
        __threadfence();
        if (ld_volatile_u64(C->line + CW_COMPLETE) == nc) {
            const long long g = (long long)nc;
            /* run totals <- the generation's statistics (only the holder of the lock writes them) */
            unsigned long long *cnt = A.A.counters;
            const unsigned long long n_scatt = ld_volatile_u64(cnt + 1) + ld_volatile_u64(C->acc_scatt + t);
            const unsigned long long n_rec = ld_volatile_u64(cnt + 2) + ld_volatile_u64(C->acc_rec + t);
            const unsigned long long mt0 = ld_volatile_u64(A.A.max_tau_bits), mt1 = ld_volatile_u64(C->acc_maxtau + t);
            const unsigned long long mt = mt0 > mt1 ? mt0 : mt1; /* non-negative doubles order like their bit patterns */
            st_volatile_u64(cnt + 1, n_scatt);
            st_volatile_u64(cnt + 2, n_rec);
            st_volatile_u64(A.A.max_tau_bits, mt);
            /* bias denominator of the generations that open now: everything recorded in generations <= g (operation
             * order of make_bias_stats, gm_params.h) */
            const double avg = __ddiv_rn((double)n_scatt, __dadd_rn((double)n_rec, 1.0));
            const double den = __dmul_rn(__dmul_rn(C->bias_norm, __longlong_as_double((long long)mt)), __dadd_rn(avg, 2.0));
            /* the ring slot now belongs to generation g + 4 */
            st_volatile_u64(C->acc_scatt + t, 0ull);
            st_volatile_u64(C->acc_rec + t, 0ull);
            st_volatile_u64(C->acc_maxtau + t, 0ull);
            st_volatile_u64(C->done + t, 0ull);
            st_volatile_u64(C->alloc + t, g + 4 < C->n_desc ? C->desc[g + 4].count : 0ull);
            /* Which generations open now?  g + 1 unless it was opened one generation early (lag 1: at the completion
             * of g - 1); g + 2 if IT may start early.  Only generations of this launch open: nothing of a generation
             * >= g_end may run here (the launch ends when g_end - 1 is complete). */
            GenDesc *dsc = const_cast<GenDesc *>(C->desc);
            const unsigned long long now = global_timer_ns();
            dsc[g].t_done = now;
            const int q = (int)(g & 1), q1 = q ^ 1;
            const bool has1 = g + 1 < C->n_desc, has2 = g + 2 < C->n_desc && C->desc[g + 2].lag == 1;
            const bool in1 = has1 && g + 1 < (long long)C->g_end, in2 = has2 && g + 2 < (long long)C->g_end;
            /* (the denominators are frozen here even for generations of the next launch: the ring lives on) */
            if (has1 && C->desc[g + 1].lag == 0)
                *reinterpret_cast<volatile double *>(C->bias_den + ((g + 1) & 3)) = den;
            if (has2)
                *reinterpret_cast<volatile double *>(C->bias_den + ((g + 2) & 3)) = den;
            if (in1) {
                if (C->desc[g + 1].lag == 0)
                    dsc[g + 1].t_open = now;
                /* limbo queue of g + 1's parity: all entries for g + 1 are in (g is complete) and the next ones, for
                 * g + 3, must wait for that generation: the gate is (re)set to the current tail */
                st_volatile_u64(C->line + CW_L0_LIM + kCwLStride * q1, ld_volatile_u64(C->line + CW_L0_TAIL + kCwLStride * q1));
                st_volatile_u64(C->line + CW_PRIM_LIM, C->desc[g + 1].prim_end);
            }
            if (in2) {
                dsc[g + 2].t_open = now;
                /* its limbo entries come from g + 1, which is running: every published entry may be taken */
                st_volatile_u64(C->line + CW_L0_LIM + kCwLStride * q, kGateLive);
                st_volatile_u64(C->line + CW_PRIM_LIM, C->desc[g + 2].prim_end);
            }
            __threadfence();
            st_release_u64(C->line + CW_COMPLETE, nc + 1ull);
        }
        __threadfence();
        atomicExch(&C->lock, 0u);
    }
}

__device__ __noinline__ void record_call_pipe(const TransportArgs *Ag, unsigned int slot, double x2, double x3, double w,
                                              double tau_abs, double tau_scatt, int tag) {
    record_super_photon<true>(*Ag, slot, x2, x3, w, tau_abs, tau_scatt, tag);
}

constexpr int kTicketShift = 56;

/* publish a warp's finished / suspended counts (16 bits per generation tag in done_pk); all lanes call it together */
__device__ __noinline__ void pipe_flush_done(GenCtl *C, unsigned long long done_pk) {
    __threadfence(); /* records, children and suspensions of these photons come first */
    unsigned long long v = done_pk;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const unsigned long long c = (v >> (16 * t)) & 0xffffull;
            if (c)
                atomicAdd(C->done + t, c);
        }
    }
}


#ifndef GM_PIPE_CLAIM_EVERY
#define GM_PIPE_CLAIM_EVERY 4 /* claim only in every n-th loop iteration (power of two): fewer, larger claims;
                               * measured 1 -> 593, 2 -> 592, 4 -> 574 ms per configs[1] run */
#endif

constexpr int kDoneFlushEvery = 8; /* inner iterations between two publications of a warp's finished counts */

template <int BLOCK, int MIN_BLOCKS>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS) pipeline_kernel(const TransportArgs A) {
    __shared__ unsigned long long s_base; /* first scatter-queue position claimed for the block */
    __shared__ int s_count;               /* number of scatter-queue entries claimed (0: no service) */
    __shared__ int s_quit;
    __shared__ double s_bias[4];          /* bias denominators by generation tag (copy of GenCtl::bias_den) */
    /* work hints, refreshed by thread 0 once per outer iteration: which sources had work (bit 0 / 1: limbo and
     * runnable queue of the older open generation, 2 / 3: of the newer one, 4: primaries), the older generation's
     * parity, and the end of the issuable primaries (it only grows: a stale copy is conservative).  A warp that needs
     * work reads ONE source's counters, the first one hinted; the control lines are touched by claims and pushes only */
    __shared__ volatile int s_hint, s_pa;
    __shared__ volatile unsigned long long s_prim_lim;
    const int lane = threadIdx.x & 31;
    double *snap = gm_smem + threadIdx.x; /* 13 rows of BLOCK doubles */
    /* Record stage compaction as in transport_kernel (gm_kernels.cuh): a lane whose photon escaped leaves what the
     * record needs here (x2 x3 w tau_abs tau_scatt, slot / generation tag) and takes its next photon; the warp makes
     * the records of kRecordBatch lanes together -- and at the end of every outer iteration at the latest, because
     * here the photon counts as DONE in its generation only once its record is made (the generation's statistics are
     * folded into the run totals when done == alloc). */
    double *pend = gm_smem + (size_t)13 * BLOCK + threadIdx.x;
    bool pending = false;
    GenCtl *const C = A.ctl;
    unsigned long long *const line = C->line;
    Live L;
    bool has = false;
    long long ticket = -1; /* kind << 56 | position: 0 / 1 runnable queue, 2 / 3 limbo queue (by parity), 4 primary */
    Work wk = {0u, 0u, 0u, 0u, 0u, 0u, 0u};
    unsigned long long done_pk = 0ull; /* finished / suspended records by generation tag, 16 bits each */
    unsigned int idle_spins = 0;
    bool was_idle = true;
    if (threadIdx.x < 4)
        s_bias[threadIdx.x] = __ldcg(C->bias_den + threadIdx.x);
    if (threadIdx.x == 0) {
        s_hint = 0;
        s_pa = 0;
        s_prim_lim = 0ull;
    }
    if (threadIdx.x == 0 && blockIdx.x == 0)
        atomicCAS(&C->t_start, 0ull, global_timer_ns());
    __syncthreads();

    for (;;) {
        /* ---- block control (one thread): the generation clock, and parked photons for the scattering stage ---- */
        if (threadIdx.x == 0) {
            gen_try_complete(A.self);
            int cnt = 0;
            const unsigned long long h = ld_volatile_u64(A.scatter.head);
            unsigned long long t = ld_volatile_u64(A.scatter.tail);
            const unsigned long long avail = t > h ? t - h : 0ull;
            if (avail >= (unsigned long long)BLOCK || (was_idle && avail > 0)) {
                cnt = avail < (unsigned long long)BLOCK ? (int)avail : BLOCK;
                if (atomicCAS(A.scatter.head, h, h + cnt) == h)
                    s_base = h;
                else
                    cnt = 0;
            }
            s_count = cnt;
            /* work hints for the block's warps */
            {
                const int pa = (int)(ld_volatile_u64(line + CW_COMPLETE) & 1ull);
                int hint = 0;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int p = k ^ pa;
                    const unsigned long long lh = ld_volatile_u64(line + CW_L0_HEAD + kCwLStride * p);
                    unsigned long long lt = ld_acquire_u64(line + CW_L0_TAIL + kCwLStride * p);
                    const unsigned long long ll = ld_volatile_u64(line + CW_L0_LIM + kCwLStride * p);
                    lt = lt < ll ? lt : ll;
                    if (lt > lh)
                        hint |= 1 << (2 * k);
                    if (ld_volatile_u64(line + CW_R0_TAIL + kCwRStride * p) > ld_volatile_u64(line + CW_R0_HEAD + kCwRStride * p))
                        hint |= 2 << (2 * k);
                }
                const unsigned long long pl = ld_volatile_u64(line + CW_PRIM_LIM);
                if (pl > ld_volatile_u64(line + CW_PRIM_CUR))
                    hint |= 16;
                s_pa = pa;
                s_prim_lim = pl;
                s_hint = hint;
            }
        }
        __syncthreads();
        /* ---- scattering stage, all lanes of the block that got a parked photon ---- */
        {
            const int cnt = s_count;
            if ((int)threadIdx.x < cnt) {
                unsigned int *ent = A.scatter.entries + ((s_base + threadIdx.x) & (unsigned long long)(A.scatter.capacity - 1u));
                unsigned int v, spins = 0;
                while ((v = ld_volatile_u32(ent)) == 0u) {
                    if (++spins > (1u << 26)) {
                        atomicOr(A.A.error, 2u);
                        break;
                    }
                }
                __threadfence();
                if (v) {
                    *reinterpret_cast<volatile unsigned int *>(ent) = 0u; /* ring: the entry is free again */
                    const ScatterStageResult sr = scatter_stage_t<true>(A.self, v - 1u);
                    if (sr.done)
                        done_pk += 1ull << (16 * sr.tag);
                    wk.attempts += sr.attempts;
                    wk.scatters += sr.scatters;
                    wk.tracked += sr.children;
                }
            }
        }
#pragma unroll 1
        for (int sub = 0; sub < kItersPerSync; ++sub) {
            /* ---- refill: lanes without a photon and without a ticket claim one (pipe_claim) ---- */
            {
                const bool want = !has && ticket < 0;
                const unsigned int need = __ballot_sync(0xffffffffu, want);
                if (need && (GM_PIPE_CLAIM_EVERY == 1 || (sub & (GM_PIPE_CLAIM_EVERY - 1)) == 0)) {
                    const int hint = s_hint;
                    if (hint) {
                        /* the first hinted source: its counters are read by lane 0 only (one line), which claims for
                         * the warp with one fetch-add -- never fails; a claim that ran ahead of what is published is a
                         * ticket the lane keeps polling.  (Measured and rejected, profiles/r2_pipeline_ab.txt: claims
                         * capped at the warp's share of what is available; block-local exact claims refilled once per
                         * outer iteration.) */
                        const int src = __ffs(hint) - 1;
                        const int p = ((src >> 1) & 1) ^ s_pa;
                        const int kind = src == 4 ? 4 : ((src & 1) ? p : 2 + p);
                        unsigned long long base = 0ull;
                        int n = 0;
                        if (lane == 0) {
                            unsigned long long *hp = line + (src == 4 ? CW_PRIM_CUR
                                                                      : (src & 1) ? CW_R0_HEAD + kCwRStride * p
                                                                                  : CW_L0_HEAD + kCwLStride * p);
                            const unsigned long long h = ld_volatile_u64(hp);
                            unsigned long long t = ld_acquire_u64(hp + 1); /* tail (primaries: limit) before the gate */
                            if (src < 4 && !(src & 1)) {
                                const unsigned long long lim = ld_volatile_u64(hp + 2);
                                t = t < lim ? t : lim;
                            }
                            const unsigned long long avail = t > h ? t - h : 0ull;
                            const int n_want = __popc(need);
                            n = avail < (unsigned long long)n_want ? (int)avail : n_want;
                            if (n > 0)
                                base = atomicAdd(hp, (unsigned long long)n);
                            else
                                atomicAnd(const_cast<int *>(&s_hint), ~(1 << src)); /* stale hint */
                        }
                        base = __shfl_sync(0xffffffffu, base, 0);
                        n = __shfl_sync(0xffffffffu, n, 0);
                        if (want) {
                            const int my = __popc(need & ((1u << lane) - 1u));
                            if (my < n)
                                ticket = ((long long)kind << kTicketShift) | (long long)(base + my);
                        }
                    }
                }
                if (!has && ticket >= 0) {
                    const int kind = (int)(ticket >> kTicketShift);
                    const unsigned long long pos = (unsigned long long)ticket & ((1ull << kTicketShift) - 1ull);
                    unsigned int slot = 0u;
                    bool ok = false;
                    if (kind == 4) {
                        ok = pos < s_prim_lim; /* the block's copy lags by at most one outer iteration */
                        slot = (unsigned int)pos;
                    } else {
                        /* acquire: pairs with the producer's fence + store in queue_push_ring */
                        unsigned int *ent = A.qent + (size_t)kind * A.qcap + (pos & (unsigned long long)(A.qcap - 1u));
                        const unsigned int v = ld_acquire_u32(ent);
                        ok = v != 0u;
                        /* limbo: the entry first, then the gate (an entry of a generation that is not open yet was
                         * published after its queue's gate had been frozen below it) */
                        if (ok && kind >= 2)
                            ok = pos < ld_volatile_u64(line + CW_L0_LIM + kCwLStride * (kind - 2));
                        slot = v - 1u;
                        if (ok)
                            *reinterpret_cast<volatile unsigned int *>(ent) = 0u; /* ring: the entry is free again */
                    }
                    if (ok) {
                        ticket = -1;
                        const int ns = live_load(A, slot, L);
                        const int tag = (ns >> kTagShift) & 3;
                        L.n_step = ns & kNStepMask;
                        L.status = tag << 4;
                        /* the generation is open, so its bias denominator is final: refresh the block's copy (all
                         * writers store the same value) and give a fresh primary its bias */
                        const double den = __ldcg(C->bias_den + tag);
                        s_bias[tag] = den;
                        if (ns & kFreshBit)
                            L.bi = bias_func_den(L.bi, L.w, den);
                        bool bad = (L.w == 0.0);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            bad = bad || isnan(L.x[i]) || isnan(L.k[i]);
                        if (bad)
                            done_pk += 1ull << (16 * tag); /* invalid photon (reference :895-900): dropped */
                        else
                            has = true;
                    }
                }
            }
            const unsigned int live_mask = __ballot_sync(0xffffffffu, has);
            if (!live_mask)
                break; /* nothing to do in this warp: go to the barrier */
            ++wk.slot_iters;
            if (has) {
                ++wk.live_iters;
                bool record;
                const StepResult r = advance<true>(A, L, live_mask, snap, BLOCK, wk, record, s_bias);
                if (r == STEP_FINISHED) {
                    if (record && kRecordBatch > 0) {
                        if (pending) { /* rare: a second photon escaped before the first was recorded */
                            const int2 st = *reinterpret_cast<const int2 *>(pend + 5 * BLOCK);
                            record_call_pipe(A.self, (unsigned int)st.x, pend[0], pend[BLOCK], pend[2 * BLOCK],
                                             pend[3 * BLOCK], pend[4 * BLOCK], st.y);
                            done_pk += 1ull << (16 * st.y);
                        }
                        pend[0] = L.x[2];
                        pend[BLOCK] = L.x[3];
                        pend[2 * BLOCK] = L.w;
                        pend[3 * BLOCK] = L.tau_abs;
                        pend[4 * BLOCK] = L.tau_scatt;
                        *reinterpret_cast<int2 *>(pend + 5 * BLOCK) = make_int2((int)L.slot, (L.status >> 4) & 3);
                        pending = true; /* counted as done when the record is made */
                    } else {
                        if (record)
                            record_call_pipe(A.self, L.slot, L.x[2], L.x[3], L.w, L.tau_abs, L.tau_scatt, (L.status >> 4) & 3);
                        done_pk += 1ull << (16 * ((L.status >> 4) & 3));
                    }
                    has = false;
                } else if (r == STEP_SCATTER) {
                    has = false; /* parked for the scattering stage */
                } else if (r == STEP_SUSPEND) {
                    suspend_photon_pipe(A.self, L.slot, L.x[0], L.x[1], L.x[2], L.x[3], L.k[0], L.k[1], L.k[2], L.k[3],
                                        L.dk[0], L.dk[1], L.dk[2], L.dk[3], L.w, L.e_0_s, L.tau_abs, L.tau_scatt,
                                        L.alpha_scatt, L.alpha_abs, L.bi, L.ne_pos, L.rng.id0, L.rng.id1, L.rng.id2,
                                        L.rng.ctr, L.n_step, (L.status >> 4) & 3);
                    has = false;
                    done_pk += 1ull << (16 * ((L.status >> 4) & 3)); /* left this generation */
                }
            }
            if (kRecordBatch > 0 && __popc(__ballot_sync(0xffffffffu, pending)) >= kRecordBatch && pending) {
                const int2 st = *reinterpret_cast<const int2 *>(pend + 5 * BLOCK);
                record_call_pipe(A.self, (unsigned int)st.x, pend[0], pend[BLOCK], pend[2 * BLOCK], pend[3 * BLOCK],
                                 pend[4 * BLOCK], st.y);
                done_pk += 1ull << (16 * st.y);
                pending = false;
            }
            if ((sub & (kDoneFlushEvery - 1)) == kDoneFlushEvery - 1 && __ballot_sync(0xffffffffu, done_pk != 0ull)) {
                pipe_flush_done(C, done_pk);
                done_pk = 0ull;
            }
        }
        if (pending) { /* end of the outer iteration: no record waits longer than this */
            const int2 st = *reinterpret_cast<const int2 *>(pend + 5 * BLOCK);
            record_call_pipe(A.self, (unsigned int)st.x, pend[0], pend[BLOCK], pend[2 * BLOCK], pend[3 * BLOCK],
                             pend[4 * BLOCK], st.y);
            done_pk += 1ull << (16 * st.y);
            pending = false;
        }
        if (__ballot_sync(0xffffffffu, done_pk != 0ull)) {
            pipe_flush_done(C, done_pk);
            done_pk = 0ull;
        }
        const int block_live = __syncthreads_or(has ? 1 : 0);
        if (!block_live) {
            if (threadIdx.x == 0) {
                const unsigned long long nc = ld_volatile_u64(line + CW_COMPLETE);
                unsigned int err = ld_volatile_u32(A.A.error);
                if (global_timer_ns() - ld_volatile_u64(&C->t_start) > C->t_limit) {
                    atomicOr(A.A.error, 8u); /* watchdog: something is lost; end the launch instead of hanging */
                    err |= 8u;
                }
                s_quit = ((long long)nc >= (long long)C->g_end) || (err & (1u | 2u | 8u));
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

} /* namespace gm */
