/*
 * grmonty_b200.h -- C ABI of the B200-native superphoton transport path.
 *
 * This is the drop-in boundary.  The reference has no plugin/FFI interface; its only
 * seam is the `#ifdef CUDA` block of HARMModel::run_simulation (reference cuda_grmonty/harm_model.cpp:345-362)
 * that calls three C++ free functions declared in cuda_grmonty/super_photon.cuh:29-61:
 *
 *     cuda_super_photon::alloc_memory(header, data, units, hotcross_table, f, k2)     -> grmonty_b200_create
 *     cuda_super_photon::track_super_photons(bias_norm, max_tau_scatt, photon_queue,
 *                                            stop_sem, spectrum, n_recorded, n_scatt) -> grmonty_b200_run
 *                                                                                       + grmonty_b200_result
 *     cuda_super_photon::free_memory()                                                -> grmonty_b200_destroy
 *
 * Differences by design: photon generation (reference harm_model.cpp:673-811, host threads feeding a queue
 * in the reference GPU build, :842-892) happens on the device, so there is no photon queue/semaphore in the
 * interface; errors are returned as codes (the reference's gpuErrchk calls exit(), utils.cuh:33-40); a context
 * object replaces the reference's file-scope statics (super_photon.cu:36-71), so several contexts (one per
 * GPU) can live in one process.
 *
 * Conventions: plain C structs and pointers, no C++ or torch types.  The caller owns every input buffer and
 * may free it as soon as grmonty_b200_create returns (inputs are copied to the device).  Outputs are
 * caller-allocated.  All functions return 0 on success or a negative GRMONTY_B200_E* code and never throw or
 * exit.  A context is bound to one CUDA device and must be used by one host thread at a time.
 * There is NO CPU fallback: without a CUDA device grmonty_b200_create fails with GRMONTY_B200_ECUDA.
 */
#ifndef GRMONTY_B200_H
#define GRMONTY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRMONTY_B200_ABI_VERSION 4

#define GRMONTY_B200_N_TH_BINS 6     /* reference consts.hpp:26 */
#define GRMONTY_B200_N_E_BINS 200    /* reference consts.hpp:25 */
#define GRMONTY_B200_SPEC_FIELDS 13  /* reference harm_data.hpp:129-143, same field order */
#define GRMONTY_B200_HOTCROSS_N (221 * 81)
#define GRMONTY_B200_TABLE_N 201
#define GRMONTY_B200_NINT_N 20001
#define GRMONTY_B200_NCCL_ID_BYTES 128 /* sizeof(ncclUniqueId) */

enum {
    GRMONTY_B200_OK = 0,
    GRMONTY_B200_EINVAL = -1,   /* bad argument / config */
    GRMONTY_B200_ECUDA = -2,    /* CUDA runtime error (message in last_error) */
    GRMONTY_B200_EQUEUE = -3,   /* device photon queue overflowed: raise queue_capacity or lower gen_cap */
    GRMONTY_B200_ENCCL = -4,    /* NCCL not loadable / collective failed */
    GRMONTY_B200_ESTATE = -5    /* call order violation */
};

typedef struct grmonty_b200_ctx grmonty_b200_ctx;

/* Everything the transport path reads; mirrors what HARMModel holds after read_file() + init(). */
typedef struct grmonty_b200_config {
    uint32_t abi_version; /* GRMONTY_B200_ABI_VERSION */
    uint32_t struct_size; /* sizeof(grmonty_b200_config) */

    /* harm::Header subset, reference harm_data.hpp:19-44 */
    int32_t n0, n1;
    double x_start1, x_start2; /* x_start[1..2] */
    double dx1, dx2, dx3;      /* dx[1..3] */
    double x_stop1, x_stop2;   /* x_stop[1..2] */
    double a, h_slope, r_0;

    /* harm::Units, reference harm_data.hpp:63-72 */
    double mass_unit, l_unit, t_unit, rho_unit, u_unit, b_unit, theta_e_unit, n_e_unit;

    /* harm::Data, reference harm_data.hpp:49-58: eight row-major [n0][n1] host arrays */
    const double *k_rho, *u, *u_1, *u_2, *u_3, *b_1, *b_2, *b_3;
    /* geometry_.det, reference harm_model.cpp:261: sqrt|det g_cov| at zone centres, row-major [n0][n1] */
    const double *geom_det;

    /* tables, reference harm_model.hpp:159-195 */
    const double *hotcross;   /* [221][81] log10 of the hot cross-section, hotcross.cpp:60-79 */
    const double *f;          /* [201] ln F(K), jnu_mixed.cpp:57-65 */
    const double *k2;         /* [201] ln K2(1/theta_e), jnu_mixed.cpp:67-70 */
    const double *weight;     /* [201] ln weight, harm_model.cpp:268-306 */
    const double *nint;       /* [20001] ln nint, harm_model.cpp:308-338 */
    const double *dndlnu_max; /* [20001] ln dndlnu_max */

    /* scalars, reference harm_model.hpp:104-125 */
    double photon_n;       /* estimate of photon number (CLI -photon_n) */
    double bias_norm;      /* harm_model.cpp:206,219 */
    double max_tau_scatt0; /* initial max_tau_scatt, harm_model.cpp:72 */
    uint64_t seed;         /* Philox key; the reference seeds mt19937 with 123 (main.cpp:49) */

    /* sharding: this context tracks the primaries whose global index i satisfies i % world == rank */
    int32_t rank, world;
    int32_t device; /* CUDA device ordinal */

    /* tuning, 0 = default */
    int32_t threads_per_block;
    int32_t blocks_per_sm;
    int64_t queue_capacity; /* photon slots in the device queue */
    /* Generation schedule.  The scattering-bias statistics are frozen within a generation, so the generations
     * must stay short relative to the run so far (the reference updates its statistics after every photon):
     * the generation starting at run position s holds gen0 positions if s < gen0, (gen_ramp - 1) s positions (the
     * cumulative count grows gen_ramp-fold) while s < gen_fine_from, then s / gen_fine_div positions, and never
     * more than gen_cap.  Positions are rank-local here (global position / world), so each GPU of a multi-GPU job
     * runs the schedule of a stand-alone run of its own share. */
    int64_t gen0;           /* default 32 */
    int64_t gen_cap;        /* default 2^20 */
    int64_t gen_budget;     /* push attempts a photon lineage may make per generation before it is carried over
                               to the next one (default 384); bounds the tail of every generation */
    int64_t gen_budget_spread; /* > 0: the t-th of a generation's n primaries (rank-local) may make gen_budget +
                               (n - 1 - t) / gen_budget_spread attempts and a carried lineage gen_budget + n /
                               gen_budget_spread: lanes take the primaries in order, so a lineage that starts early
                               can run for the rest of the generation without delaying its end (~100 attempts per
                               primary on 37 888 lanes = one attempt per 384 positions); negative: off */
    int64_t gen_fine_from;  /* default 16384 */
    int64_t gen_ramp;       /* default 8 */
    int64_t gen_fine_div;   /* default 6 (each generation adds 1/6 to the run so far); <= 1: keep doubling.
                               Measured at configs[0] against 20 reference runs (profiles/r1_bias_schedule.txt):
                               doubling with budget 256 gives +5.8 % scattered / +2.5 % recorded counts,
                               div 6 with budget 384 gives +0.6 % / +0.1 %. */
    /* Transport kernel.  0 (default) or 1: the fused per-lane loop (threads_per_block x blocks_per_sm as compiled in
     * csrc/gm_api.cu; default 32 x 8: one warp per block, eight blocks per SM).  2: the state-compacting wavefront kernel (csrc/gm_wavefront.cuh) -- photons
     * resident in shared memory, `slots_per_thread` of them per thread (default 2), the block runs push / interact /
     * service / scatter phases chosen by lane counts; threads_per_block 128 ... 512 (default 384).  The wavefront
     * kernel is the slower of the two on B200 (profiles/r2_wavefront_vs_fused.txt) and is kept as a selectable
     * alternative.  Results do not depend on the choice (tests/test_gpu_invariance.py). */
    int32_t kernel;
    int32_t slots_per_thread;
    /* wavefront phase thresholds in 1/256 (0 = default): see TransportArgs in csrc/gm_transport.cuh */
    int32_t wf_thr_interact, wf_thr_service;
    /* Generation scheduler.  0 (default) or 2: one launch per generation, with the statistics of all earlier
     * generations (the round-1 scheduler; the only one the wavefront kernel has).
     * 1: the PIPELINED scheduler (csrc/gm_pipeline.cuh; fused kernel only) -- one persistent launch runs a window of
     * generations and the generation clock lives on the device.  The generations at the size cap (gen_cap: the bulk
     * of a large run, where the bias statistics are settled) start one generation early, with the statistics of the
     * generations <= g - 2, so that the drain of one generation is covered by the bulk of the next; the generations of
     * the start-up ramp start when their predecessor is complete -- for a run without capped generations the integer
     * results are those of scheduler 2.  3: pipelined with EVERY generation starting early.
     * All are deterministic and independent of launch geometry, GPU scheduling and queue_capacity.  Measured
     * (profiles/r2_pipeline_ab.txt): scheduler 3 runs configs[0] 32 % and configs[1] 2 % faster than scheduler 2 but
     * its scattering counts stand 1.3 % above the reference ensemble at configs[0] (the statistics lag where they move
     * fastest) -- outside the 1 % bar, so it is not the default; scheduler 1 keeps the counts but is slower than 2. */
    int32_t gen_overlap;
    int32_t reserved0;
} grmonty_b200_config;

/* Device-side work counters and timings (filled by grmonty_b200_result when `stats` is not NULL). */
typedef struct grmonty_b200_stats {
    uint64_t n_tracked;        /* photons tracked (primaries + scattered children) */
    uint64_t n_steps;          /* accepted geodesic steps */
    uint64_t n_push_attempts;  /* push_photon attempts incl. halved sub-steps and scatter back-ups */
    uint64_t n_interactions;   /* in-fluid interaction evaluations (reference harm_model.cpp:937 gate) */
    uint64_t n_scatter_events; /* scatter_super_photon calls */
    uint64_t n_generations;
    uint64_t n_kernel_launches;
    uint64_t queue_high_water;
    uint64_t n_live_iterations; /* loop iterations executed by a lane that held a live photon ... */
    uint64_t n_slot_iterations; /* ... out of all loop iterations of all lanes: lane occupancy.  Counted by the test
                                   library and the optional kernels only; the product's default kernel reports 0 (the
                                   counters cost 1 % of the run time) */
    double kernel_ms;          /* device time of all kernels of the last run (CUDA events) */
    double transport_ms;       /* of which: the persistent transport kernel */
} grmonty_b200_stats;

/* ---- product entry points ---------------------------------------------------------------------------- */

/* Allocate device state, copy the model to the device, build per-zone emission data (init_zone / tetrads,
 * reference harm_model.cpp:1337-1389, :717-731) and the zone -> primary-index prefix table. */
int grmonty_b200_create(grmonty_b200_ctx **out, const grmonty_b200_config *cfg);

/* Total number of primary superphotons of the whole job (all ranks), = sum of the per-zone counts of
 * reference harm_model.cpp:673-704. */
int grmonty_b200_total_primaries(grmonty_b200_ctx *ctx, int64_t *total);

/* Generate, transport and record this rank's share (j % world == rank) of positions [first, last) of the
 * processing sequence (last < 0: to the end); position j handles primary (j * mult) mod total, a fixed
 * permutation of the zone-ordered primaries (see DESIGN.md "processing order").  Blocking.
 * Replaces the CPU loop at reference harm_model.cpp:366-404. */
int grmonty_b200_run_range(grmonty_b200_ctx *ctx, int64_t first, int64_t last);
/* = run_range(ctx, 0, -1) */
int grmonty_b200_run(grmonty_b200_ctx *ctx);

/* Optional progress callback: called on the host thread of grmonty_b200_run / run_range after every generation with
 * the number of positions of the processing sequence completed so far and their total (all ranks).  What the
 * reference logs once a second from its CPU loop (harm_model.cpp:397-403).  cb == NULL: off. */
typedef void (*grmonty_b200_progress_fn)(void *user, int64_t positions_done, int64_t positions_total);
int grmonty_b200_set_progress(grmonty_b200_ctx *ctx, grmonty_b200_progress_fn cb, void *user);

/* Sum the spectrum and counters (and max the scattering depth) over the ranks of `nccl_comm`
 * (an ncclComm_t passed as void*), in place on the device accumulators: the path's only collective.
 * NCCL is resolved at run time (dlopen; the copy already loaded in the process if there is one); world == 1: no-op. */
int grmonty_b200_allreduce(grmonty_b200_ctx *ctx, void *nccl_comm, void *cuda_stream);

/* Communicators for grmonty_b200_allreduce, for hosts that do not have an ncclComm_t of their own (the reference's
 * GPU build is single-GPU and has none).  Thin forwards to ncclGetUniqueId / ncclCommInitRank / ncclCommInitAll /
 * ncclCommDestroy of the NCCL copy grmonty_b200_allreduce uses:
 *   one process per GPU (torchrun, MPI): rank 0 calls nccl_unique_id, ships the GRMONTY_B200_NCCL_ID_BYTES bytes to
 *     the other ranks by whatever means the launcher offers, every rank calls nccl_comm_init_rank;
 *   one process, one host thread per GPU (the command line's --gpus N): nccl_comm_init_all (devices == NULL: 0..n-1). */
int grmonty_b200_nccl_unique_id(void *id);
int grmonty_b200_nccl_comm_init_rank(void **comm, const void *id, int rank, int world, int device);
int grmonty_b200_nccl_comm_init_all(void **comms, int n_devices, const int *devices);
int grmonty_b200_nccl_comm_destroy(void *comm);

/* Device pointers of the accumulators, for hosts that do the reduction with their own collective library
 * (e.g. torch.distributed):  spectrum: double[6*200*13]; counters: uint64[3] = created, scattered, recorded;
 * max_tau_scatt: double[1]. */
int grmonty_b200_device_accumulators(grmonty_b200_ctx *ctx, void **spectrum, void **counters, void **max_tau_scatt);

/* Copy results to the host.  spectrum: [6][200][13] doubles in harm::Spectrum field order;
 * counts: created, scattered (sum of n_scatt over recorded photons), recorded -- the three counters the
 * reference logs (harm_model.cpp:409-413).  Any pointer may be NULL. */
int grmonty_b200_result(grmonty_b200_ctx *ctx, double *spectrum, uint64_t counts[3], double *max_tau_scatt,
                        grmonty_b200_stats *stats);

/* Zero the accumulators and counters so the same context can run again. */
int grmonty_b200_reset(grmonty_b200_ctx *ctx);

/* Frees the context.  The photon pool (the bulk of the device memory, ~1.7 GB at the default capacity) is parked in a
 * per-device cache and reused by the next grmonty_b200_create with the same queue_capacity: cudaFree of it costs
 * more than a whole run.  grmonty_b200_trim_cache() returns the cached memory to the driver. */
void grmonty_b200_destroy(grmonty_b200_ctx *ctx);
void grmonty_b200_trim_cache(void);

/* Message of the last error on this context (ctx == NULL: of the last failed context-free call on this thread). */
const char *grmonty_b200_last_error(grmonty_b200_ctx *ctx);

/* Hot Compton cross-section table on the device: table[221][81] = log10 of the numeric integral of reference
 * hotcross.cpp:108-142 on the grid of consts.hpp:97-112 (what hotcross::init_table fills, hotcross.cpp:60-79, 33 s on
 * one CPU core).  Needs no context; `table` is a host buffer of GRMONTY_B200_HOTCROSS_N doubles. */
int grmonty_b200_hotcross_table(int device, double *table);

/* Grid-dependent initialisation tables on the device (what HARMModel::init_geometry, init_weight_table and
 * init_nint_table compute: reference harm_model.cpp:242-266, :268-306, :308-338).  Reads from `cfg` the grid geometry,
 * the units, dx3, photon_n, the eight primitive grids and the F(K) / K2 tables (cfg->f, cfg->k2; everything else,
 * including cfg->geom_det / weight / nint / dndlnu_max, is ignored) and writes
 *   geom_det[n0*n1]  sqrt|det g_cov| at zone centres,        weight[201]  ln of the super-photon weight per frequency,
 *   nint[20001], dndlnu_max[20001]  ln of the zone photon-number integral and of max dN/dln(nu) per B theta_e^2.
 * Any output may be NULL.  `device_ms` (optional) receives the device time of the three kernels.  Needs no context. */
int grmonty_b200_init_tables(const grmonty_b200_config *cfg, double *geom_det, double *weight, double *nint,
                             double *dndlnu_max, double *device_ms);

/* FP64 FMA throughput micro-benchmark on the context's device (TFLOP/s, FMA = 2 flop); the roofline
 * denominator of this path (MEASURED_PEAKS.json has no FP64 entry). */
int grmonty_b200_fp64_peak(grmonty_b200_ctx *ctx, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* GRMONTY_B200_H */
