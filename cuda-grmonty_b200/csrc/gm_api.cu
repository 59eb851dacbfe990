/*
 * gm_api.cu -- implementation of the C ABI declared in include/grmonty_b200.h.
 *
 * Replaces the seam cuda_super_photon::{alloc_memory, track_super_photons, free_memory}
 * (reference cuda_grmonty/super_photon.cuh:29-61, super_photon.cu:447-1037) with a context object,
 * on-device photon generation and one persistent kernel per generation.  No CPU fallback exists: every
 * entry point needs a CUDA device and fails with GRMONTY_B200_ECUDA otherwise.
 *
 * Two generation schedulers live here: run_batch / grmonty_b200_run_range (one launch per generation, statistics of
 * all earlier generations; the default) and run_range_pipelined (one persistent launch per window of generations,
 * generation clock on the device, csrc/gm_pipeline.cuh; grmonty_b200_config::gen_overlap = 1 / 3).  Three kernels:
 * transport_kernel (fused per-lane loop, the default), wavefront_kernel (csrc/gm_wavefront.cuh, config kernel = 2),
 * pipeline_kernel.  DESIGN.md section 3 has the measurements behind the defaults.
 */
#include <cuda_profiler_api.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h> /* types and enum values only: the functions are resolved with dlopen, see nccl_api() */

#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/grmonty_b200.h"
#include <cub/device/device_radix_sort.cuh>

#include "gm_kernels.cuh"
#include "gm_pipeline.cuh"
#include "gm_tables.cuh"
#include "gm_wavefront.cuh"

using namespace gm;

/* Device memory of a destroyed context, kept for the next context on the same device.  Every device buffer of a
 * context is carved out of ONE cudaMalloc block; cudaFree costs 10 - 100 ms per call on the B200 boxes and 0.2 - 1.4 s
 * for the photon pool (measured, profiles/r1_e2e_split.txt) -- more than the whole transport of 16 M superphotons --
 * so destroy parks the block here (one per device) and create reuses it when it is large enough.
 * grmonty_b200_trim_cache() really frees it. */
struct DeviceArena {
    int device = -1;
    char *base = nullptr;
    size_t bytes = 0;
};
static std::mutex g_arena_mutex;
static std::vector<DeviceArena> g_arena_cache;

struct WfVariant;

struct grmonty_b200_ctx {
    grmonty_b200_config cfg;
    GmParams P;
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    /* device buffers */
    double *d_grid = nullptr, *d_det = nullptr, *d_hotcross = nullptr, *d_f = nullptr, *d_k2 = nullptr;
    double *d_weight = nullptr, *d_nint = nullptr, *d_dnmax = nullptr;
    ZoneData *d_zones = nullptr;
    long long *d_num = nullptr, *d_prefix = nullptr;
    double *d_nz = nullptr;
    PhotonPool pool{};
    SlotQueue ready{}, scatter{};
    SlotQueue carry{};
    PhotonPool stage{}; /* staging pool: suspended photons between two batches */
    unsigned long long n_carry = 0; /* records in `stage` waiting for the next batch */
    unsigned long long h_qc[8] = {0}; /* staging of the queue counters written at each batch start */
    TransportArgs h_args{};           /* staging of the device-global argument block (see run_batch) */
    /* issue order within a generation: primaries of long-lived zones first (see run_batch) */
    unsigned long long *d_zone_cost = nullptr; /* [2][n0] */
    std::vector<unsigned long long> h_zone_cost;
    unsigned char *d_bin_rank = nullptr, *d_keys_in = nullptr, *d_keys_out = nullptr;
    long long *d_vals_in = nullptr, *d_vals_out = nullptr;
    void *d_sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    long long sort_cap = 0;
    int order_mode = 1; /* 0: positions in sequence, 1: sorted by measured zone lifetime */
    unsigned long long h_counters[3] = {0, 0, 0}, h_maxtau_bits = 0; /* accumulators as of the last batch end */
    bool h_bias_valid = false;
    unsigned long long host_tracked = 0; /* primaries started (the device counts scattered children) */
    cudaEvent_t ev2 = nullptr, ev3 = nullptr;
    DeviceArena arena;     /* the one device allocation all buffers below live in */
    size_t arena_used = 0;
    int budget = 384;               /* attempts a lineage may make per generation */
    unsigned long long *d_qctr = nullptr; /* n_alloc, finished, ready head/tail, scatter head/tail, carry head/tail */
    unsigned long long used_ready = 0, used_scatter = 0, used_carry = 0; /* entries to clear before the next batch */
    Accumulators A{};
    double *d_spectrum = nullptr;
    unsigned long long *d_counters = nullptr, *d_maxtau = nullptr, *d_work = nullptr;
    unsigned int *d_error = nullptr;
    TransportArgs *d_args = nullptr; /* device-global copy of the kernel arguments (cold stages) */
    /* host state */
    std::vector<long long> prefix; /* [nzones+1] */
    long long total = 0;
    long long perm_mult = 1; /* Weyl multiplier of the processing order */
    unsigned int gen_tag = 0;
    int threads = 128, blocks_per_sm = 0, grid_blocks = 0;
    int want_bps = 1; /* blocks per SM of the compiled variant in use */
    int kernel = 0;                 /* 0: wavefront, 1: fused loop of round 1 */
    /* pipelined generations (gm_pipeline.cuh): device-side generation clock, one launch per window of generations */
    int overlap = 1;
    int lag_mode = 1; /* 1: generations at the size cap start one generation early, 3: all generations do */
    GenCtl *d_ctl = nullptr;
    GenDesc *d_desc = nullptr; /* own allocation, grown on demand (one entry per generation of a run) */
    size_t desc_cap = 0;
    unsigned int *d_qent = nullptr; /* runnable queues R0 R1 and limbo queues L0 L1, qcap entries each */
    unsigned int qcap = 0;
    bool q_dirty = false; /* a launch did not end cleanly: clear the rings before the next one */
    unsigned long long *h_pin = nullptr;         /* pinned staging: control line + error word + n_alloc read-back */
    const WfVariant *wf = nullptr;  /* the wavefront variant in use */
    double *d_snap = nullptr;       /* wavefront pre-step snapshots [13][snap_stride] */
    unsigned int snap_stride = 0;
    int wf_thr_interact = 192, wf_thr_service = 32;
    long long gen0 = 32, gen_cap = 1 << 20, gen_fine_from = 16384, gen_fine_div = 6, gen_ramp = 8, gen_budget_spread = 0;
    grmonty_b200_stats stats{};
    grmonty_b200_progress_fn progress = nullptr;
    void *progress_user = nullptr;
    std::string err;
};

static thread_local std::string g_create_err;

static int fail(grmonty_b200_ctx *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c)
        c->err = buf;
    else
        g_create_err = buf;
    return code;
}

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return fail(ctx, GRMONTY_B200_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),      \
                        __FILE__, __LINE__);                                                                  \
    } while (0)

/* bump allocation inside the context's arena (256-byte aligned); sizes are summed by arena_bytes_needed() first */
static void *arena_take(grmonty_b200_ctx *ctx, size_t bytes) {
    const size_t off = (ctx->arena_used + 255) & ~(size_t)255;
    if (off + bytes > ctx->arena.bytes)
        return nullptr;
    ctx->arena_used = off + bytes;
    return ctx->arena.base + off;
}
template <typename T> static cudaError_t arena_alloc(grmonty_b200_ctx *ctx, T **dst, size_t n) {
    *dst = static_cast<T *>(arena_take(ctx, n * sizeof(T)));
    return *dst ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <typename T> static cudaError_t upload(grmonty_b200_ctx *ctx, T **dst, const T *src, size_t n) {
    cudaError_t e = arena_alloc(ctx, dst, n);
    if (e != cudaSuccess)
        return e;
    return cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
}

/* dynamic shared memory of the transport kernel: 13 snapshot rows + 6 pending-record rows per thread */
static size_t transport_smem_bytes(int threads) { return (size_t)kTransportSmemRows * threads * sizeof(double); }

/* ---- kernel dispatch over the compiled (block, min-blocks) variants ------------------------------------- */
typedef void (*TransportFn)(const TransportArgs);
struct Variant {
    int block, min_blocks;
    TransportFn fn;      /* one launch per generation (round 1) */
    TransportFn pipe_fn; /* overlapping generations (gm_pipeline.cuh) */
};
#define GM_V(B, M) {B, M, transport_kernel<B, M>, pipeline_kernel<B, M>}
static const Variant kVariants[] = {GM_V(128, 2), GM_V(128, 3), GM_V(256, 1), GM_V(64, 4), GM_V(32, 8),
                                    GM_V(384, 1), GM_V(512, 1)};
#undef GM_V
static const Variant *find_variant(int block, int min_blocks) {
    for (const Variant &v : kVariants)
        if (v.block == block && v.min_blocks == min_blocks)
            return &v;
    return nullptr;
}
/* wavefront kernel (gm_wavefront.cuh): threads per block x photon slots per thread, one block per SM */
struct WfVariant {
    int block, slots, min_blocks;
    TransportFn fn;
    size_t smem;
};
#define GM_WF(B, R, M) {B, R, M, wavefront_kernel<B, R, M>, wavefront_smem_bytes<B, R>()}
/* the third number is the __launch_bounds__ minimum of blocks per SM (register cap); blocks actually resident follow
 * from registers and the shared-memory slots (228 bytes per slot) */
static const WfVariant kWfVariants[] = {GM_WF(384, 2, 1), GM_WF(256, 2, 1), GM_WF(256, 3, 1), GM_WF(512, 1, 1),
                                        GM_WF(384, 1, 1), GM_WF(128, 2, 2), GM_WF(128, 3, 2), GM_WF(128, 2, 3),
                                        GM_WF(192, 2, 2)};
#undef GM_WF
static const WfVariant *find_wf_variant(int block, int slots, int min_blocks) {
    for (const WfVariant &v : kWfVariants)
        if (v.block == block && v.slots == slots && (min_blocks <= 0 || v.min_blocks == min_blocks))
            return &v;
    return nullptr;
}

extern "C" {

const char *grmonty_b200_last_error(grmonty_b200_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

/* scalar part of the device parameter block (pointers are filled in by the caller) */
static void fill_params(GmParams &P, const grmonty_b200_config *cfg) {
    memset(&P, 0, sizeof(P));
    P.n0 = cfg->n0;
    P.n1 = cfg->n1;
    P.x_start1 = cfg->x_start1;
    P.x_start2 = cfg->x_start2;
    P.dx1 = cfg->dx1;
    P.dx2 = cfg->dx2;
    P.x_stop1 = cfg->x_stop1;
    P.x_stop2 = cfg->x_stop2;
    P.a = cfg->a;
    P.h_slope = cfg->h_slope;
    P.r_0 = cfg->r_0;
    P.b_unit = cfg->b_unit;
    P.theta_e_unit = cfg->theta_e_unit;
    P.n_e_unit = cfg->n_e_unit;
    P.photon_n = cfg->photon_n;
    P.bias_norm = cfg->bias_norm;
    /* reference harm_model.cpp:73, :228-229 */
    P.d_tau_k = 2.0 * kPi * cfg->l_unit / (kME * kCL * kCL / kHBAR);
    P.x1_min = std::log(1.0 + std::sqrt(1.0 - cfg->a * cfg->a));
    P.x1_max = std::log(kRMax);
    P.seed_lo = (uint32_t)cfg->seed;
    P.seed_hi = (uint32_t)(cfg->seed >> 32);
    P.l_nu_min = std::log(kNuMin);
    P.n_l_n = std::log(kNuMax) - std::log(kNuMin);
    P.d_l_nu = (std::log(kNuMax) - std::log(kNuMin)) / kNESamp;
    P.l_b_min = std::log(kBthsqMin);
    P.d_l_b = std::log(kBthsqMax / kBthsqMin) / kNint;
    P.hc_l_min_w = std::log10(kHcMinW);
    P.hc_l_min_t = std::log10(kHcMinT);
    P.hc_d_l_w = std::log10(kHcMaxW / kHcMinW) / kHcNW;
    P.hc_d_l_t = std::log10(kHcMaxT / kHcMinT) / kHcNT;
    P.jnu_l_min_k = std::log(kJnuMinK);
    P.jnu_d_l_k = std::log(kJnuMaxK / kJnuMinK) / kNESamp;
    P.jnu_l_min_t = std::log(kThetaEMin);
    P.jnu_d_l_t = std::log(kJnuMaxT / kThetaEMin) / kNESamp;
    P.spec_l_e_0 = std::log(1.0e-12);
    P.nz_max = cfg->photon_n * std::log(kNuMax / kNuMin);
    P.inv_dx1 = 1.0 / P.dx1;
    P.inv_dx2 = 1.0 / P.dx2;
    P.inv_b_unit = 1.0 / P.b_unit;
    P.inv_hc_d_l_w = 1.0 / P.hc_d_l_w;
    P.inv_hc_d_l_t = 1.0 / P.hc_d_l_t;
    P.inv_jnu_d_l_t = 1.0 / P.jnu_d_l_t;
}

int grmonty_b200_create(grmonty_b200_ctx **out, const grmonty_b200_config *cfg) {
    grmonty_b200_ctx *ctx = nullptr;
    if (!out || !cfg)
        return fail(nullptr, GRMONTY_B200_EINVAL, "null argument");
    *out = nullptr;
    if (cfg->abi_version != GRMONTY_B200_ABI_VERSION || cfg->struct_size != sizeof(grmonty_b200_config))
        return fail(nullptr, GRMONTY_B200_EINVAL, "ABI mismatch: version %u size %u (library: %u, %zu)",
                    cfg->abi_version, cfg->struct_size, GRMONTY_B200_ABI_VERSION, sizeof(grmonty_b200_config));
    if (cfg->n0 < 2 || cfg->n1 < 2 || cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world)
        return fail(nullptr, GRMONTY_B200_EINVAL, "bad grid or sharding: n0=%d n1=%d rank=%d world=%d", cfg->n0,
                    cfg->n1, cfg->rank, cfg->world);
    const void *ptrs[] = {cfg->k_rho, cfg->u,  cfg->u_1, cfg->u_2,    cfg->u_3,  cfg->b_1,       cfg->b_2, cfg->b_3,
                          cfg->geom_det, cfg->hotcross, cfg->f, cfg->k2, cfg->weight, cfg->nint, cfg->dndlnu_max};
    for (const void *p : ptrs)
        if (!p)
            return fail(nullptr, GRMONTY_B200_EINVAL, "null input array");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, GRMONTY_B200_ECUDA, "no CUDA device: %s (there is no CPU fallback)",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev)
        return fail(nullptr, GRMONTY_B200_EINVAL, "device %d out of range (%d devices)", cfg->device, ndev);

    ctx = new grmonty_b200_ctx();
    ctx->cfg = *cfg;
    ctx->device = cfg->device;
    int rc = [&]() -> int {
        CK(cudaSetDevice(ctx->device));
        /* one attribute, not cudaGetDeviceProperties: the full query costs up to 0.2 s per call on this driver */
        CK(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, ctx->device));
        const bool trace_create = getenv("GRMONTY_B200_TRACE") != nullptr;
        auto t_last = std::chrono::steady_clock::now();
        auto mark = [&](const char *what) {
            if (!trace_create)
                return;
            const auto now = std::chrono::steady_clock::now();
            fprintf(stderr, "[grmonty_b200] create: %-22s %8.3f ms\n", what,
                    std::chrono::duration<double, std::milli>(now - t_last).count());
            t_last = now;
        };
        CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        CK(cudaEventCreate(&ctx->ev0));
        CK(cudaEventCreate(&ctx->ev1));
        CK(cudaEventCreate(&ctx->ev2));
        CK(cudaEventCreate(&ctx->ev3));

        /* ---- parameter block ---- */
        GmParams &P = ctx->P;
        fill_params(P, cfg);

        mark("streams and events");
        /* ---- one device arena for everything (reused from the cache when possible) ---- */
        const size_t nz = (size_t)cfg->n0 * cfg->n1;
        /* kernel / scheduler choice first: the pool is sized for it.  config: kernel 0 = default (the fused loop, the
         * faster of the two on B200), 1 = fused, 2 = wavefront; gen_overlap 0 = default = 2 = one launch per
         * generation, 1 = pipelined (capped generations start early), 3 = pipelined, every generation starts early */
        if (cfg->kernel < 0 || cfg->kernel > 2)
            return fail(ctx, GRMONTY_B200_EINVAL, "kernel must be 0 (default), 1 (fused) or 2 (wavefront)");
        if (cfg->gen_overlap < 0 || cfg->gen_overlap > 3)
            return fail(ctx, GRMONTY_B200_EINVAL, "gen_overlap must be 0 (default), 1 (on), 2 (off) or 3 (lag 1 everywhere)");
        ctx->lag_mode = cfg->gen_overlap == 3 ? 3 : 1;
        ctx->kernel = cfg->kernel == 2 ? 0 : 1;
        if (const char *e = getenv("GRMONTY_B200_KERNEL")) /* A/B switch for tools/: "fused" or "wavefront" */
            ctx->kernel = strcmp(e, "wavefront") == 0 ? 0 : 1;
        ctx->overlap = (cfg->gen_overlap == 1 || cfg->gen_overlap == 3) ? 1 : 0;
        if (const char *e = getenv("GRMONTY_B200_OVERLAP")) /* A/B switch for tools/ */
            ctx->overlap = atoi(e) != 0;
        if (ctx->kernel == 0) {
            if (cfg->gen_overlap == 1 || cfg->gen_overlap == 3)
                return fail(ctx, GRMONTY_B200_EINVAL, "gen_overlap = 1 / 3 needs the fused kernel");
            ctx->overlap = 0;
        }
        /* pool records: a launch of the pipelined scheduler holds a window of generations (a third of the pool in
         * primaries), the round-1 scheduler one generation */
        unsigned long long cap = cfg->queue_capacity > 0 ? (unsigned long long)cfg->queue_capacity
                                                         : (ctx->overlap ? 32ull << 20 : 1ull << 22);
        if (cap > 0x3fffffffull)
            cap = 0x3fffffffull; /* slots are addressed with 32 bits */
        const unsigned long long stage_cap = std::max<unsigned long long>(1024, ctx->overlap ? cap / 4 : cap / 2);
        /* rings R0 R1 L0 L1 + scatter of the pipelined kernel: a power of two >= the pool (live entries never exceed
         * the records) and well above the number of lanes (tickets outstanding) */
        ctx->qcap = 0u;
        if (ctx->overlap) {
            unsigned long long q = 1ull << 17;
            while (q < cap)
                q <<= 1;
            ctx->qcap = (unsigned int)q;
        }
        /* (the per-generation queues serve the test exports' preloaded batches in overlap mode: a quarter will do) */
        ctx->ready.capacity = (unsigned int)std::min<unsigned long long>(ctx->overlap ? cap / 2 : 2 * cap, 0x7fffffffull);
        ctx->scatter.capacity = (unsigned int)cap;
        ctx->carry.capacity = (unsigned int)stage_cap;
        {
            const size_t per_slot = (size_t)P_NFIELDS * sizeof(double) + 2 * sizeof(uint4) + 3 * sizeof(int);
            size_t need = nz * (8 + 1 + 1) * sizeof(double) + nz * sizeof(ZoneData) + (2 * nz + 1) * sizeof(long long) +
                          (size_t)(GRMONTY_B200_HOTCROSS_N + 3 * GRMONTY_B200_TABLE_N + 2 * GRMONTY_B200_NINT_N) *
                              sizeof(double) +
                          per_slot * (cap + stage_cap) +
                          ((size_t)ctx->ready.capacity + ctx->scatter.capacity + ctx->carry.capacity +
                           (size_t)5 * ctx->qcap) * sizeof(unsigned int) + sizeof(GenCtl) +
                          (size_t)kNThBins * kNEBins * kSpecFields * sizeof(double) + sizeof(TransportArgs) + 4096;
            /* issue-order sort buffers for one batch (keys 1 B, values 8 B, double-buffered) + cub temporary storage */
            ctx->sort_cap = (long long)std::max<unsigned long long>(1024, cap / 4);
            cub::DeviceRadixSort::SortPairs(nullptr, ctx->sort_tmp_bytes, (const unsigned char *)nullptr,
                                            (unsigned char *)nullptr, (const long long *)nullptr, (long long *)nullptr,
                                            (int)ctx->sort_cap, 0, 8, ctx->stream);
            need += (size_t)ctx->sort_cap * 2 * (1 + sizeof(long long)) + ctx->sort_tmp_bytes +
                    (size_t)cfg->n0 * (2 * sizeof(unsigned long long) + 1);
            need += (size_t)13 * ctx->sm_count * 1024 * sizeof(double); /* wavefront snapshot rows: <= 1024 slots / SM */
            need += 64 * 256; /* alignment padding of the ~40 sub-allocations */
            {
                std::lock_guard<std::mutex> lock(g_arena_mutex);
                for (size_t i = 0; i < g_arena_cache.size(); ++i)
                    if (g_arena_cache[i].device == ctx->device && g_arena_cache[i].bytes >= need &&
                        g_arena_cache[i].bytes <= need + need / 2) {
                        ctx->arena = g_arena_cache[i];
                        g_arena_cache.erase(g_arena_cache.begin() + i);
                        break;
                    }
            }
            if (!ctx->arena.base) {
                void *b = nullptr;
                cudaError_t me = cudaMalloc(&b, need);
                if (me == cudaErrorMemoryAllocation) {
                    /* a parked block of another size may be what is in the way: free it and try once more */
                    cudaGetLastError();
                    std::lock_guard<std::mutex> lock(g_arena_mutex);
                    for (size_t i = 0; i < g_arena_cache.size();)
                        if (g_arena_cache[i].device == ctx->device) {
                            cudaFree(g_arena_cache[i].base);
                            g_arena_cache.erase(g_arena_cache.begin() + i);
                        } else {
                            ++i;
                        }
                    me = cudaMalloc(&b, need);
                }
                CK(me);
                ctx->arena.device = ctx->device;
                ctx->arena.base = static_cast<char *>(b);
                ctx->arena.bytes = need;
            }
            ctx->arena_used = 0;
        }

        mark("arena");
        /* ---- model upload: primitives interleaved [n0][n1][8] ---- */
        {
            std::vector<double> inter(nz * 8);
            const double *src[8] = {cfg->k_rho, cfg->u, cfg->u_1, cfg->u_2, cfg->u_3, cfg->b_1, cfg->b_2, cfg->b_3};
            for (size_t z = 0; z < nz; ++z)
                for (int v = 0; v < 8; ++v)
                    inter[z * 8 + v] = src[v][z];
            CK(upload(ctx, &ctx->d_grid, inter.data(), nz * 8));
        }
        CK(upload(ctx, &ctx->d_det, cfg->geom_det, nz));
        CK(upload(ctx, &ctx->d_hotcross, cfg->hotcross, (size_t)GRMONTY_B200_HOTCROSS_N));
        CK(upload(ctx, &ctx->d_f, cfg->f, (size_t)GRMONTY_B200_TABLE_N));
        CK(upload(ctx, &ctx->d_k2, cfg->k2, (size_t)GRMONTY_B200_TABLE_N));
        CK(upload(ctx, &ctx->d_weight, cfg->weight, (size_t)GRMONTY_B200_TABLE_N));
        CK(upload(ctx, &ctx->d_nint, cfg->nint, (size_t)GRMONTY_B200_NINT_N));
        CK(upload(ctx, &ctx->d_dnmax, cfg->dndlnu_max, (size_t)GRMONTY_B200_NINT_N));
        P.grid = ctx->d_grid;
        P.geom_det = ctx->d_det;
        P.hotcross = ctx->d_hotcross;
        P.f = ctx->d_f;
        P.k2 = ctx->d_k2;
        P.weight = ctx->d_weight;
        P.nint = ctx->d_nint;
        P.dndlnu_max = ctx->d_dnmax;

        mark("uploads");
        /* keep the fluid grid L2-resident (access-policy window; best effort) */
        {
            size_t bytes = nz * 8 * sizeof(double);
            int max_persist = 0, max_window = 0;
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
            cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
            /* GRMONTY_B200_L2_WINDOW=0 switches the window off (A/B evidence for the 1024^2 grid, profiles/) */
            const char *l2_env = getenv("GRMONTY_B200_L2_WINDOW");
            if (max_persist > 0 && max_window > 0 && !(l2_env && atoi(l2_env) == 0)) {
                /* the persisting-L2 carve-out is a device-wide limit (and setting it synchronises the device):
                 * once per process and device, and only ever grown */
                static std::mutex limit_mutex;
                static size_t limit_set[64] = {0};
                {
                    std::lock_guard<std::mutex> lock(limit_mutex);
                    const size_t want = std::min(bytes, (size_t)max_persist);
                    if (ctx->device < 64 && limit_set[ctx->device] < want) {
                        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
                        limit_set[ctx->device] = want;
                    }
                }
                cudaStreamAttrValue attr;
                memset(&attr, 0, sizeof(attr));
                attr.accessPolicyWindow.base_ptr = ctx->d_grid;
                attr.accessPolicyWindow.num_bytes = std::min(bytes, (size_t)max_window);
                attr.accessPolicyWindow.hitRatio =
                    (float)std::min(1.0, (double)max_persist / (double)attr.accessPolicyWindow.num_bytes);
                attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
                cudaGetLastError(); /* best effort */
            }
        }

        mark("L2 policy");
        /* ---- per-zone emission data and the zone -> primary index prefix ---- */
        CK(arena_alloc(ctx, &ctx->d_zones, nz));
        CK(arena_alloc(ctx, &ctx->d_num, nz));
        CK(arena_alloc(ctx, &ctx->d_nz, nz));
        CK(arena_alloc(ctx, &ctx->d_prefix, nz + 1));
        zone_kernel<<<(unsigned)((nz + 127) / 128), 128, 0, ctx->stream>>>(P, ctx->d_zones, ctx->d_nz, ctx->d_num);
        CK(cudaGetLastError());
        std::vector<long long> num(nz);
        CK(cudaMemcpyAsync(num.data(), ctx->d_num, nz * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->prefix.assign(nz + 1, 0);
        for (size_t z = 0; z < nz; ++z)
            ctx->prefix[z + 1] = ctx->prefix[z] + num[z];
        ctx->total = ctx->prefix[nz];
        {
            /* multiplier ~ total / golden ratio, coprime to total (same rule as the oracle's orc_perm_multiplier) */
            long long mult = 1;
            if (ctx->total >= 3) {
                mult = std::max<long long>(1, (long long)((double)ctx->total * 0.6180339887498949));
                auto gcd = [](long long a, long long b) {
                    while (b) {
                        const long long t = a % b;
                        a = b;
                        b = t;
                    }
                    return a;
                };
                while (gcd(mult, ctx->total) != 1)
                    ++mult;
            }
            ctx->perm_mult = mult;
        }
        /* on the context's own (non-blocking) stream: a NULL-stream copy would not be ordered before the kernels
         * that read the table */
        CK(cudaMemcpyAsync(ctx->d_prefix, ctx->prefix.data(), (nz + 1) * sizeof(long long), cudaMemcpyHostToDevice,
                           ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));

        mark("zone kernel + prefix");
        /* ---- photon pool and stage queues ---- */
        auto alloc_pool = [&](PhotonPool &pl, unsigned long long c) -> cudaError_t {
            pl.capacity = (unsigned int)c;
            cudaError_t e;
            if ((e = arena_alloc(ctx, &pl.f, (size_t)P_NFIELDS * c)) != cudaSuccess) return e;
            if ((e = arena_alloc(ctx, &pl.rng, c)) != cudaSuccess) return e;
            if ((e = arena_alloc(ctx, &pl.crng, c)) != cudaSuccess) return e;
            if ((e = arena_alloc(ctx, &pl.n_scatt, c)) != cudaSuccess) return e;
            if ((e = arena_alloc(ctx, &pl.n_step, c)) != cudaSuccess) return e;
            return arena_alloc(ctx, &pl.gclock, c);
        };
        CK(alloc_pool(ctx->pool, cap));
        CK(alloc_pool(ctx->stage, stage_cap));
        CK(arena_alloc(ctx, &ctx->ready.entries, (size_t)ctx->ready.capacity));
        CK(arena_alloc(ctx, &ctx->scatter.entries, (size_t)ctx->scatter.capacity));
        CK(arena_alloc(ctx, &ctx->carry.entries, (size_t)ctx->carry.capacity));
        CK(cudaMemsetAsync(ctx->ready.entries, 0, (size_t)ctx->ready.capacity * sizeof(unsigned int), ctx->stream));
        CK(cudaMemsetAsync(ctx->scatter.entries, 0, (size_t)ctx->scatter.capacity * sizeof(unsigned int), ctx->stream));
        CK(cudaMemsetAsync(ctx->carry.entries, 0, (size_t)ctx->carry.capacity * sizeof(unsigned int), ctx->stream));
        CK(arena_alloc(ctx, &ctx->d_zone_cost, (size_t)2 * cfg->n0));
        CK(arena_alloc(ctx, &ctx->d_bin_rank, (size_t)cfg->n0));
        CK(arena_alloc(ctx, &ctx->d_keys_in, (size_t)ctx->sort_cap));
        CK(arena_alloc(ctx, &ctx->d_keys_out, (size_t)ctx->sort_cap));
        CK(arena_alloc(ctx, &ctx->d_vals_in, (size_t)ctx->sort_cap));
        CK(arena_alloc(ctx, &ctx->d_vals_out, (size_t)ctx->sort_cap));
        {
            char *tmp = nullptr;
            CK(arena_alloc(ctx, &tmp, ctx->sort_tmp_bytes + 256));
            ctx->d_sort_tmp = tmp;
        }
        ctx->h_zone_cost.assign((size_t)2 * cfg->n0, 0ull);
        if (const char *e = getenv("GRMONTY_B200_ORDER"))
            ctx->order_mode = atoi(e);
        if (ctx->overlap) {
            CK(arena_alloc(ctx, &ctx->d_qent, (size_t)5 * ctx->qcap));
            CK(cudaMemsetAsync(ctx->d_qent, 0, (size_t)5 * ctx->qcap * sizeof(unsigned int), ctx->stream));
            CK(arena_alloc(ctx, &ctx->d_ctl, 1));
            CK(cudaMemsetAsync(ctx->d_ctl, 0, sizeof(GenCtl), ctx->stream));
            CK(cudaMallocHost(&ctx->h_pin, (CW_WORDS + 8) * sizeof(unsigned long long)));
        }
        CK(arena_alloc(ctx, &ctx->d_qctr, 8));
        CK(cudaMemsetAsync(ctx->d_qctr, 0, 8 * sizeof(unsigned long long), ctx->stream));
        ctx->pool.n_alloc = ctx->d_qctr;
        ctx->pool.finished = ctx->d_qctr + 1;
        ctx->ready.head = ctx->d_qctr + 2;
        ctx->ready.tail = ctx->d_qctr + 3;
        ctx->scatter.head = ctx->d_qctr + 4;
        ctx->scatter.tail = ctx->d_qctr + 5;
        ctx->carry.head = ctx->d_qctr + 6;
        ctx->carry.tail = ctx->d_qctr + 7;
        if (cfg->gen_budget > 0)
            ctx->budget = (int)std::min<long long>(cfg->gen_budget, INT_MAX);

        /* ---- accumulators ---- */
        const size_t nspec = (size_t)kNThBins * kNEBins * kSpecFields;
        CK(arena_alloc(ctx, &ctx->d_spectrum, nspec));
        CK(arena_alloc(ctx, &ctx->d_counters, 3));
        CK(arena_alloc(ctx, &ctx->d_maxtau, 1));
        CK(arena_alloc(ctx, &ctx->d_work, 24));
        CK(arena_alloc(ctx, &ctx->d_error, 1));
        CK(arena_alloc(ctx, &ctx->d_args, 1));
        ctx->A.spectrum = ctx->d_spectrum;
        ctx->A.counters = ctx->d_counters;
        ctx->A.max_tau_bits = ctx->d_maxtau;
        ctx->A.work = ctx->d_work;
        ctx->A.error = ctx->d_error;

        mark("pool and queues");
        /* ---- launch geometry ---- */
        if (ctx->kernel == 0) {
            ctx->threads = cfg->threads_per_block > 0 ? cfg->threads_per_block : 384;
            const int slots = cfg->slots_per_thread > 0 ? cfg->slots_per_thread : 2;
            ctx->wf = find_wf_variant(ctx->threads, slots, cfg->blocks_per_sm);
            if (!ctx->wf)
                return fail(ctx, GRMONTY_B200_EINVAL,
                            "no compiled wavefront variant for %d threads x %d slots (x %d blocks/SM)", ctx->threads,
                            slots, cfg->blocks_per_sm);
            CK(cudaFuncSetAttribute((const void *)ctx->wf->fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)ctx->wf->smem));
            int occ = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)ctx->wf->fn, ctx->threads,
                                                             ctx->wf->smem));
            if (occ < 1)
                return fail(ctx, GRMONTY_B200_ECUDA, "wavefront kernel does not fit on an SM");
            ctx->blocks_per_sm = occ;
            ctx->grid_blocks = occ * ctx->sm_count;
            ctx->snap_stride = (unsigned int)(ctx->grid_blocks * ctx->threads * slots);
            CK(arena_alloc(ctx, &ctx->d_snap, (size_t)13 * ctx->snap_stride));
            if (cfg->wf_thr_interact > 0)
                ctx->wf_thr_interact = cfg->wf_thr_interact;
            if (cfg->wf_thr_service > 0)
                ctx->wf_thr_service = cfg->wf_thr_service;
            if (const char *e = getenv("GRMONTY_B200_WF_THR")) /* "interact,service" in 1/256 (tools/ sweeps) */
                sscanf(e, "%d,%d", &ctx->wf_thr_interact, &ctx->wf_thr_service);
        } else {
            /* default geometry 32 x 8: 8 warps/SM like 256 x 1 (255 registers: two warps per scheduler is what the
             * register file holds), each warp a block of its own -- no warp ever waits at a block barrier for another
             * one's rare slow path.  Measured per configs[1] run: 578 (32 x 8), 586 (64 x 4), 589 (128 x 2), 593 ms
             * (256 x 1); configs[0]: 113 against 128 ms; fewer warps lose in proportion (7: 619, 6: 668 ms), a ninth
             * would need 168 registers (three warps on one scheduler).  profiles/r2_ab_microopts.txt */
            ctx->threads = cfg->threads_per_block > 0 ? cfg->threads_per_block : 32;
            int want_bps = cfg->blocks_per_sm > 0 ? cfg->blocks_per_sm : (cfg->threads_per_block > 0 ? 1 : 8);
            ctx->want_bps = want_bps;
            const Variant *v = find_variant(ctx->threads, want_bps);
            if (!v)
                return fail(ctx, GRMONTY_B200_EINVAL, "no compiled kernel variant for %d threads x %d blocks/SM",
                            ctx->threads, want_bps);
            const size_t smem = transport_smem_bytes(ctx->threads);
            CK(cudaFuncSetAttribute((const void *)v->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaFuncSetAttribute((const void *)v->pipe_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int occ = 0, occ_pipe = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)v->fn, ctx->threads, smem));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_pipe, (const void *)v->pipe_fn, ctx->threads, smem));
            if (ctx->overlap)
                occ = std::min(occ, occ_pipe); /* all blocks of the persistent launches must be co-resident */
            if (occ < 1)
                return fail(ctx, GRMONTY_B200_ECUDA, "transport kernel does not fit on an SM");
            ctx->blocks_per_sm = std::min(occ, want_bps);
            ctx->grid_blocks = ctx->blocks_per_sm * ctx->sm_count;
        }
        if (cfg->gen0 > 0)
            ctx->gen0 = cfg->gen0;
        if (cfg->gen_cap > 0)
            ctx->gen_cap = cfg->gen_cap;
        if (cfg->gen_fine_from > 0)
            ctx->gen_fine_from = cfg->gen_fine_from;
        if (cfg->gen_fine_div != 0)
            ctx->gen_fine_div = cfg->gen_fine_div;
        if (cfg->gen_ramp > 1)
            ctx->gen_ramp = cfg->gen_ramp;
        if (cfg->gen_budget_spread != 0)
            ctx->gen_budget_spread = cfg->gen_budget_spread; /* negative: off */
        mark("launch geometry");
        return GRMONTY_B200_OK;
    }();
    if (rc != GRMONTY_B200_OK) {
        g_create_err = ctx->err;
        grmonty_b200_destroy(ctx);
        return rc;
    }
    rc = grmonty_b200_reset(ctx);
    if (rc != GRMONTY_B200_OK) {
        g_create_err = ctx->err;
        grmonty_b200_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return GRMONTY_B200_OK;
}

int grmonty_b200_reset(grmonty_b200_ctx *ctx) {
    if (!ctx)
        return GRMONTY_B200_EINVAL;
    CK(cudaSetDevice(ctx->device));
    const size_t nspec = (size_t)kNThBins * kNEBins * kSpecFields;
    CK(cudaMemsetAsync(ctx->d_spectrum, 0, nspec * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_counters, 0, 3 * sizeof(unsigned long long), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_work, 0, 24 * sizeof(unsigned long long), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_error, 0, sizeof(unsigned int), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_zone_cost, 0, (size_t)2 * ctx->P.n0 * sizeof(unsigned long long), ctx->stream));
    std::fill(ctx->h_zone_cost.begin(), ctx->h_zone_cost.end(), 0ull);
    unsigned long long bits;
    const double mt = ctx->cfg.max_tau_scatt0;
    memcpy(&bits, &mt, sizeof(bits));
    CK(cudaMemcpyAsync(ctx->d_maxtau, &bits, sizeof(bits), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    ctx->h_bias_valid = false;
    ctx->host_tracked = 0;
    /* suspended photons of an aborted run must not enter the next one; used_ready / used_scatter / used_carry keep
     * their values so that begin_batch still clears the stale queue entries */
    ctx->n_carry = 0;
    return GRMONTY_B200_OK;
}

int grmonty_b200_total_primaries(grmonty_b200_ctx *ctx, int64_t *total) {
    if (!ctx || !total)
        return GRMONTY_B200_EINVAL;
    *total = ctx->total;
    return GRMONTY_B200_OK;
}

/* size of the generation starting at run position g_start (same rule as the oracle's orc_generation_size;
 * documented at grmonty_b200_config::gen0) */
static long long generation_size(long long g_start, long long gen0, long long cap, long long fine_from,
                                 long long fine_div, long long ramp) {
    long long s;
    if (g_start < gen0)
        s = gen0;
    else if (g_start < fine_from || fine_div <= 1)
        s = g_start * (std::max<long long>(ramp, 2) - 1);
    else
        s = g_start / fine_div;
    s = std::max(s, gen0);
    return std::min(s, cap);
}

static void fill_args(grmonty_b200_ctx *ctx, const GmBiasStats &bias, const DebugOut &dbg, int budget,
                      TransportArgs &args) {
    args.P = ctx->P;
    args.bias = bias;
    args.pool = ctx->pool;
    args.ready = ctx->ready;
    args.scatter = ctx->scatter;
    args.carry = ctx->carry;
    args.budget = budget;
    args.A = ctx->A;
    args.D = dbg;
    args.zone_cost = ctx->d_zone_cost;
    args.snap = ctx->d_snap;
    args.snap_stride = ctx->snap_stride;
    args.wf_thr_interact = ctx->wf_thr_interact;
    args.wf_thr_service = ctx->wf_thr_service;
    args.ctl = ctx->d_ctl;
    args.qent = ctx->d_qent;
    args.qcap = ctx->qcap;
    args.self = ctx->d_args;
}

/* reset pool and queues for a batch whose first `count` records / ready entries are (or will be) filled */
static int begin_batch(grmonty_b200_ctx *ctx, long long count) {
    if (ctx->used_ready)
        CK(cudaMemsetAsync(ctx->ready.entries, 0, ctx->used_ready * sizeof(unsigned int), ctx->stream));
    if (ctx->used_scatter)
        CK(cudaMemsetAsync(ctx->scatter.entries, 0, ctx->used_scatter * sizeof(unsigned int), ctx->stream));
    if (ctx->used_carry)
        CK(cudaMemsetAsync(ctx->carry.entries, 0, ctx->used_carry * sizeof(unsigned int), ctx->stream));
    ctx->used_ready = ctx->used_scatter = ctx->used_carry = 0;
    /* the source buffer lives in the context: it is rewritten only after the batch's final synchronisation */
    unsigned long long *qc = ctx->h_qc;
    qc[0] = qc[3] = (unsigned long long)count;
    qc[1] = qc[2] = qc[4] = qc[5] = qc[6] = qc[7] = 0ull;
    CK(cudaMemcpyAsync(ctx->d_qctr, qc, 8 * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
    return GRMONTY_B200_OK;
}

/* Run one batch with frozen bias statistics: `count` new primaries at positions first, first+stride, ...
 * plus the photons carried over from the previous batch (ctx->n_carry records in the staging pool).
 * preloaded: records 0..count-1 and their ready entries were already written (test export).
 * budget: attempts a lineage may make in this batch before it is suspended and carried over. */
static int run_batch(grmonty_b200_ctx *ctx, long long first, long long stride, long long count,
                     const GmBiasStats &bias, const DebugOut &dbg, bool preloaded, int budget) {
    const Variant *v = ctx->kernel == 0 ? nullptr
                                        : find_variant(ctx->threads, ctx->want_bps);
    /* the device-global copy of the arguments is written on the batch's own stream, ahead of the kernels that read it
     * through A.self; the source lives in the context and is rewritten only after the batch's final synchronisation */
    TransportArgs &args = ctx->h_args;
    fill_args(ctx, bias, dbg, budget, args);
    CK(cudaMemcpyAsync(ctx->d_args, &args, sizeof(args), cudaMemcpyHostToDevice, ctx->stream));
    float ms = 0.f;
    /* extra attempts for lineages that start early in the batch (see birth_kernel); none in the final drain */
    const long long spread = (budget != INT_MAX && ctx->gen_budget_spread > 0) ? ctx->gen_budget_spread : 0;
    const long long n_carry = preloaded ? 0 : (long long)ctx->n_carry;
    const long long n_start = count + n_carry;
    if (!preloaded) {
        int rc = begin_batch(ctx, n_start);
        if (rc)
            return rc;
        CK(cudaEventRecord(ctx->ev2, ctx->stream));
        /* Issue order.  Lanes take the batch's pool slots in sequence, and a launch ends when its last lineage has
         * finished or used up its budget, so the primaries that live longest should start first (longest-processing-
         * time-first): lineages born at r = 3..10 M make ~400 attempts on the bench dump, those born inside the
         * photon orbit or beyond 15 M fewer than 50.  The transport kernel counts steps of finished primaries per
         * radial bin of their birth zone; once 4096 primaries have finished, each batch is sorted by the measured
         * mean lifetime of its zones' bins (cub radix sort on an 8-bit rank).  Results do not depend on the order. */
        const long long *order = nullptr;
        if (count > 0 && ctx->order_mode == 1 && count <= ctx->sort_cap && ctx->P.n0 <= 4096) {
            const int n0 = ctx->P.n0;
            unsigned long long finished = 0;
            for (int i = 0; i < n0; ++i)
                finished += ctx->h_zone_cost[n0 + i];
            if (finished >= 4096) {
                std::vector<std::pair<double, int>> by_cost(n0);
                for (int i = 0; i < n0; ++i) {
                    const unsigned long long c = ctx->h_zone_cost[n0 + i];
                    by_cost[i] = {c ? (double)ctx->h_zone_cost[i] / (double)c : 0.0, i};
                }
                std::stable_sort(by_cost.begin(), by_cost.end(),
                                 [](const std::pair<double, int> &a, const std::pair<double, int> &b) { return a.first > b.first; });
                std::vector<unsigned char> rank(n0);
                for (int r = 0; r < n0; ++r) /* at most 256 distinct ranks: bins of similar lifetime share one */
                    rank[by_cost[r].second] = (unsigned char)((long long)r * 256 / n0);
                CK(cudaMemcpyAsync(ctx->d_bin_rank, rank.data(), n0, cudaMemcpyHostToDevice, ctx->stream));
                order_key_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(
                    ctx->d_prefix, n0, ctx->P.n1, ctx->d_bin_rank, first, stride, count, ctx->perm_mult, ctx->total,
                    ctx->d_keys_in, ctx->d_vals_in);
                CK(cudaGetLastError());
                size_t tmp_bytes = ctx->sort_tmp_bytes;
                CK(cub::DeviceRadixSort::SortPairs(ctx->d_sort_tmp, tmp_bytes, ctx->d_keys_in, ctx->d_keys_out,
                                                   ctx->d_vals_in, ctx->d_vals_out, (int)count, 0, 8, ctx->stream));
                CK(cudaStreamSynchronize(ctx->stream)); /* `rank` is a stack buffer */
                order = ctx->d_vals_out;
                ctx->stats.n_kernel_launches += 2;
            }
        }
        if (count > 0) {
            const int bb = 128;
            const long long want = (count + bb - 1) / bb;
            const int nb = (int)std::min<long long>(want, (long long)ctx->sm_count * 16);
            birth_kernel<<<nb, bb, 0, ctx->stream>>>(args, ctx->d_zones, ctx->d_prefix, first, stride, count,
                                                     ctx->perm_mult, ctx->total, spread, order);
            CK(cudaGetLastError());
            ctx->stats.n_kernel_launches += 1;
        }
        if (n_carry > 0) {
            carry_copy_kernel<<<(unsigned)((n_carry + 127) / 128), 128, 0, ctx->stream>>>(
                ctx->pool, (unsigned int)count, ctx->stage, 0u, nullptr, (unsigned int)n_carry, ctx->ready.entries,
                spread > 0 ? -(int)(count / spread) : 0);
            CK(cudaGetLastError());
            ctx->stats.n_kernel_launches += 1;
            ctx->n_carry = 0;
        }
        CK(cudaEventRecord(ctx->ev3, ctx->stream)); /* read after the batch's one synchronisation below */
    }
    const size_t smem = ctx->kernel == 0 ? ctx->wf->smem : transport_smem_bytes(ctx->threads);
    /* do not launch far more threads than there are photons to start with (tiny generations / test batches) */
    long long blocks = std::min<long long>(ctx->grid_blocks,
                                           std::max<long long>(1, (n_start * 2 + ctx->threads - 1) / ctx->threads));
    blocks = std::max<long long>(blocks, std::min<long long>(ctx->grid_blocks, ctx->sm_count));
    /* profiling hook: GRMONTY_B200_PROFILE_MIN_COUNT=n brackets the first transport launch that starts with at
     * least n photons with cudaProfilerStart/Stop (use with `ncu --profile-from-start off`) */
    static long long prof_min = -2;
    if (prof_min == -2) {
        const char *e = getenv("GRMONTY_B200_PROFILE_MIN_COUNT");
        prof_min = e ? atoll(e) : -1;
    }
    /* GRMONTY_B200_PROFILE_DRAIN=k: bracket the k-th final-drain launch of the process instead */
    static int prof_drain = -1;
    if (prof_drain < 0)
        prof_drain = getenv("GRMONTY_B200_PROFILE_DRAIN") ? atoi(getenv("GRMONTY_B200_PROFILE_DRAIN")) : 0;
    const bool is_drain = budget == INT_MAX && !preloaded;
    const bool prof = (prof_min >= 0 && n_start >= prof_min) || (prof_drain == 1 && is_drain);
    if (is_drain && prof_drain > 1)
        --prof_drain; /* not yet */
    if (prof)
        cudaProfilerStart();
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    (ctx->kernel == 0 ? ctx->wf->fn : v->fn)<<<(unsigned)blocks, ctx->threads, smem, ctx->stream>>>(args);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    /* one host synchronisation per batch: error word, queue counters and the accumulators the next generation's
     * bias statistics are made of come back together */
    unsigned int err = 0;
    unsigned long long qc[8];
    CK(cudaMemcpyAsync(&err, ctx->d_error, sizeof(err), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(qc, ctx->d_qctr, sizeof(qc), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                       ctx->stream));
    CK(cudaMemcpyAsync(&ctx->h_maxtau_bits, ctx->d_maxtau, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                       ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_zone_cost.data(), ctx->d_zone_cost, ctx->h_zone_cost.size() * sizeof(unsigned long long),
                       cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->h_bias_valid = true;
    if (prof) {
        cudaProfilerStop();
        prof_min = -1;
        prof_drain = 0;
    }
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->stats.kernel_ms += ms;
    ctx->stats.transport_ms += ms;
    ctx->stats.n_kernel_launches += 1;
    {
        static int trace = -1; /* GRMONTY_B200_TRACE=1: one line per transport launch on stderr */
        if (trace < 0)
            trace = getenv("GRMONTY_B200_TRACE") ? 1 : 0;
        if (trace) {
            fprintf(stderr, "[grmonty_b200] batch first=%lld count=%lld carried_in=%lld budget=%d blocks=%lld: %.3f ms, "
                            "records=%llu ready=%llu scatter=%llu carried_out=%llu\n",
                    first, count, n_carry, budget, blocks, ms, qc[0], qc[3], qc[5], qc[7]);
            if (ctx->kernel == 0) { /* wavefront phase statistics of this launch (differences of the run totals) */
                static unsigned long long prev[24] = {0};
                unsigned long long w[24];
                CK(cudaMemcpy(w, ctx->d_work, sizeof(w), cudaMemcpyDeviceToHost));
                unsigned long long d[24];
                for (int i = 0; i < 24; ++i)
                    d[i] = w[i] >= prev[i] ? w[i] - prev[i] : w[i];
                memcpy(prev, w, sizeof(w));
                const double lanes = (double)blocks * ctx->threads;
                fprintf(stderr, "    phases/block: push %.0f interact %.0f service %.0f scatter %.0f idle %.0f; lane "
                                "efficiency push %.3f interact %.3f; us/phase push %.2f interact %.2f service %.2f "
                                "scatter %.2f idle %.2f\n",
                        d[8] / (double)blocks, d[9] / (double)blocks, d[10] / (double)blocks, d[11] / (double)blocks,
                        d[12] / (double)blocks, d[13] / std::max(1.0, d[8] * (double)ctx->threads),
                        d[14] / std::max(1.0, d[9] * (double)ctx->threads), d[16] / 1965.0 / std::max(1.0, (double)d[8]),
                        d[17] / 1965.0 / std::max(1.0, (double)d[9]), d[18] / 1965.0 / std::max(1.0, (double)d[10]),
                        d[19] / 1965.0 / std::max(1.0, (double)d[11]), d[20] / 1965.0 / std::max(1.0, (double)d[12]));
                (void)lanes;
            }
        }
    }
    if (!preloaded) {
        CK(cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3));
        ctx->stats.kernel_ms += ms;
    }
    ctx->used_ready = std::min<unsigned long long>(qc[3], ctx->ready.capacity);
    ctx->used_scatter = std::min<unsigned long long>(qc[5], ctx->scatter.capacity);
    ctx->used_carry = std::min<unsigned long long>(qc[7], ctx->carry.capacity);
    ctx->stats.queue_high_water = std::max<uint64_t>(ctx->stats.queue_high_water, qc[0]);
    /* photons tracked = records created: primaries counted here, scattered children on the device */
    ctx->host_tracked += (unsigned long long)count;
    if (err & 1u)
        return fail(ctx, GRMONTY_B200_EQUEUE,
                    "device photon pool/queue overflow: %llu records, %llu ready, %llu scatter, %llu carry entries "
                    "needed (capacity %u); raise queue_capacity or lower gen_cap",
                    qc[0], qc[3], qc[5], qc[7], ctx->pool.capacity);
    if (err & 2u)
        return fail(ctx, GRMONTY_B200_ECUDA, "device photon queue: entry publication timeout");
    /* stash the suspended photons for the next batch */
    if (qc[7] > 0) {
        if (qc[7] > ctx->stage.capacity)
            return fail(ctx, GRMONTY_B200_EQUEUE, "carry-over staging pool overflow: %llu records (capacity %u)", qc[7],
                        ctx->stage.capacity);
        carry_copy_kernel<<<(unsigned)((qc[7] + 127) / 128), 128, 0, ctx->stream>>>(
            ctx->stage, 0u, ctx->pool, 0u, ctx->carry.entries, (unsigned int)qc[7], nullptr, 0);
        CK(cudaGetLastError()); /* stream order keeps the next batch's queue clearing behind this copy */
        ctx->stats.n_kernel_launches += 1;
        ctx->n_carry = qc[7];
    }
    return GRMONTY_B200_OK;
}

static int read_bias_stats(grmonty_b200_ctx *ctx, GmBiasStats *b) {
    unsigned long long c[3], bits;
    if (ctx->h_bias_valid) { /* fetched with the previous batch's results */
        memcpy(c, ctx->h_counters, sizeof(c));
        bits = ctx->h_maxtau_bits;
    } else {
        CK(cudaMemcpyAsync(c, ctx->d_counters, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(&bits, ctx->d_maxtau, sizeof(bits), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    double mt;
    memcpy(&mt, &bits, sizeof(mt));
    *b = make_bias_stats(ctx->P.bias_norm, mt, (double)c[1], (double)c[2]);
    return GRMONTY_B200_OK;
}

/* The pipelined scheduler (gm_pipeline.cuh): the generations of [first, last) in windows of as many generations as
 * the pool holds, ONE persistent launch per window.  The host only prepares a window (birth of its primaries, queue
 * reset, carried-over lineages) and collects it; the generation clock -- statistics, completion, opening of the next
 * generations -- runs on the device and continues across windows. */
static int run_range_pipelined(grmonty_b200_ctx *ctx, long long first, long long last) {
    const long long world = ctx->cfg.world, rank = ctx->cfg.rank;
    const Variant *v = find_variant(ctx->threads, ctx->want_bps);
    /* ---- the generations of this call (same partition as the round-1 scheduler and the oracle) ---- */
    struct Gen {
        long long f0, count, pos_end;
        bool capped;
    };
    std::vector<Gen> gens;
    for (long long g_start = 0; g_start < last;) {
        const long long size = generation_size(g_start / world, ctx->gen0, ctx->gen_cap, ctx->gen_fine_from,
                                               ctx->gen_fine_div, ctx->gen_ramp);
        const long long g_end = g_start + world * size;
        const long long lo = std::max<long long>(g_start, first), hi = std::min<long long>(g_end, last);
        if (lo < hi) {
            const long long f0 = lo + ((rank - lo % world) % world + world) % world;
            gens.push_back({f0, f0 < hi ? (hi - f0 + world - 1) / world : 0, hi, size >= ctx->gen_cap});
        }
        g_start = g_end;
    }
    const int n_real = (int)gens.size();
    if (n_real == 0)
        return GRMONTY_B200_OK;
    const int n_desc = n_real + 1; /* + the final drain generation: no primaries, no attempt budget */
    const long long spread = ctx->gen_budget_spread > 0 ? ctx->gen_budget_spread : 0;
    std::vector<GenDesc> desc((size_t)n_desc);
    for (int g = 0; g < n_real; ++g) {
        desc[g].count = (unsigned long long)gens[g].count;
        desc[g].prim_end = 0ull;
        desc[g].carry_clock0 = spread > 0 ? -(int)(gens[g].count / spread) : 0;
        /* early start (statistics one generation behind): the generations at the size cap, or all (study knob) */
        desc[g].lag = ctx->lag_mode == 3 ? 1 : (gens[g].capped ? 1 : 0);
        desc[g].t_open = desc[g].t_done = 0ull;
    }
    desc[n_real] = GenDesc{0ull, 0ull, kClockForever, 0, 0ull, 0ull};
    if ((size_t)n_desc > ctx->desc_cap) {
        if (ctx->d_desc)
            CK(cudaFree(ctx->d_desc));
        ctx->d_desc = nullptr;
        ctx->desc_cap = std::max<size_t>(1024, (size_t)n_desc * 2);
        CK(cudaMalloc(&ctx->d_desc, ctx->desc_cap * sizeof(GenDesc)));
    }
    /* ---- generation clock: the first two generations use the statistics as they stand (lag 1 from there on) ---- */
    GmBiasStats bias;
    ctx->h_bias_valid = false;
    int rc = read_bias_stats(ctx, &bias);
    if (rc)
        return rc;
    GenCtl h;
    memset(&h, 0, sizeof(h));
    for (int i = 0; i < 4; ++i) {
        h.alloc[i] = i < n_desc ? desc[i].count : 0ull;
        h.bias_den[i] = bias.bias_den;
    }
    h.n_desc = n_desc;
    h.bias_norm = ctx->P.bias_norm;
    h.desc = ctx->d_desc;
    {
        const char *e = getenv("GRMONTY_B200_WATCHDOG_S"); /* a launch that lost a record ends itself after this long */
        h.t_limit = (unsigned long long)((e ? atof(e) : 120.0) * 1e9);
    }
    TransportArgs &args = ctx->h_args;
    const DebugOut nodbg = {nullptr, nullptr, 0};
    fill_args(ctx, bias, nodbg, ctx->budget, args);
    /* the scatter queue of this scheduler: the fifth ring, its counters in the control line */
    args.scatter.entries = ctx->d_qent + (size_t)4 * ctx->qcap;
    args.scatter.capacity = ctx->qcap;
    args.scatter.head = ctx->d_ctl->line + CW_SC_HEAD;
    args.scatter.tail = ctx->d_ctl->line + CW_SC_TAIL;
    args.zone_cost = nullptr;
    CK(cudaMemcpyAsync(ctx->d_args, &args, sizeof(args), cudaMemcpyHostToDevice, ctx->stream));
    static int trace = -1; /* GRMONTY_B200_TRACE=1: one line per launch on stderr */
    if (trace < 0)
        trace = getenv("GRMONTY_B200_TRACE") ? 1 : 0;

    /* Primaries per window.  A window's records are its primaries, the lineages carried in and the scattered children
     * it creates; how many children a primary has depends on the model (0.46 on the bench dump, 3 on the small test
     * model, 9 in its first generations).  The first window assumes 15 per primary, later ones the measured ratio of
     * the run so far with a margin of 1.5. */
    long long w_cap = std::max<long long>(1024, (long long)(ctx->pool.capacity / 16));
    unsigned long long n_carry_in = 0, created = 0, children_so_far = 0;
    const size_t smem = transport_smem_bytes(ctx->threads);
    for (int g0 = 0; g0 < n_desc;) {
        /* ---- the window: generations g0 .. g1 - 1 (at least one), then the drain generation if they are the last ---- */
        long long W = 0;
        int g1 = g0;
        while (g1 < n_real && (g1 == g0 || W + gens[g1].count <= w_cap)) {
            W += gens[g1].count;
            desc[g1].prim_end = (unsigned long long)W;
            ++g1;
        }
        const bool final_window = g1 == n_real;
        const int g_end = final_window ? n_desc : g1;
        if (final_window)
            desc[n_real].prim_end = (unsigned long long)W;
        if ((unsigned long long)W + n_carry_in + (unsigned long long)W / 8 > ctx->pool.capacity)
            return fail(ctx, GRMONTY_B200_EQUEUE,
                        "a generation of %lld primaries (+ %llu carried lineages) does not fit the photon pool "
                        "(capacity %u): raise queue_capacity or lower gen_cap", W, n_carry_in, ctx->pool.capacity);
        CK(cudaMemcpyAsync(ctx->d_desc, desc.data(), (size_t)n_desc * sizeof(GenDesc), cudaMemcpyHostToDevice,
                           ctx->stream));
        /* ---- queues and control line (the rings clean themselves: a consumer writes 0 back) ---- */
        if (ctx->q_dirty)
            CK(cudaMemsetAsync(ctx->d_qent, 0, (size_t)5 * ctx->qcap * sizeof(unsigned int), ctx->stream));
        ctx->q_dirty = true; /* until this launch has ended cleanly */
        memset(h.line, 0, sizeof(h.line));
        const int p0 = g0 & 1;
        h.line[CW_L0_TAIL + kCwLStride * p0] = n_carry_in; /* lineages suspended into g0 by the previous window */
        h.line[CW_L0_LIM + kCwLStride * p0] = n_carry_in;  /* frozen: what follows them will be for g0 + 2 */
        /* g0 is open; g0 + 1 as well if it may start early */
        const bool open1 = g0 + 1 < g_end && desc[g0 + 1].lag == 1;
        h.line[CW_L0_LIM + kCwLStride * (p0 ^ 1)] = open1 ? kGateLive : 0ull;
        h.line[CW_PRIM_LIM] = desc[open1 ? g0 + 1 : g0].prim_end;
        h.line[CW_COMPLETE] = (unsigned long long)g0;
        h.g_end = g_end;
        h.t_start = 0ull;
        h.lock = 0u;
        if (g0 == 0) {
            CK(cudaMemcpyAsync(ctx->d_ctl, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
        } else { /* the ring (alloc / done / statistics / denominators) lives on: only the line and the launch's end */
            CK(cudaMemcpyAsync(ctx->d_ctl->line, h.line, sizeof(h.line), cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(&ctx->d_ctl->lock, &h.lock, 4 * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(&ctx->d_ctl->t_start, &h.t_start, sizeof(h.t_start), cudaMemcpyHostToDevice, ctx->stream));
        }
        unsigned long long *qc = ctx->h_qc;
        qc[0] = (unsigned long long)W + n_carry_in; /* records in use: primaries, then the carried lineages */
        qc[1] = 0ull;
        CK(cudaMemcpyAsync(ctx->d_qctr, qc, 2 * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
        /* ---- birth of the window's primaries, generation by generation (records only: lanes take them by index) ---- */
        CK(cudaEventRecord(ctx->ev2, ctx->stream));
        for (int g = g0; g < g1; ++g) {
            if (gens[g].count <= 0)
                continue;
            const int bb = 128;
            const int nb = (int)std::min<long long>((gens[g].count + bb - 1) / bb, (long long)ctx->sm_count * 16);
            birth_kernel<<<nb, bb, 0, ctx->stream>>>(args, ctx->d_zones, ctx->d_prefix, gens[g].f0, world, gens[g].count,
                                                     ctx->perm_mult, ctx->total, spread, nullptr,
                                                     (unsigned int)(desc[g].prim_end - desc[g].count),
                                                     (g & 3) << kTagShift, 0);
            CK(cudaGetLastError());
            ctx->stats.n_kernel_launches += 1;
        }
        if (n_carry_in > 0) {
            carry_copy_kernel<<<(unsigned)((n_carry_in + 127) / 128), 128, 0, ctx->stream>>>(
                ctx->pool, (unsigned int)W, ctx->stage, 0u, nullptr, (unsigned int)n_carry_in,
                ctx->d_qent + (size_t)(2 + p0) * ctx->qcap, 0, true);
            CK(cudaGetLastError());
            ctx->stats.n_kernel_launches += 1;
        }
        CK(cudaEventRecord(ctx->ev3, ctx->stream));
        /* ---- the launch ---- */
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
        v->pipe_fn<<<(unsigned)ctx->grid_blocks, ctx->threads, smem, ctx->stream>>>(args);
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        unsigned long long *pin = ctx->h_pin; /* control words, then the error word and the records allocated */
        constexpr int kPinErr = CW_WORDS, kPinRec = CW_WORDS + 1;
        pin[kPinErr] = 0ull;
        CK(cudaMemcpyAsync(pin, ctx->d_ctl->line, CW_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(pin + kPinErr, ctx->d_error, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(pin + kPinRec, ctx->d_qctr, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f, ms_birth = 0.f;
        CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        CK(cudaEventElapsedTime(&ms_birth, ctx->ev2, ctx->ev3));
        ctx->stats.kernel_ms += ms + ms_birth;
        ctx->stats.transport_ms += ms;
        ctx->stats.n_kernel_launches += 1;
        ctx->stats.n_generations += (uint64_t)(g_end - g0);
        ctx->stats.queue_high_water = std::max<uint64_t>(ctx->stats.queue_high_water, pin[kPinRec]);
        const unsigned int err = (unsigned int)pin[kPinErr];
        ctx->host_tracked += (unsigned long long)W;
        created += (unsigned long long)W;
        children_so_far += pin[kPinRec] > (unsigned long long)W + n_carry_in ? pin[kPinRec] - (unsigned long long)W - n_carry_in : 0ull;
        {
            const double per_primary = 1.0 + 1.5 * (double)children_so_far / (double)std::max<unsigned long long>(created, 1) + 0.15;
            w_cap = std::max<long long>(1024, (long long)((double)ctx->pool.capacity / per_primary));
        }
        if (trace)
            fprintf(stderr, "[grmonty_b200] window generations %d..%d (%lld primaries, %llu carried in): birth %.3f ms, "
                            "transport %.3f ms, records %llu, runnable %llu + %llu, limbo %llu + %llu, scatter %llu, "
                            "complete %llu, error %u\n",
                    g0, g_end - 1, W, n_carry_in, ms_birth, ms, pin[kPinRec], pin[CW_R0_TAIL], pin[CW_R1_TAIL],
                    pin[CW_L0_TAIL], pin[CW_L1_TAIL], pin[CW_SC_TAIL], pin[CW_COMPLETE], err);
        if (err & 1u)
            return fail(ctx, GRMONTY_B200_EQUEUE,
                        "device photon pool/queue overflow: %llu records, queue tails %llu %llu %llu %llu, scatter %llu "
                        "(capacity %u); raise queue_capacity or lower gen_cap", pin[kPinRec], pin[CW_R0_TAIL],
                        pin[CW_R1_TAIL], pin[CW_L0_TAIL], pin[CW_L1_TAIL], pin[CW_SC_TAIL], ctx->pool.capacity);
        if (err & 2u)
            return fail(ctx, GRMONTY_B200_ECUDA, "device photon queue: entry publication timeout");
        if ((err & 8u) || (long long)pin[CW_COMPLETE] < (long long)g_end) {
            GenCtl dbg;
            cudaMemcpy(&dbg, ctx->d_ctl, sizeof(dbg), cudaMemcpyDeviceToHost);
            return fail(ctx, GRMONTY_B200_ECUDA,
                        "generation pipeline stalled: %llu of %d generations complete (watchdog); ring alloc %llu %llu %llu "
                        "%llu done %llu %llu %llu %llu; R %llu/%llu %llu/%llu L %llu/%llu<%llu %llu/%llu<%llu prim %llu/%llu "
                        "scatter %llu/%llu", pin[CW_COMPLETE], g_end, dbg.alloc[0], dbg.alloc[1], dbg.alloc[2], dbg.alloc[3],
                        dbg.done[0], dbg.done[1], dbg.done[2], dbg.done[3], pin[CW_R0_HEAD], pin[CW_R0_TAIL], pin[CW_R1_HEAD],
                        pin[CW_R1_TAIL], pin[CW_L0_HEAD], pin[CW_L0_TAIL], pin[CW_L0_LIM], pin[CW_L1_HEAD], pin[CW_L1_TAIL],
                        pin[CW_L1_LIM], pin[CW_PRIM_CUR], pin[CW_PRIM_LIM], pin[CW_SC_HEAD], pin[CW_SC_TAIL]);
        }
        /* ---- lineages suspended into the next window's first generation: to the staging pool ---- */
        n_carry_in = 0;
        if (!final_window) {
            const int pe = g_end & 1;
            /* from the frozen gate, not from the head: claims may have run ahead of the gate (tickets that never resolved) */
            const unsigned long long lh = pin[CW_L0_LIM + kCwLStride * pe], lt = pin[CW_L0_TAIL + kCwLStride * pe];
            const unsigned long long n = lt > lh ? lt - lh : 0ull;
            if (n > ctx->stage.capacity)
                return fail(ctx, GRMONTY_B200_EQUEUE, "carry-over staging pool overflow: %llu records (capacity %u)", n,
                            ctx->stage.capacity);
            /* the entries lh .. lt - 1 of the ring, in up to two contiguous pieces; they are cleared afterwards */
            unsigned long long done_n = 0;
            while (done_n < n) {
                const unsigned long long at = (lh + done_n) & (unsigned long long)(ctx->qcap - 1u);
                const unsigned long long piece = std::min<unsigned long long>(n - done_n, ctx->qcap - at);
                unsigned int *list = ctx->d_qent + (size_t)(2 + pe) * ctx->qcap + at;
                carry_copy_kernel<<<(unsigned)((piece + 127) / 128), 128, 0, ctx->stream>>>(
                    ctx->stage, (unsigned int)done_n, ctx->pool, 0u, list, (unsigned int)piece, nullptr, 0, true);
                CK(cudaGetLastError());
                CK(cudaMemsetAsync(list, 0, piece * sizeof(unsigned int), ctx->stream));
                ctx->stats.n_kernel_launches += 1;
                done_n += piece;
            }
            n_carry_in = n;
        }
        ctx->q_dirty = false;
        if (trace > 0) { /* per-generation timeline of the launch (device clock, ms since the first opening) */
            CK(cudaMemcpy(desc.data(), ctx->d_desc, (size_t)n_desc * sizeof(GenDesc), cudaMemcpyDeviceToHost));
            unsigned long long t0 = ~0ull;
            for (int g = g0; g < g_end; ++g)
                if (desc[g].t_done)
                    t0 = std::min(t0, desc[g].t_open ? desc[g].t_open : desc[g].t_done);
            for (int g = g0; g < g_end; ++g)
                fprintf(stderr, "    generation %d: %llu primaries, opened %.3f ms, complete %.3f ms\n", g, desc[g].count,
                        desc[g].t_open ? (desc[g].t_open - t0) * 1e-6 : 0.0, (desc[g].t_done - t0) * 1e-6);
        }
        if (ctx->progress)
            ctx->progress(ctx->progress_user, gens[g1 - 1].pos_end, ctx->total);
        g0 = g_end;
    }
    ctx->h_bias_valid = false;
    /* counters[0] = created (host-side count; the reference counts primaries only, harm_model.cpp:395) */
    add_u64_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_counters, created);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return GRMONTY_B200_OK;
}

int grmonty_b200_run_range(grmonty_b200_ctx *ctx, int64_t first, int64_t last) {
    if (!ctx)
        return GRMONTY_B200_EINVAL;
    CK(cudaSetDevice(ctx->device));
    if (last < 0 || last > ctx->total)
        last = ctx->total;
    if (first < 0)
        first = 0;
    if (ctx->overlap && ctx->kernel == 1)
        return run_range_pipelined(ctx, first, last);
    const long long world = ctx->cfg.world, rank = ctx->cfg.rank;
    const DebugOut nodbg = {nullptr, nullptr, 0};
    /* a batch never fills more than a quarter of the pool with primaries: the rest is room for carried
     * photons and scattered children */
    const long long chunk_cap = std::max<long long>(1024, (long long)(ctx->pool.capacity / 4));
    long long g_start = 0;
    unsigned long long created = 0;
    GmBiasStats bias;
    for (long long g = 0; g_start < last; ++g) {
        /* the schedule runs on rank-local positions (global position / world): every rank sees the generation sizes
         * of a stand-alone run of its own share, whatever the number of GPUs */
        const long long g_end =
            g_start + world * generation_size(g_start / world, ctx->gen0, ctx->gen_cap, ctx->gen_fine_from,
                                              ctx->gen_fine_div, ctx->gen_ramp);
        const long long lo = std::max<long long>(g_start, first), hi = std::min<long long>(g_end, last);
        if (lo < hi) {
            long long f0 = lo + ((rank - lo % world) % world + world) % world; /* first index >= lo, = rank mod world */
            long long count = f0 < hi ? (hi - f0 + world - 1) / world : 0;
            if (count > 0 || ctx->n_carry > 0) {
                int rc = read_bias_stats(ctx, &bias);
                if (rc)
                    return rc;
                do {
                    const long long n = std::min(count, chunk_cap);
                    rc = run_batch(ctx, f0, world, n, bias, nodbg, false, ctx->budget);
                    if (rc) {
                        ctx->n_carry = 0; /* nothing of a failed run is carried into the next one */
                        return rc;
                    }
                    created += (unsigned long long)n;
                    f0 += n * world;
                    count -= n;
                } while (count > 0);
                ++ctx->stats.n_generations;
                if (ctx->progress)
                    ctx->progress(ctx->progress_user, std::min<long long>(g_end, last), ctx->total);
            }
        }
        g_start = g_end;
    }
    /* drain: photons still suspended after the last generation run to completion (no budget) */
    while (ctx->n_carry > 0) {
        int rc = read_bias_stats(ctx, &bias);
        if (rc)
            return rc;
        rc = run_batch(ctx, 0, 1, 0, bias, nodbg, false, INT_MAX);
        if (rc) {
            ctx->n_carry = 0;
            return rc;
        }
        ++ctx->stats.n_generations;
    }
    /* counters[0] = created (host-side count; the reference counts primaries only, harm_model.cpp:395) */
    add_u64_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_counters, created);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return GRMONTY_B200_OK;
}

int grmonty_b200_run(grmonty_b200_ctx *ctx) { return grmonty_b200_run_range(ctx, 0, -1); }

int grmonty_b200_set_progress(grmonty_b200_ctx *ctx, grmonty_b200_progress_fn cb, void *user) {
    if (!ctx)
        return GRMONTY_B200_EINVAL;
    ctx->progress = cb;
    ctx->progress_user = user;
    return GRMONTY_B200_OK;
}

int grmonty_b200_device_accumulators(grmonty_b200_ctx *ctx, void **spectrum, void **counters, void **max_tau) {
    if (!ctx)
        return GRMONTY_B200_EINVAL;
    if (spectrum)
        *spectrum = ctx->d_spectrum;
    if (counters)
        *counters = ctx->d_counters;
    if (max_tau)
        *max_tau = ctx->d_maxtau;
    ctx->h_bias_valid = false; /* the caller may reduce into these buffers */
    return GRMONTY_B200_OK;
}

/* NCCL is resolved at run time so that the library has no link-time NCCL dependency and, inside a process
 * that already loaded NCCL (e.g. PyTorch's bundled copy), uses that very copy.  Types and enum values come from
 * <nccl.h> at build time. */
struct NcclApi {
    decltype(&ncclAllReduce) all_reduce = nullptr;
    decltype(&ncclGroupStart) group_start = nullptr;
    decltype(&ncclGroupEnd) group_end = nullptr;
    decltype(&ncclGetUniqueId) get_unique_id = nullptr;
    decltype(&ncclCommInitRank) comm_init_rank = nullptr;
    decltype(&ncclCommInitAll) comm_init_all = nullptr;
    decltype(&ncclCommDestroy) comm_destroy = nullptr;
    decltype(&ncclGetErrorString) error_string = nullptr;
    bool ok = false;
    std::string why;
};
static const NcclApi &nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        /* a copy already in the process (same soname) is what dlopen returns; GRMONTY_B200_NCCL_LIB overrides */
        std::vector<std::string> names;
        if (const char *e = getenv("GRMONTY_B200_NCCL_LIB"))
            names.push_back(e);
        names.push_back("libnccl.so.2");
        names.push_back("libnccl.so");
        void *h = nullptr;
        for (const std::string &n : names)
            if ((h = dlopen(n.c_str(), RTLD_NOW | RTLD_GLOBAL)))
                break;
        if (!h) {
            api.why = std::string("libnccl.so.2 not loadable: ") + dlerror();
            return;
        }
#define GM_NCCL_SYM(field, name)                                     \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name)); \
    if (!api.field) {                                                \
        api.why = std::string(name) + " not found in libnccl";       \
        return;                                                      \
    }
        GM_NCCL_SYM(all_reduce, "ncclAllReduce")
        GM_NCCL_SYM(group_start, "ncclGroupStart")
        GM_NCCL_SYM(group_end, "ncclGroupEnd")
        GM_NCCL_SYM(get_unique_id, "ncclGetUniqueId")
        GM_NCCL_SYM(comm_init_rank, "ncclCommInitRank")
        GM_NCCL_SYM(comm_init_all, "ncclCommInitAll")
        GM_NCCL_SYM(comm_destroy, "ncclCommDestroy")
        GM_NCCL_SYM(error_string, "ncclGetErrorString")
#undef GM_NCCL_SYM
        api.ok = true;
    });
    return api;
}

int grmonty_b200_nccl_unique_id(void *id) {
    static_assert(sizeof(ncclUniqueId) == GRMONTY_B200_NCCL_ID_BYTES, "ncclUniqueId size");
    if (!id)
        return fail(nullptr, GRMONTY_B200_EINVAL, "null id");
    const NcclApi &N = nccl_api();
    if (!N.ok)
        return fail(nullptr, GRMONTY_B200_ENCCL, "%s", N.why.c_str());
    const ncclResult_t r = N.get_unique_id(static_cast<ncclUniqueId *>(id));
    if (r != ncclSuccess)
        return fail(nullptr, GRMONTY_B200_ENCCL, "ncclGetUniqueId: %s", N.error_string(r));
    return GRMONTY_B200_OK;
}

int grmonty_b200_nccl_comm_init_rank(void **comm, const void *id, int rank, int world, int device) {
    grmonty_b200_ctx *ctx = nullptr; /* for the CK macro */
    if (!comm || !id || world < 1 || rank < 0 || rank >= world)
        return fail(nullptr, GRMONTY_B200_EINVAL, "bad communicator arguments (rank %d of %d)", rank, world);
    const NcclApi &N = nccl_api();
    if (!N.ok)
        return fail(nullptr, GRMONTY_B200_ENCCL, "%s", N.why.c_str());
    CK(cudaSetDevice(device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t c = nullptr;
    const ncclResult_t r = N.comm_init_rank(&c, world, uid, rank);
    if (r != ncclSuccess)
        return fail(nullptr, GRMONTY_B200_ENCCL, "ncclCommInitRank: %s", N.error_string(r));
    *comm = c;
    return GRMONTY_B200_OK;
}

int grmonty_b200_nccl_comm_init_all(void **comms, int n_devices, const int *devices) {
    if (!comms || n_devices < 1)
        return fail(nullptr, GRMONTY_B200_EINVAL, "bad communicator arguments (%d devices)", n_devices);
    const NcclApi &N = nccl_api();
    if (!N.ok)
        return fail(nullptr, GRMONTY_B200_ENCCL, "%s", N.why.c_str());
    std::vector<ncclComm_t> c((size_t)n_devices, nullptr);
    const ncclResult_t r = N.comm_init_all(c.data(), n_devices, devices);
    if (r != ncclSuccess)
        return fail(nullptr, GRMONTY_B200_ENCCL, "ncclCommInitAll: %s", N.error_string(r));
    for (int i = 0; i < n_devices; ++i)
        comms[i] = c[(size_t)i];
    return GRMONTY_B200_OK;
}

int grmonty_b200_nccl_comm_destroy(void *comm) {
    if (!comm)
        return GRMONTY_B200_OK;
    const NcclApi &N = nccl_api();
    if (!N.ok)
        return fail(nullptr, GRMONTY_B200_ENCCL, "%s", N.why.c_str());
    const ncclResult_t r = N.comm_destroy(static_cast<ncclComm_t>(comm));
    return r == ncclSuccess ? GRMONTY_B200_OK : fail(nullptr, GRMONTY_B200_ENCCL, "ncclCommDestroy: %s", N.error_string(r));
}

int grmonty_b200_allreduce(grmonty_b200_ctx *ctx, void *nccl_comm, void *cuda_stream) {
    if (!ctx)
        return GRMONTY_B200_EINVAL;
    if (ctx->cfg.world == 1)
        return GRMONTY_B200_OK;
    if (!nccl_comm)
        return fail(ctx, GRMONTY_B200_EINVAL, "world > 1 needs an ncclComm_t");
    const NcclApi &N = nccl_api();
    if (!N.ok)
        return fail(ctx, GRMONTY_B200_ENCCL, "%s", N.why.c_str());
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    ncclComm_t comm = static_cast<ncclComm_t>(nccl_comm);
    const size_t nspec = (size_t)kNThBins * kNEBins * kSpecFields;
    /* one group: spectrum (sum), the three counters (sum), max_tau_scatt (max; non-negative doubles order like their
     * bit patterns, so a max over uint64 is the max over the doubles) */
    ncclResult_t r[4];
    N.group_start();
    r[0] = N.all_reduce(ctx->d_spectrum, ctx->d_spectrum, nspec, ncclFloat64, ncclSum, comm, s);
    r[1] = N.all_reduce(ctx->d_counters, ctx->d_counters, 3, ncclUint64, ncclSum, comm, s);
    r[2] = N.all_reduce(ctx->d_maxtau, ctx->d_maxtau, 1, ncclUint64, ncclMax, comm, s);
    r[3] = N.group_end();
    for (ncclResult_t e : r)
        if (e != ncclSuccess)
            return fail(ctx, GRMONTY_B200_ENCCL, "ncclAllReduce failed: %s", N.error_string(e));
    CK(cudaStreamSynchronize(s));
    ctx->h_bias_valid = false;
    return GRMONTY_B200_OK;
}

int grmonty_b200_result(grmonty_b200_ctx *ctx, double *spectrum, uint64_t counts[3], double *max_tau_scatt,
                        grmonty_b200_stats *stats) {
    if (!ctx)
        return GRMONTY_B200_EINVAL;
    CK(cudaSetDevice(ctx->device));
    const size_t nspec = (size_t)kNThBins * kNEBins * kSpecFields;
    if (spectrum)
        CK(cudaMemcpy(spectrum, ctx->d_spectrum, nspec * sizeof(double), cudaMemcpyDeviceToHost));
    if (counts) {
        unsigned long long c[3];
        CK(cudaMemcpy(c, ctx->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
        counts[0] = c[0];
        counts[1] = c[1];
        counts[2] = c[2];
    }
    if (max_tau_scatt) {
        unsigned long long bits;
        CK(cudaMemcpy(&bits, ctx->d_maxtau, sizeof(bits), cudaMemcpyDeviceToHost));
        memcpy(max_tau_scatt, &bits, sizeof(double));
    }
    if (stats) {
        unsigned long long w[24];
        CK(cudaMemcpy(w, ctx->d_work, sizeof(w), cudaMemcpyDeviceToHost));
        if (getenv("GRMONTY_B200_TRACE"))
            fprintf(stderr, "[grmonty_b200] wavefront block-phases: push %llu interact %llu service %llu scatter %llu idle "
                            "%llu; lane-phases with work: push %llu interact %llu; Mcycles of thread 0 per phase kind (sum "
                            "over blocks): push %.0f interact %.0f service %.0f scatter %.0f idle %.0f\n",
                    w[8], w[9], w[10], w[11], w[12], w[13], w[14], w[16] * 1e-6, w[17] * 1e-6, w[18] * 1e-6, w[19] * 1e-6,
                    w[20] * 1e-6);
        ctx->stats.n_tracked = w[0] + ctx->host_tracked;
        ctx->stats.n_steps = w[1];
        ctx->stats.n_push_attempts = w[2];
        ctx->stats.n_interactions = w[3];
        ctx->stats.n_scatter_events = w[4];
        ctx->stats.n_live_iterations = w[5];
        ctx->stats.n_slot_iterations = w[6];
        *stats = ctx->stats;
    }
    return GRMONTY_B200_OK;
}

void grmonty_b200_destroy(grmonty_b200_ctx *ctx) {
    if (!ctx)
        return;
    cudaSetDevice(ctx->device);
    if (ctx->stream)
        cudaStreamSynchronize(ctx->stream);
    if (ctx->arena.base) {
        /* park the device block for the next context on this device (see DeviceArena): one block per device */
        std::lock_guard<std::mutex> lock(g_arena_mutex);
        for (size_t i = 0; i < g_arena_cache.size(); ++i)
            if (g_arena_cache[i].device == ctx->arena.device) {
                cudaFree(g_arena_cache[i].base);
                g_arena_cache.erase(g_arena_cache.begin() + i);
                break;
            }
        g_arena_cache.push_back(ctx->arena);
        ctx->arena = DeviceArena{};
    }
    if (ctx->d_desc)
        cudaFree(ctx->d_desc);
    if (ctx->h_pin)
        cudaFreeHost(ctx->h_pin);
    if (ctx->ev0)
        cudaEventDestroy(ctx->ev0);
    if (ctx->ev1)
        cudaEventDestroy(ctx->ev1);
    if (ctx->ev2)
        cudaEventDestroy(ctx->ev2);
    if (ctx->ev3)
        cudaEventDestroy(ctx->ev3);
    if (ctx->stream)
        cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int grmonty_b200_hotcross_table(int device, double *table) {
    grmonty_b200_ctx *ctx = nullptr; /* for the CK macro: errors go to the thread's create-error slot */
    if (!table)
        return fail(nullptr, GRMONTY_B200_EINVAL, "null table");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
        return fail(nullptr, GRMONTY_B200_ECUDA, "no such CUDA device %d (there is no CPU fallback)", device);
    CK(cudaSetDevice(device));
    const int n = GRMONTY_B200_HOTCROSS_N;
    const double l_min_w = std::log10(kHcMinW), l_min_t = std::log10(kHcMinT);
    const double d_l_w = std::log10(kHcMaxW / kHcMinW) / kHcNW, d_l_t = std::log10(kHcMaxT / kHcMinT) / kHcNT;
    std::vector<double> axes(kHcNW + 1 + kHcNT + 1);
    for (int i = 0; i <= kHcNW; ++i)
        axes[i] = std::pow(10.0, l_min_w + i * d_l_w);
    for (int j = 0; j <= kHcNT; ++j)
        axes[kHcNW + 1 + j] = std::pow(10.0, l_min_t + j * d_l_t);
    double *d = nullptr;
    CK(cudaMalloc(&d, (n + axes.size()) * sizeof(double)));
    cudaError_t e = cudaMemcpy(d + n, axes.data(), axes.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        hotcross_table_kernel<<<(n + 63) / 64, 64>>>(d, d + n, d + n + kHcNW + 1);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess)
        e = cudaMemcpy(table, d, n * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess)
        return fail(nullptr, GRMONTY_B200_ECUDA, "hotcross_table_kernel: %s", cudaGetErrorString(e));
    return GRMONTY_B200_OK;
}

int grmonty_b200_init_tables(const grmonty_b200_config *cfg, double *geom_det, double *weight, double *nint,
                             double *dndlnu_max, double *device_ms) {
    grmonty_b200_ctx *ctx = nullptr; /* for the CK macro: errors go to the thread's create-error slot */
    if (!cfg)
        return fail(nullptr, GRMONTY_B200_EINVAL, "null config");
    if (cfg->abi_version != GRMONTY_B200_ABI_VERSION || cfg->struct_size != sizeof(grmonty_b200_config))
        return fail(nullptr, GRMONTY_B200_EINVAL, "ABI mismatch: version %u size %u (library: %u, %zu)",
                    cfg->abi_version, cfg->struct_size, GRMONTY_B200_ABI_VERSION, sizeof(grmonty_b200_config));
    if (cfg->n0 < 2 || cfg->n1 < 2 || !(cfg->photon_n > 0))
        return fail(nullptr, GRMONTY_B200_EINVAL, "bad grid %d x %d or photon_n", cfg->n0, cfg->n1);
    const double *src[8] = {cfg->k_rho, cfg->u, cfg->u_1, cfg->u_2, cfg->u_3, cfg->b_1, cfg->b_2, cfg->b_3};
    for (const double *g : src)
        if (!g)
            return fail(nullptr, GRMONTY_B200_EINVAL, "null input array");
    if (!cfg->f || !cfg->k2)
        return fail(nullptr, GRMONTY_B200_EINVAL, "init_tables needs the F(K) and K2 tables");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || cfg->device < 0 || cfg->device >= ndev)
        return fail(nullptr, GRMONTY_B200_ECUDA, "no such CUDA device %d (there is no CPU fallback)", cfg->device);
    CK(cudaSetDevice(cfg->device));

    GmParams P;
    fill_params(P, cfg);
    const size_t nz = (size_t)cfg->n0 * cfg->n1;
    const size_t n_tab = GRMONTY_B200_TABLE_N, n_nint = GRMONTY_B200_NINT_N;
    /* one allocation: grid[nz][8] det fac te bb [nz] | f k2 weight [201] | nint dnmax [20001] */
    const size_t total = nz * 12 + 3 * n_tab + 2 * n_nint;
    double *d = nullptr;
    CK(cudaMalloc(&d, total * sizeof(double)));
    double *d_grid = d, *d_det = d + nz * 8, *d_fac = d_det + nz, *d_te = d_fac + nz, *d_bb = d_te + nz;
    double *d_f = d_bb + nz, *d_k2 = d_f + n_tab, *d_w = d_k2 + n_tab, *d_nint = d_w + n_tab, *d_dn = d_nint + n_nint;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t e = cudaSuccess;
    auto step = [&](cudaError_t r) {
        if (e == cudaSuccess)
            e = r;
    };
    {
        std::vector<double> inter(nz * 8);
        for (size_t z = 0; z < nz; ++z)
            for (int v = 0; v < 8; ++v)
                inter[z * 8 + v] = src[v][z];
        step(cudaMemcpy(d_grid, inter.data(), nz * 8 * sizeof(double), cudaMemcpyHostToDevice));
    }
    step(cudaMemcpy(d_f, cfg->f, n_tab * sizeof(double), cudaMemcpyHostToDevice));
    step(cudaMemcpy(d_k2, cfg->k2, n_tab * sizeof(double), cudaMemcpyHostToDevice));
    P.grid = d_grid;
    P.f = d_f;
    P.k2 = d_k2;
    step(cudaEventCreate(&e0));
    step(cudaEventCreate(&e1));
    if (e == cudaSuccess) {
        /* reference :283-285 (zone volume in cm^3) and :329-333 */
        const double s_fac = cfg->dx1 * cfg->dx2 * cfg->dx3 * cfg->l_unit * cfg->l_unit * cfg->l_unit;
        const double n_fac = s_fac * 1.41421356237309504880 * kEE * kEE * kEE / (27.0 * kME * kCL * kCL) * (1.0 / kHPL);
        cudaEventRecord(e0);
        zone_table_kernel<<<(unsigned)((nz + 127) / 128), 128>>>(P, s_fac, d_det, d_fac, d_te, d_bb);
        weight_table_kernel<512><<<(unsigned)n_tab, 512>>>(P, d_fac, d_te, d_bb, d_w);
        nint_table_kernel<<<(unsigned)((n_nint + 127) / 128), 128>>>(P, d_w, n_fac, d_nint, d_dn);
        cudaEventRecord(e1);
        step(cudaGetLastError());
        step(cudaEventSynchronize(e1));
        float ms = 0.f;
        if (e == cudaSuccess)
            cudaEventElapsedTime(&ms, e0, e1);
        if (device_ms)
            *device_ms = ms;
    }
    if (geom_det)
        step(cudaMemcpy(geom_det, d_det, nz * sizeof(double), cudaMemcpyDeviceToHost));
    if (weight)
        step(cudaMemcpy(weight, d_w, n_tab * sizeof(double), cudaMemcpyDeviceToHost));
    if (nint)
        step(cudaMemcpy(nint, d_nint, n_nint * sizeof(double), cudaMemcpyDeviceToHost));
    if (dndlnu_max)
        step(cudaMemcpy(dndlnu_max, d_dn, n_nint * sizeof(double), cudaMemcpyDeviceToHost));
    if (e0)
        cudaEventDestroy(e0);
    if (e1)
        cudaEventDestroy(e1);
    cudaFree(d);
    if (e != cudaSuccess)
        return fail(nullptr, GRMONTY_B200_ECUDA, "init_tables: %s", cudaGetErrorString(e));
    return GRMONTY_B200_OK;
}

void grmonty_b200_trim_cache(void) {
    std::lock_guard<std::mutex> lock(g_arena_mutex);
    for (DeviceArena &a : g_arena_cache) {
        cudaSetDevice(a.device);
        cudaFree(a.base);
    }
    g_arena_cache.clear();
}

int grmonty_b200_fp64_peak(grmonty_b200_ctx *ctx, double *tflops) {
    if (!ctx || !tflops)
        return GRMONTY_B200_EINVAL;
    CK(cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 1 << 16;
    double *d = nullptr;
    CK(cudaMalloc(&d, (size_t)blocks * threads * sizeof(double)));
    fp64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d, 1024); /* warm-up */
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
        fp64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d, iters);
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        CK(cudaEventSynchronize(ctx->ev1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    CK(cudaFree(d));
    *tflops = best;
    return GRMONTY_B200_OK;
}

} /* extern "C" */

#ifdef GRMONTY_B200_TEST_EXPORTS
#include "gm_test_exports.inc"
#endif
