/*
 * capi.cpp -- extern "C" face of the host library (libgrmonty_b200_host.so) so that Python (tests, bench.py)
 * can drive the same HARMModel object the CLI uses.  One opaque handle per model.
 */
#include <cstring>
#include <stdexcept>
#include <string>

#include "harm_model.hpp"

using harm::HARMModel;

static thread_local std::string g_err;
#define GUARD(body)                 \
    try {                           \
        body;                       \
        return 0;                   \
    } catch (const std::exception &e) { \
        g_err = e.what();           \
        return -1;                  \
    }

extern "C" {

const char *gmh_last_error() { return g_err.c_str(); }

void *gmh_create(int photon_n, double mass_unit, int verbosity) {
    harm::set_verbosity(verbosity);
    return new HARMModel(photon_n, mass_unit);
}
void gmh_destroy(void *h) { delete static_cast<HARMModel *>(h); }

int gmh_read_file(void *h, const char *path) { GUARD(static_cast<HARMModel *>(h)->read_file(path)) }
/* mode 0: off, 1: load/store `<dump>.b200cache` (dir: NULL or "" = next to the dump) */
int gmh_set_dump_cache(void *h, int mode, const char *dir) {
    auto *m = static_cast<HARMModel *>(h);
    m->dump_cache = mode != 0;
    m->dump_cache_dir = dir ? dir : "";
    return 0;
}
int gmh_set_device_tables(void *h, int on) {
    static_cast<HARMModel *>(h)->options.device_tables = on != 0;
    return 0;
}
int gmh_set_hotcross_cache(void *h, const char *path) {
    static_cast<HARMModel *>(h)->hotcross_cache = path ? path : "";
    return 0;
}
int gmh_hotcross_from_cache(void *h) { return static_cast<HARMModel *>(h)->hotcross_from_cache() ? 1 : 0; }
int gmh_read_from_cache(void *h) { return static_cast<HARMModel *>(h)->read_from_cache() ? 1 : 0; }
int gmh_report_spectrum_binary(void *h, const char *path) {
    GUARD(static_cast<HARMModel *>(h)->report_spectrum_binary(path))
}
int gmh_init(void *h, int threads) {
    auto *m = static_cast<HARMModel *>(h);
    m->init_threads = threads;
    GUARD(m->init())
}
/* which: 0 geometry 1 hotcross 2 emiss 3 weight 4 nint (host); 5, 6: device builders */
int gmh_init_stage(void *h, int which, int threads) {
    auto *m = static_cast<HARMModel *>(h);
    m->init_threads = threads;
    GUARD(switch (which) {
        case 0: m->init_geometry(); break;
        case 1: m->init_hotcross_table(); break;
        case 2: m->init_emiss_tables(); break;
        case 3: m->init_weight_table(); break;
        case 4: m->init_nint_table(); break;
        case 5: m->init_tables_on_device(false); break; /* geometry + weight + nint on the GPU */
        case 6: m->init_tables_on_device(true); break;  /* ... and the hot cross-section table */
        default: throw std::runtime_error("bad init stage");
    })
}
int gmh_set_options(void *h, uint64_t seed, int rank, int world, int device, int threads_per_block,
                    int blocks_per_sm, int64_t queue_capacity, int64_t gen0, int64_t gen_cap, int64_t gen_budget,
                    int64_t gen_fine_from, int64_t gen_fine_div, void *nccl_comm,
                    const char *cuda_library) {
    auto &o = static_cast<HARMModel *>(h)->options;
    o.seed = seed;
    o.rank = rank;
    o.world = world;
    o.device = device;
    o.threads_per_block = threads_per_block;
    o.blocks_per_sm = blocks_per_sm;
    o.queue_capacity = queue_capacity;
    o.gen0 = gen0;
    o.gen_cap = gen_cap;
    o.gen_budget = gen_budget;
    o.gen_fine_from = gen_fine_from;
    o.gen_fine_div = gen_fine_div;
    o.nccl_comm = nccl_comm;
    o.cuda_library = cuda_library ? cuda_library : "";
    return 0;
}
int gmh_set_gpus(void *h, int gpus) {
    static_cast<HARMModel *>(h)->options.gpus = gpus;
    return 0;
}
int gmh_set_external_reduce(void *h, int on) {
    static_cast<HARMModel *>(h)->options.external_reduce = on != 0;
    return 0;
}
int gmh_run_simulation(void *h) { GUARD(static_cast<HARMModel *>(h)->run_simulation()) }
int gmh_report_spectrum(void *h, const char *path) { GUARD(static_cast<HARMModel *>(h)->report_spectrum(path)) }

/* out[13]: n0 n1 x_start1 x_start2 dx1 dx2 dx3 x_stop1 x_stop2 a h_slope r_0 gamma */
void gmh_get_header(void *h, double *out) {
    const harm::Header &H = *static_cast<HARMModel *>(h)->get_header();
    const double v[13] = {(double)H.n[0], (double)H.n[1], H.x_start[1], H.x_start[2], H.dx[1], H.dx[2], H.dx[3],
                          H.x_stop[1],   H.x_stop[2],    H.a,          H.h_slope,    H.r_0,   H.gamma};
    std::memcpy(out, v, sizeof(v));
}
/* all 26 header fields in file order (for the loader test) */
void gmh_get_header_raw(void *h, double *out) {
    const harm::Header &H = *static_cast<HARMModel *>(h)->get_header();
    const double v[26] = {H.t,        (double)H.n[0],     (double)H.n[1],    H.x_start[1],        H.x_start[2],
                          H.dx[1],    H.dx[2],            H.t_final,         (double)H.n_step,    H.a,
                          H.gamma,    H.courant,          H.dt_dump,         H.dt_log,            H.dt_img,
                          (double)H.dt_rdump, (double)H.cnt_dump, (double)H.cnt_img, (double)H.cnt_rdump, H.dt,
                          (double)H.lim, (double)H.failed, H.r_in,           H.r_out,             H.h_slope,
                          H.r_0};
    std::memcpy(out, v, sizeof(v));
}
/* out[8]: mass_unit l_unit t_unit rho_unit u_unit b_unit theta_e_unit n_e_unit */
void gmh_get_units(void *h, double *out) {
    const harm::Units &u = static_cast<HARMModel *>(h)->units();
    const double v[8] = {u.mass_unit, u.l_unit, u.t_unit, u.rho_unit, u.u_unit, u.b_unit, u.theta_e_unit, u.n_e_unit};
    std::memcpy(out, v, sizeof(v));
}
/* out[3]: bias_norm max_tau_scatt0 photon_n */
void gmh_get_scalars(void *h, double *out) {
    auto *m = static_cast<HARMModel *>(h);
    out[0] = m->bias_norm();
    out[1] = m->max_tau_scatt0();
    out[2] = m->photon_n();
}
/* which: 0..7 primitives (k_rho u u_1 u_2 u_3 b_1 b_2 b_3), 8 geom_det */
int gmh_get_grid(void *h, int which, double *out) {
    auto *m = static_cast<HARMModel *>(h);
    const harm::Data &d = *m->get_data();
    const std::vector<double> *a[9] = {&d.k_rho, &d.u, &d.u_1, &d.u_2, &d.u_3, &d.b_1, &d.b_2, &d.b_3, &m->geom_det()};
    if (which < 0 || which > 8)
        return -1;
    std::memcpy(out, a[which]->data(), a[which]->size() * sizeof(double));
    return 0;
}
/* which: 0 hotcross[221*81] 1 f[201] 2 k2[201] 3 weight[201] 4 nint[20001] 5 dndlnu_max[20001] */
int gmh_get_table(void *h, int which, double *out) {
    auto *m = static_cast<HARMModel *>(h);
    switch (which) {
    case 0: std::memcpy(out, m->hotcross_table().data(), m->hotcross_table().size() * sizeof(double)); break;
    case 1: std::memcpy(out, m->f_table().data(), sizeof(double) * 201); break;
    case 2: std::memcpy(out, m->k2_table().data(), sizeof(double) * 201); break;
    case 3: std::memcpy(out, m->weight_table().data(), sizeof(double) * 201); break;
    case 4: std::memcpy(out, m->nint_table().data(), m->nint_table().size() * sizeof(double)); break;
    case 5: std::memcpy(out, m->dndlnu_max_table().data(), m->dndlnu_max_table().size() * sizeof(double)); break;
    default: return -1;
    }
    return 0;
}
void gmh_get_spectrum(void *h, double *out) {
    auto &s = static_cast<HARMModel *>(h)->spectrum();
    std::memcpy(out, s.data(), s.size() * sizeof(double));
}
void gmh_set_spectrum(void *h, const double *in) {
    auto &s = static_cast<HARMModel *>(h)->spectrum();
    std::memcpy(s.data(), in, s.size() * sizeof(double));
}
/* out[16]: created scattered recorded max_tau seconds kernel_ms transport_ms tracked steps attempts interactions
 *          scatter_events generations launches luminosity max_tau_reported */
void gmh_get_stats(void *h, double *out) {
    auto *m = static_cast<HARMModel *>(h);
    const harm::RunStats &s = m->stats();
    const double v[16] = {(double)s.created, (double)s.scattered, (double)s.recorded, s.max_tau_scatt, s.seconds,
                          s.kernel_ms, s.transport_ms, (double)s.n_tracked, (double)s.n_steps,
                          (double)s.n_push_attempts, (double)s.n_interactions, (double)s.n_scatter_events,
                          (double)s.n_generations, (double)s.n_kernel_launches, m->luminosity(),
                          m->max_tau_scatt_reported()};
    std::memcpy(out, v, sizeof(v));
}

} /* extern "C" */
