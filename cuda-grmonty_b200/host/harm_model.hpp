/*
 * harm_model.hpp -- C++20 host side of the B200 path: the reference's HARMModel surface.
 *
 * Mirrors the public interface of the reference's harm::HARMModel (cuda_grmonty/harm_model.hpp:30-85):
 *     HARMModel(photon_n, mass_unit); read_file(path); init(); run_simulation(); report_spectrum(path);
 *     get_header(); get_data();
 * with the same call order as the reference main (main.cpp:43-53).  read_file keeps the HARM dump text format
 * (harm_model.cpp:81-232, fixture in tests/harm_model_test.cpp:224-262), report_spectrum keeps the spectrum
 * file format byte for byte (harm_model.cpp:416-471).  run_simulation replaces both the CPU loop
 * (harm_model.cpp:366-404) and the reference GPU seam (:345-361) with calls into the C ABI
 * (include/grmonty_b200.h); there is no CPU transport path here.
 *
 * Everything is written from scratch for this project (plain std::vector storage instead of NDArray, a
 * from_chars dump parser, threaded table builders); it links against nothing but libstdc++ and dlopen()s the
 * CUDA library only when run_simulation is called.
 */
#pragma once
#include <array>
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

struct grmonty_b200_config;

namespace harm {

/* reference harm_data.hpp:19-44 */
struct Header {
    double t = 0;
    int n[2] = {0, 0};
    double x_start[4] = {0, 0, 0, 0};
    double x_stop[4] = {0, 0, 0, 0};
    double dx[4] = {0, 0, 0, 0};
    double t_final = 0;
    int n_step = 0;
    double a = 0, gamma = 0, courant = 0, dt_dump = 0, dt_log = 0, dt_img = 0;
    int dt_rdump = 0, cnt_dump = 0, cnt_img = 0, cnt_rdump = 0;
    double dt = 0;
    int lim = 0, failed = 0;
    double r_in = 0, r_out = 0, h_slope = 0, r_0 = 0;
};

/* reference harm_data.hpp:49-58; row-major [n0][n1] */
struct Data {
    std::vector<double> k_rho, u, u_1, u_2, u_3, b_1, b_2, b_3;
};

/* reference harm_data.hpp:63-72 */
struct Units {
    double mass_unit = 0, l_unit = 0, t_unit = 0, rho_unit = 0, u_unit = 0, b_unit = 0, theta_e_unit = 0,
           n_e_unit = 0;
};

constexpr int kNThBins = 6, kNEBins = 200, kSpecFields = 13;
constexpr int kNESamp = 200, kNint = 20000, kHcNW = 220, kHcNT = 80;

/* options of the B200 run that have no counterpart in the reference */
struct RunOptions {
    uint64_t seed = 123; /* reference consts::rng_seed, main.cpp:49 */
    int rank = 0, world = 1, device = 0;
    /* > 1: run_simulation() shards the run over devices 0..gpus-1 of this box, one host thread per GPU, and
     * all-reduces the spectrum at the end (the command line's --gpus N); needs world == 1 */
    int gpus = 1;
    int threads_per_block = 0, blocks_per_sm = 0;
    int kernel = 0;      /* grmonty_b200_config::kernel: 0 default (fused loop), 1 fused, 2 wavefront */
    int gen_overlap = 0; /* grmonty_b200_config::gen_overlap: 0 default (overlapping generations), 1 on, 2 off */
    int64_t queue_capacity = 0, gen0 = 0, gen_cap = 0, gen_budget = 0, gen_fine_from = 0, gen_fine_div = 0;
    bool device_tables = false; /* init(): geometry / weight / nint / hot cross-section tables on the GPU */
    void *nccl_comm = nullptr; /* ncclComm_t for world > 1: run_simulation() ends with grmonty_b200_allreduce */
    /* world > 1 without a communicator is an error (the spectrum would silently be one rank's share) unless the
     * caller states that it reduces the per-rank spectra itself */
    bool external_reduce = false;
    /* log a "Rate ... ph/s" line about once a second while the run is in progress (reference harm_model.cpp:397-403);
     * only runs longer than a second ever print one */
    bool progress = true;
    std::string cuda_library;  /* path of libgrmonty_b200.so; empty: $GRMONTY_B200_LIB or next to this library */
};

struct RunStats {
    uint64_t created = 0, scattered = 0, recorded = 0;
    double max_tau_scatt = 0, seconds = 0, kernel_ms = 0, transport_ms = 0;
    uint64_t n_tracked = 0, n_steps = 0, n_push_attempts = 0, n_interactions = 0, n_scatter_events = 0,
             n_generations = 0, n_kernel_launches = 0;
};

class HARMModel {
public:
    explicit HARMModel(int photon_n, double mass_unit);
    HARMModel(const HARMModel &) = delete;
    HARMModel &operator=(const HARMModel &) = delete;

    void read_file(std::string filepath); /* throws std::runtime_error if the file cannot be read */
    void init();                          /* geometry, hot cross-section, emissivity, weight and nint tables */
    void run_simulation();                /* on the GPU, through the C ABI; throws on failure */
    void report_spectrum(std::string filepath);
    /* the whole [6][200][13] accumulator + run counters, binary (the text file drops 7 of the 13 fields) */
    void report_spectrum_binary(std::string filepath);

    const Header *get_header() const { return &header_; }
    const Data *get_data() const { return &data_; }

    /* ---- beyond the reference surface (used by the CLI, the tests and bench.py) ---- */
    RunOptions options;
    int init_threads = 0; /* 0: hardware concurrency */
    /* binary dump cache: read_file() loads `<dump>.b200cache` when it matches the text dump, and writes it after
     * parsing otherwise.  Off by default (the reference leaves no side files). */
    bool dump_cache = false;
    std::string dump_cache_dir; /* empty: next to the dump */
    std::string dump_cache_path(const std::string &filepath) const;
    bool read_from_cache() const { return read_from_cache_; }
    /* on-disk copy of the (grid-independent) hot cross-section table; empty: always rebuild it */
    std::string hotcross_cache;
    bool hotcross_from_cache() const { return hotcross_from_cache_; }
    const Units &units() const { return units_; }
    const RunStats &stats() const { return stats_; }
    double bias_norm() const { return bias_norm_; }
    double max_tau_scatt0() const { return max_tau_scatt_; }
    int photon_n() const { return photon_n_; }
    const std::vector<double> &geom_det() const { return det_; }
    const std::vector<double> &hotcross_table() const { return hotcross_; }
    const std::array<double, kNESamp + 1> &f_table() const { return f_; }
    const std::array<double, kNESamp + 1> &k2_table() const { return k2_; }
    const std::array<double, kNESamp + 1> &weight_table() const { return weight_; }
    const std::vector<double> &nint_table() const { return nint_; }
    const std::vector<double> &dndlnu_max_table() const { return dndlnu_max_; }
    std::vector<double> &spectrum() { return spectrum_; } /* [6][200][13], harm::Spectrum field order */
    double luminosity() const { return luminosity_; }      /* logged by report_spectrum, reference :469 */
    double max_tau_scatt_reported() const { return max_tau_reported_; }

    /* individual init stages (reference harm_model.cpp:242-338, hotcross.cpp:60-79, jnu_mixed.cpp:57-73) */
    void init_geometry();
    void init_hotcross_table();
    void init_emiss_tables();
    void init_weight_table();
    void init_nint_table();
    /* the same tables through the C ABI on options.device (needs init_emiss_tables() first) */
    void init_tables_on_device(bool with_hotcross);

private:
    Header header_;
    Data data_;
    Units units_;
    double bias_norm_ = 0, rh_ = 0, max_tau_scatt_ = 0, d_tau_k_ = 0, x1_min_ = 0;
    int photon_n_;
    std::vector<double> det_, hotcross_, nint_, dndlnu_max_, spectrum_;
    std::array<double, kNESamp + 1> f_{}, k2_{}, weight_{};
    RunStats stats_;
    double luminosity_ = 0, max_tau_reported_ = 0;
    bool read_from_cache_ = false, hotcross_from_cache_ = false;
    bool load_hotcross_cache();
    void store_hotcross_cache() const;

    void fill_config(struct grmonty_b200_config &cfg) const;
    void run_share(int rank, int world, int device, void *nccl_comm, bool collect,
                   const std::function<bool(bool)> &sync);
    void apply_header(const double h[26], const std::string &filepath);
    bool load_dump_cache(const std::string &filepath);
    void store_dump_cache(const std::string &filepath, const double h[26]) const;

    void gcov(const double x[4], double g[4][4]) const;
    void gcon(const double x[4], double g[4][4]) const;
    double d_omega(double x2i, double x2f) const;
    struct ZoneFluid {
        double n_e, theta_e, b;
    };
    ZoneFluid fluid_zone(int i, int j) const;
    double k2_eval(double theta_e) const;
    double f_eval(double theta_e, double b, double nu) const;
};

void log_info(const char *fmt, ...);
void set_verbosity(int level); /* 0 trace .. 2 info (default) .. 6 off, like spdlog levels */

} /* namespace harm */
