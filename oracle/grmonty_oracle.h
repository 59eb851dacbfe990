/*
 * oracle/grmonty_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C (C99) CPU restatement of the reference's superphoton transport path
 * (m-torhan/cuda-grmonty, CPU branch of HARMModel::run_simulation,
 * cuda_grmonty/harm_model.cpp:362-405 and everything it calls).  Each function in
 * grmonty_oracle.c cites the reference file:line it follows.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker.  The product (csrc/, host/)
 * never links or calls it.
 *
 * Parity status: PINNED against the reference itself -- the unmodified reference CPU
 * sources are compiled in this container by oracle/Makefile (`make ref`) and
 * tests/test_oracle_vs_golden.py compares every function below with golden vectors
 * generated from that build (tests/golden/, generator oracle/make_golden.py).  The
 * reference's own tests hold no vectors for this path.
 *
 * Two deliberate differences from the reference, both mandated by the north star
 * (BASELINE.json) and shared with the CUDA path so that the two can be compared
 * photon by photon:
 *   1. random numbers come from counter-based Philox4x32-10 streams keyed by photon
 *      identity instead of one global mt19937 (reference monty_rand.cpp:19-31);
 *   2. the scattering-bias statistics (harm_model.cpp:1391-1404) are frozen per
 *      "generation" of primaries in ORC_STATS_FROZEN mode; ORC_STATS_LIVE mode keeps
 *      the reference's running statistics for statistical comparison with it.
 */
#ifndef GRMONTY_ORACLE_H
#define GRMONTY_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NDIM 4
#define ORC_N_TH_BINS 6
#define ORC_N_E_BINS 200
#define ORC_SPEC_FIELDS 13
#define ORC_N_E_SAMP 200
#define ORC_NINT 20000
#define ORC_HC_NW 220
#define ORC_HC_NT 80

/* Everything the transport path reads. Pointers are borrowed (caller keeps them alive). */
typedef struct orc_model {
    /* header subset, reference harm_data.hpp:19-44 */
    int n0, n1;
    double x_start1, x_start2, dx1, dx2, dx3, x_stop1, x_stop2;
    double a, h_slope, r_0;
    /* units, reference harm_data.hpp:63-72 */
    double l_unit, rho_unit, b_unit, theta_e_unit, n_e_unit;
    /* primitives, row-major [n0][n1], reference harm_data.hpp:49-58 */
    const double *k_rho, *u, *u_1, *u_2, *u_3, *b_1, *b_2, *b_3;
    const double *geom_det; /* sqrt|det g_cov| at zone centres, reference harm_model.cpp:261 */
    /* tables, reference harm_model.hpp:159-195 */
    const double *hotcross;   /* [221][81] log10 sigma */
    const double *f;          /* [201] */
    const double *k2;         /* [201] */
    const double *weight;     /* [201] */
    const double *nint;       /* [20001] */
    const double *dndlnu_max; /* [20001] */
    /* scalars */
    double photon_n;
    double bias_norm;
    double d_tau_k; /* reference harm_model.cpp:73 */
    double x1_min;  /* ln r_h, reference harm_model.cpp:228-229 */
    uint64_t seed;
    /* bias statistics (reference harm_model.hpp:130-138): values used by orc_bias_func ... */
    double bias_max_tau_scatt;
    double bias_n_scatt;
    double bias_n_recorded;
    /* ... and values accumulated by orc_record_super_photon */
    double acc_max_tau_scatt;
    uint64_t acc_n_scatt;
    uint64_t acc_n_recorded;
    int stats_mode; /* ORC_STATS_LIVE: bias_* track acc_* after every record */
    int zone_order; /* 1: process primaries in zone order like the reference; 0: Weyl-permuted order (CUDA path) */
    /* generation clock (CUDA path's tail bound): push attempts a photon lineage may make per
     * generation before it is suspended and continued in the next one; <= 0: unlimited (the reference) */
    int budget;
    /* generation schedule (shared with the CUDA path, see orc_generation_size) */
    int64_t gen_fine_from, gen_fine_div, gen_ramp, gen_budget_spread;
    struct orc_track_state *carry; /* suspended photons (owned by the model) */
    uint64_t n_carry, cap_carry;
    /* outputs */
    double spectrum[ORC_N_TH_BINS][ORC_N_E_BINS][ORC_SPEC_FIELDS]; /* field order of harm_data.hpp:129-143 */
    uint64_t n_created;
    /* diagnostics */
    uint64_t n_steps, n_push_attempts, n_interactions, n_scatter_events, n_tracked;
    /* Study knob for the generation-overlap design: with stats_lag = 1 a generation uses
     * the statistics that were frozen at the start of the PREVIOUS generation -- what a pipeline that starts
     * generation g+1 while the tail of generation g still runs would see.  0 (default) is the CUDA path's schedule. */
    int stats_lag;
} orc_model;

enum { ORC_STATS_FROZEN = 0, ORC_STATS_LIVE = 1 };

typedef struct orc_rng {
    uint32_t id[3];
    uint32_t ctr;
} orc_rng;

typedef struct orc_photon {
    double x[4], k[4], dkdlam[4];
    double w, e, l, x1i, x2i, tau_abs, tau_scatt, n_e_0, theta_e_0, b_0, e_0, e_0_s;
    int n_scatt;
    orc_rng rng;
} orc_photon;

/* loop state of track_super_photon, so that a photon can be suspended and resumed */
typedef struct orc_track_state {
    orc_photon ph;
    double alpha_scatti, alpha_absi, bi;
    int ne_pos, n_step, clock;
} orc_track_state;

typedef struct orc_fluid {
    double n_e, theta_e, b;
    double u_con[4], u_cov[4], b_con[4], b_cov[4];
} orc_fluid;

/* ---- RNG ------------------------------------------------------------------------------- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_rng_primary(orc_rng *r, uint64_t primary_index);
void orc_rng_zone(orc_rng *r, uint64_t zone_index);
double orc_uniform(const orc_model *m, orc_rng *r);
void orc_rng_child(const orc_model *m, orc_rng *parent, orc_rng *child);
double orc_chi_sq(const orc_model *m, orc_rng *r, int dof);

/* ---- geometry -------------------------------------------------------------------------- */
void orc_gcov(const orc_model *m, const double x[4], double g[4][4]);
void orc_gcon(const orc_model *m, const double x[4], double g[4][4]);
void orc_get_connection(const orc_model *m, const double x[4], double lconn[4][4][4]);
void orc_init_dkdlam(const orc_model *m, const double x[4], const double k[4], double dk[4]);
double orc_step_size(const orc_model *m, const double x[4], const double k[4]);
void orc_push_photon(orc_model *m, orc_photon *ph, double dl, int n);
int orc_stop_criterion(const orc_model *m, orc_photon *ph);

/* ---- fluid / radiation ----------------------------------------------------------------- */
void orc_get_fluid_params(const orc_model *m, const double x[4], double gcov[4][4], orc_fluid *f);
void orc_get_fluid_zone(const orc_model *m, int i, int j, orc_fluid *f);
void orc_init_zone(const orc_model *m, int i, int j, double *nz, double *dn_max);
double orc_bias_func(const orc_model *m, double theta_e, double w);
double orc_bk_angle(const orc_model *m, const double k[4], const double u_cov[4], const double b_cov[4], double b);
double orc_fluid_nu(const double k[4], const double u_cov[4]);
double orc_alpha_inv_scatt(const orc_model *m, double nu, double theta_e, double n_e);
double orc_alpha_inv_abs(const orc_model *m, double nu, double theta_e, double n_e, double b, double theta);
double orc_synch(const orc_model *m, double nu, double n_e, double theta_e, double b, double theta);
double orc_k2_eval(const orc_model *m, double theta_e);
double orc_f_eval(const orc_model *m, double theta_e, double b, double nu);
double orc_hotcross_lkup(const orc_model *m, double w, double theta_e);
double orc_hotcross_num(double w, double theta_e, double k2f);

/* ---- tetrads / scattering -------------------------------------------------------------- */
void orc_make_tetrad(const double u_con[4], double trial[4], double gcov[4][4], double e_con[4][4], double e_cov[4][4]);
void orc_sample_electron(const orc_model *m, orc_rng *r, const double k[4], double theta_e, double p[4]);
double orc_sample_y(const orc_model *m, orc_rng *r, double theta_e);
double orc_sample_mu(const orc_model *m, orc_rng *r, double beta_e);
double orc_sample_klein_nishina(const orc_model *m, orc_rng *r, double k0);
double orc_sample_thomson(const orc_model *m, orc_rng *r);
void orc_sample_scattered_photon(const orc_model *m, orc_rng *r, const double k[4], double p[4], double kp[4]);
/* returns 1 if the child is valid and must be tracked */
int orc_scatter_super_photon(const orc_model *m, orc_photon *ph, orc_photon *php, const orc_fluid *f, double gcov[4][4]);

/* ---- generation / transport / record ------------------------------------------------------ */
/* zone photon counts in zone order (i outer, j inner); returns the total */
uint64_t orc_zone_counts(const orc_model *m, int64_t *num_to_gen, double *dn_max);
void orc_sample_zone_photon(const orc_model *m, int i, int j, double dn_max, orc_rng *r, orc_photon *ph);
void orc_track_super_photon(orc_model *m, orc_photon *ph);
void orc_record_super_photon(orc_model *m, const orc_photon *ph);
/* Generation schedule shared with the CUDA path: size of the generation starting at position g_start. */
int64_t orc_generation_size(int64_t g_start, int64_t gen0, int64_t gen_cap, int64_t fine_from, int64_t fine_div,
                            int64_t ramp);
/* Run positions [first,last) of the processing sequence that satisfy j % world == rank; position j handles
 * primary orc_permute(j) (or j itself when m->zone_order). */
int64_t orc_perm_multiplier(int64_t total);
int64_t orc_permute(int64_t j, int64_t mult, int64_t total);
void orc_run(orc_model *m, int64_t first, int64_t last, int rank, int world, int64_t gen0, int64_t gen_cap);
/* track one primary by global index; optionally returns the flat birth state */
void orc_run_primary(orc_model *m, const int64_t *prefix, const double *dn_max, int64_t idx, int clock0);
void orc_make_primary(const orc_model *m, const int64_t *prefix, const double *dn_max, int64_t idx, orc_photon *ph);
void orc_clear_outputs(orc_model *m);

/* flat <-> struct helpers for ctypes (25 doubles: photon.hpp:19-36 order + n_scatt) */
void orc_photon_from_flat(const double *flat, orc_photon *ph);
void orc_photon_to_flat(const orc_photon *ph, double *flat);
orc_model *orc_model_alloc(void);
void orc_model_free(orc_model *m);
unsigned long orc_sizeof_model(void);

#ifdef __cplusplus
}
#endif
#endif
