/*
 * oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A thin extern "C" shim around the UNMODIFIED reference CPU sources under
 * /root/reference/cuda_grmonty (compiled where they lie, see oracle/Makefile).
 * It exists so tests, golden-vector generation and the bench's `--impl reference`
 * arm can call the reference's own functions.  Nothing here re-implements any
 * physics: every entry point forwards to a reference function.
 *
 * Private members of harm::HARMModel (harm_model.hpp:87-475) are reached with the
 * explicit-template-instantiation idiom (legal C++: an explicit instantiation may
 * name private members), so no reference source is edited or copied.
 *
 * One model per process (the reference keeps function-static state,
 * harm_model.cpp:707-709,795).
 */
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <cmath>
#include <fstream>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "spdlog/spdlog.h"

#include "cuda_grmonty/consts.hpp"
#include "cuda_grmonty/harm_data.hpp"
#include "cuda_grmonty/harm_model.hpp"
#include "cuda_grmonty/hotcross.hpp"
#include "cuda_grmonty/jnu_mixed.hpp"
#include "cuda_grmonty/monty_rand.hpp"
#include "cuda_grmonty/ndarray.hpp"
#include "cuda_grmonty/photon.hpp"
#include "cuda_grmonty/proba.hpp"
#include "cuda_grmonty/radiation.hpp"
#include "cuda_grmonty/tetrads.hpp"

using M = harm::HARMModel;
using A2 = ndarray::NDArray<double, 2>;
using Vec4 = double[4];

/* ---- private-member access ------------------------------------------------------------ */
template <typename Tag, typename Tag::type P> struct Rob {
    friend typename Tag::type get(Tag) { return P; }
};
#define ROB(TAG, TYPE, MEMBER)          \
    struct TAG {                        \
        using type = TYPE;              \
        friend type get(TAG);           \
    };                                  \
    template struct Rob<TAG, &M::MEMBER>

/* methods */
ROB(T_init_geometry, void (M::*)(), init_geometry);
ROB(T_init_weight, void (M::*)(), init_weight_table);
ROB(T_init_nint, void (M::*)(), init_nint_table);
ROB(T_gcon, void (M::*)(const Vec4 &, A2 &) const, gcon_func);
ROB(T_gcov, void (M::*)(const Vec4 &, A2 &) const, gcov_func);
ROB(T_fluid_zone, harm::FluidZone (M::*)(int, int) const, get_fluid_zone);
ROB(T_fluid_params, harm::FluidParams (M::*)(const Vec4 &, const A2 &) const, get_fluid_params);
ROB(T_sample_zone, photon::InitPhoton (M::*)(harm::Zone &), sample_zone_photon);
using MakeSpFn = std::tuple<photon::InitPhoton, bool> (M::*)();
ROB(T_make_sp, MakeSpFn, make_super_photon);
ROB(T_track, void (M::*)(photon::Photon &), track_super_photon);
ROB(T_scatter,
    void (M::*)(photon::Photon &, photon::Photon &, const harm::FluidParams &, const A2 &, double) const,
    scatter_super_photon);
ROB(T_sample_scattered, void (M::*)(const Vec4 &, Vec4 &, Vec4 &) const, sample_scattered_photon);
ROB(T_push, void (M::*)(photon::Photon &, double, int), push_photon);
ROB(T_record, void (M::*)(const photon::Photon &, int), record_super_photon);
using InitZoneFn = std::tuple<double, double> (M::*)(int, int) const;
ROB(T_init_zone, InitZoneFn, init_zone);
ROB(T_bias, double (M::*)(double, double) const, bias_func);
ROB(T_conn, void (M::*)(const Vec4 &, double (&)[4][4][4]), get_connection);
ROB(T_dkdlam, void (M::*)(const Vec4 &, const Vec4 &, Vec4 &), init_dkdlam);
ROB(T_stop, bool (M::*)(photon::Photon &) const, stop_criterion);
ROB(T_step, double (M::*)(const Vec4 &, const Vec4 &), step_size);
ROB(T_domega, double (M::*)(double, double) const, d_omega_func);
/* data */
ROB(D_header, harm::Header M::*, header_);
ROB(D_data, harm::Data M::*, data_);
ROB(D_units, harm::Units M::*, units_);
ROB(D_bias_norm, double M::*, bias_norm_);
ROB(D_rh, double M::*, rh_);
ROB(D_max_tau, double M::*, max_tau_scatt_);
ROB(D_dtauk, double M::*, d_tau_k_);
ROB(D_x1min, double M::*, x1_min_);
ROB(D_ncreated, uint64_t M::*, n_super_photon_created_);
ROB(D_nscatt, uint64_t M::*, n_super_photon_scatt_);
ROB(D_nrec, uint64_t M::*, n_super_photon_recorded_);
ROB(D_geom, harm::Geometry M::*, geometry_);
ROB(D_hotcross, A2 M::*, hotcross_table_);
using Arr201 = std::array<double, consts::n_e_samp + 1>;
using ArrNint = std::array<double, consts::nint + 1>;
ROB(D_f, Arr201 M::*, f_);
ROB(D_k2, Arr201 M::*, k2_);
ROB(D_weight, Arr201 M::*, weight_);
ROB(D_nint, ArrNint M::*, nint_);
ROB(D_dnmax, ArrNint M::*, dndlnu_max_);
using SpecArr = harm::Spectrum[consts::n_th_bins][consts::n_e_bins];
ROB(D_spec, SpecArr M::*, spectrum_);

static std::unique_ptr<M> g_model;
#define MM (*g_model)
#define CALL(TAG) (MM.*get(TAG{}))
#define FIELD(TAG) (MM.*get(TAG{}))

/* flat photon layout shared with tests: 24 doubles in photon.hpp:19-36 order + n_scatt as double */
static void photon_from_flat(const double *p, photon::Photon &ph) {
    for (int i = 0; i < 4; ++i) {
        ph.x[i] = p[i];
        ph.k[i] = p[4 + i];
        ph.dkdlam[i] = p[8 + i];
    }
    ph.w = p[12];
    ph.e = p[13];
    ph.l = p[14];
    ph.x1i = p[15];
    ph.x2i = p[16];
    ph.tau_abs = p[17];
    ph.tau_scatt = p[18];
    ph.n_e_0 = p[19];
    ph.theta_e_0 = p[20];
    ph.b_0 = p[21];
    ph.e_0 = p[22];
    ph.e_0_s = p[23];
    ph.n_scatt = static_cast<int>(p[24]);
}
static void photon_to_flat(const photon::Photon &ph, double *p) {
    for (int i = 0; i < 4; ++i) {
        p[i] = ph.x[i];
        p[4 + i] = ph.k[i];
        p[8 + i] = ph.dkdlam[i];
    }
    p[12] = ph.w;
    p[13] = ph.e;
    p[14] = ph.l;
    p[15] = ph.x1i;
    p[16] = ph.x2i;
    p[17] = ph.tau_abs;
    p[18] = ph.tau_scatt;
    p[19] = ph.n_e_0;
    p[20] = ph.theta_e_0;
    p[21] = ph.b_0;
    p[22] = ph.e_0;
    p[23] = ph.e_0_s;
    p[24] = ph.n_scatt;
}

extern "C" {

int ref_create(int photon_n, double mass_unit, int verbose) {
    spdlog::set_level(verbose ? spdlog::level::info : spdlog::level::warn);
    g_model = std::make_unique<M>(photon_n, mass_unit);
    return 0;
}

int ref_read_file(const char *path) {
    try {
        MM.read_file(path);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "ref_read_file: %s\n", e.what());
        return -1;
    }
    return 0;
}

/* init() = harm_model.cpp:234-240, but the 33 s hotcross::init_table result is cached on disk (same doubles). */
int ref_init(const char *hotcross_cache) {
    CALL(T_init_geometry)();
    A2 &tab = FIELD(D_hotcross);
    const size_t n = (consts::hotcross::n_w + 1) * (consts::hotcross::n_t + 1);
    bool loaded = false;
    if (hotcross_cache && hotcross_cache[0]) {
        std::ifstream in(hotcross_cache, std::ios::binary);
        if (in.is_open()) {
            in.read(reinterpret_cast<char *>(tab.data()), n * sizeof(double));
            loaded = static_cast<size_t>(in.gcount()) == n * sizeof(double);
        }
    }
    if (!loaded) {
        hotcross::init_table(tab);
        if (hotcross_cache && hotcross_cache[0]) {
            std::ofstream out(hotcross_cache, std::ios::binary);
            out.write(reinterpret_cast<const char *>(tab.data()), n * sizeof(double));
        }
    }
    jnu_mixed::init_emiss_tables(FIELD(D_f), FIELD(D_k2));
    CALL(T_init_weight)();
    CALL(T_init_nint)();
    return loaded ? 1 : 0;
}

void ref_rng_init(int seed) { monty_rand::init(seed); }
double ref_uniform() { return monty_rand::uniform(); }
double ref_chi_sq(int dof) { return monty_rand::chi_sq(dof); }

/* ---- scalars / tables ----------------------------------------------------------------- */
/* out[0..]: n0 n1 x_start1 x_start2 dx1 dx2 dx3 x_stop1 x_stop2 a h_slope r_0 gamma */
void ref_get_header(double *out) {
    const harm::Header &h = FIELD(D_header);
    out[0] = h.n[0];
    out[1] = h.n[1];
    out[2] = h.x_start[1];
    out[3] = h.x_start[2];
    out[4] = h.dx[1];
    out[5] = h.dx[2];
    out[6] = h.dx[3];
    out[7] = h.x_stop[1];
    out[8] = h.x_stop[2];
    out[9] = h.a;
    out[10] = h.h_slope;
    out[11] = h.r_0;
    out[12] = h.gamma;
}
/* out: mass_unit l_unit t_unit rho_unit u_unit b_unit theta_e_unit n_e_unit */
void ref_get_units(double *out) {
    const harm::Units &u = FIELD(D_units);
    out[0] = u.mass_unit;
    out[1] = u.l_unit;
    out[2] = u.t_unit;
    out[3] = u.rho_unit;
    out[4] = u.u_unit;
    out[5] = u.b_unit;
    out[6] = u.theta_e_unit;
    out[7] = u.n_e_unit;
}
/* out: bias_norm rh max_tau_scatt d_tau_k x1_min */
void ref_get_scalars(double *out) {
    out[0] = FIELD(D_bias_norm);
    out[1] = FIELD(D_rh);
    out[2] = FIELD(D_max_tau);
    out[3] = FIELD(D_dtauk);
    out[4] = FIELD(D_x1min);
}
void ref_set_bias_stats(double max_tau_scatt, uint64_t n_scatt, uint64_t n_recorded) {
    FIELD(D_max_tau) = max_tau_scatt;
    FIELD(D_nscatt) = n_scatt;
    FIELD(D_nrec) = n_recorded;
}
void ref_get_counters(uint64_t *out) {
    out[0] = FIELD(D_ncreated);
    out[1] = FIELD(D_nscatt);
    out[2] = FIELD(D_nrec);
}
/* which: 0 k_rho 1 u 2 u_1 3 u_2 4 u_3 5 b_1 6 b_2 7 b_3 8 geometry.det */
void ref_get_grid(int which, double *out) {
    const harm::Data &d = FIELD(D_data);
    const A2 *arr[9] = {&d.k_rho, &d.u, &d.u_1, &d.u_2, &d.u_3, &d.b_1, &d.b_2, &d.b_3, &FIELD(D_geom).det};
    const harm::Header &h = FIELD(D_header);
    std::memcpy(out, arr[which]->data(), sizeof(double) * h.n[0] * h.n[1]);
}
/* which: 0 hotcross[221*81] 1 f[201] 2 k2[201] 3 weight[201] 4 nint[20001] 5 dndlnu_max[20001] */
void ref_get_table(int which, double *out) {
    switch (which) {
    case 0:
        std::memcpy(out, FIELD(D_hotcross).data(), sizeof(double) * 221 * 81);
        break;
    case 1:
        std::memcpy(out, FIELD(D_f).data(), sizeof(double) * 201);
        break;
    case 2:
        std::memcpy(out, FIELD(D_k2).data(), sizeof(double) * 201);
        break;
    case 3:
        std::memcpy(out, FIELD(D_weight).data(), sizeof(double) * 201);
        break;
    case 4:
        std::memcpy(out, FIELD(D_nint).data(), sizeof(double) * 20001);
        break;
    case 5:
        std::memcpy(out, FIELD(D_dnmax).data(), sizeof(double) * 20001);
        break;
    }
}
/* 13 doubles per bin in harm_data.hpp:129-143 order */
void ref_get_spectrum(double *out) {
    static_assert(sizeof(harm::Spectrum) == 13 * sizeof(double));
    std::memcpy(out, &FIELD(D_spec)[0][0], sizeof(harm::Spectrum) * 6 * 200);
}
void ref_clear_spectrum() { std::memset(&FIELD(D_spec)[0][0], 0, sizeof(harm::Spectrum) * 6 * 200); }

/* ---- geometry ---------------------------------------------------------------------------- */
void ref_gcov(const double *x, double *out16) {
    A2 g({4, 4});
    Vec4 xx = {x[0], x[1], x[2], x[3]};
    CALL(T_gcov)(xx, g);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            out16[4 * i + j] = g(i, j);
}
void ref_gcon(const double *x, double *out16) {
    A2 g({4, 4});
    Vec4 xx = {x[0], x[1], x[2], x[3]};
    CALL(T_gcon)(xx, g);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            out16[4 * i + j] = g(i, j);
}
/* full symmetric 4x4x4 (the reference fills only j<=k; we mirror for convenience) */
void ref_get_connection(const double *x, double *out64) {
    double lc[4][4][4];
    Vec4 xx = {x[0], x[1], x[2], x[3]};
    CALL(T_conn)(xx, lc);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            for (int k = 0; k < 4; ++k)
                out64[16 * i + 4 * j + k] = (j <= k) ? lc[i][j][k] : lc[i][k][j];
}
void ref_init_dkdlam(const double *x, const double *k, double *dk) {
    Vec4 xx = {x[0], x[1], x[2], x[3]};
    Vec4 kk = {k[0], k[1], k[2], k[3]};
    Vec4 out;
    CALL(T_dkdlam)(xx, kk, out);
    std::memcpy(dk, out, sizeof(out));
}
double ref_step_size(const double *x, const double *k) {
    Vec4 xx = {x[0], x[1], x[2], x[3]};
    Vec4 kk = {k[0], k[1], k[2], k[3]};
    return CALL(T_step)(xx, kk);
}
void ref_push_photon(double *flat, double dl) {
    photon::Photon ph;
    photon_from_flat(flat, ph);
    CALL(T_push)(ph, dl, 0);
    photon_to_flat(ph, flat);
}
/* returns 1 if stopped; consumes reference RNG draws exactly like the reference */
int ref_stop_criterion(double *flat) {
    photon::Photon ph;
    photon_from_flat(flat, ph);
    bool s = CALL(T_stop)(ph);
    photon_to_flat(ph, flat);
    return s ? 1 : 0;
}
double ref_d_omega(double x2i, double x2f) { return CALL(T_domega)(x2i, x2f); }

/* ---- fluid -------------------------------------------------------------------------------- */
/* out19: n_e theta_e b u_con[4] u_cov[4] b_con[4] b_cov[4]; only out[0] valid if outside the grid */
void ref_get_fluid_params(const double *x, double *out19) {
    A2 g({4, 4});
    Vec4 xx = {x[0], x[1], x[2], x[3]};
    CALL(T_gcov)(xx, g);
    harm::FluidParams fp = CALL(T_fluid_params)(xx, g);
    out19[0] = fp.n_e;
    if (fp.n_e == 0.0) {
        for (int i = 1; i < 19; ++i)
            out19[i] = 0.0;
        return;
    }
    out19[1] = fp.theta_e;
    out19[2] = fp.b;
    for (int i = 0; i < 4; ++i) {
        out19[3 + i] = fp.u_con[i];
        out19[7 + i] = fp.u_cov[i];
        out19[11 + i] = fp.b_con[i];
        out19[15 + i] = fp.b_cov[i];
    }
}
/* out11: n_e theta_e b u_con[4] b_con[4] */
void ref_get_fluid_zone(int i, int j, double *out11) {
    harm::FluidZone fz = CALL(T_fluid_zone)(i, j);
    out11[0] = fz.n_e;
    out11[1] = fz.theta_e;
    out11[2] = fz.b;
    for (int m = 0; m < 4; ++m) {
        out11[3 + m] = fz.u_con[m];
        out11[7 + m] = fz.b_con[m];
    }
}
void ref_init_zone(int i, int j, double *out2) {
    auto [nz, dnmax] = CALL(T_init_zone)(i, j);
    out2[0] = nz;
    out2[1] = dnmax;
}
double ref_bias_func(double te, double w) { return CALL(T_bias)(te, w); }

/* ---- radiation (public free functions of the reference) --------------------------------- */
double ref_bk_angle(const double *x, const double *k, const double *ucov, const double *bcov, double b) {
    Vec4 xx = {x[0], x[1], x[2], x[3]}, kk = {k[0], k[1], k[2], k[3]};
    Vec4 uu = {ucov[0], ucov[1], ucov[2], ucov[3]}, bb = {bcov[0], bcov[1], bcov[2], bcov[3]};
    return radiation::bk_angle(xx, kk, uu, bb, b, FIELD(D_units).b_unit);
}
double ref_fluid_nu(const double *x, const double *k, const double *ucov) {
    Vec4 xx = {x[0], x[1], x[2], x[3]}, kk = {k[0], k[1], k[2], k[3]};
    Vec4 uu = {ucov[0], ucov[1], ucov[2], ucov[3]};
    return radiation::fluid_nu(xx, kk, uu);
}
double ref_alpha_inv_scatt(double nu, double te, double ne) {
    return radiation::alpha_inv_scatt(nu, te, ne, FIELD(D_hotcross));
}
double ref_alpha_inv_abs(double nu, double te, double ne, double b, double theta) {
    return radiation::alpha_inv_abs(nu, te, ne, b, theta, FIELD(D_k2));
}
double ref_synch(double nu, double ne, double te, double b, double theta) {
    return jnu_mixed::synch(nu, ne, te, b, theta, FIELD(D_k2));
}
double ref_k2_eval(double te) { return jnu_mixed::k2_eval(te, FIELD(D_k2)); }
double ref_f_eval(double te, double b, double nu) { return jnu_mixed::f_eval(te, b, nu, FIELD(D_f)); }
double ref_hotcross_lkup(double w, double te) { return hotcross::total_compton_cross_lkup(w, te, FIELD(D_hotcross)); }

/* ---- tetrads / samplers ----------------------------------------------------------------- */
void ref_make_tetrad(const double *ucon, const double *trial, const double *gcov16, double *econ16, double *ecov16) {
    A2 g({4, 4});
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            g(i, j) = gcov16[4 * i + j];
    Vec4 uu = {ucon[0], ucon[1], ucon[2], ucon[3]};
    Vec4 tt = {trial[0], trial[1], trial[2], trial[3]};
    double ec[4][4], ev[4][4];
    tetrads::make_tetrad(uu, tt, g, ec, ev);
    std::memcpy(econ16, ec, sizeof(ec));
    std::memcpy(ecov16, ev, sizeof(ev));
}
void ref_sample_electron(const double *k, double te, double *p) {
    Vec4 kk = {k[0], k[1], k[2], k[3]}, pp;
    proba::sample_electron_distr_p(kk, pp, te);
    std::memcpy(p, pp, sizeof(pp));
}
double ref_sample_y(double te) { return proba::sample_y_distr(te); }
double ref_sample_mu(double beta) { return proba::sample_mu_distr(beta); }
double ref_sample_kn(double k0) { return proba::sample_klein_nishina(k0); }
double ref_sample_thomson() { return proba::sample_thomson(); }
void ref_sample_scattered_photon(const double *k, double *p, double *kp) {
    Vec4 kk = {k[0], k[1], k[2], k[3]}, pp = {p[0], p[1], p[2], p[3]}, out;
    CALL(T_sample_scattered)(kk, pp, out);
    std::memcpy(p, pp, sizeof(pp));
    std::memcpy(kp, out, sizeof(out));
}
/* scatter at the photon's position; child returned in flat_p (w preset by caller as in harm_model.cpp:983) */
int ref_scatter_super_photon(double *flat, double *flat_p) {
    photon::Photon ph, php;
    photon_from_flat(flat, ph);
    photon_from_flat(flat_p, php);
    A2 g({4, 4});
    CALL(T_gcov)(ph.x, g);
    harm::FluidParams fp = CALL(T_fluid_params)(ph.x, g);
    if (fp.n_e <= 0.0)
        return 0;
    CALL(T_scatter)(ph, php, fp, g, FIELD(D_units).b_unit);
    photon_to_flat(ph, flat);
    photon_to_flat(php, flat_p);
    return 1;
}

/* ---- generation / transport ----------------------------------------------------------- */
/* out16: x[4] k[4] w e l n_e_0 theta_e_0 b_0 e_0 n_scatt; returns 1 when the zone walk is finished */
static void init_photon_to_flat(const photon::InitPhoton &ip, double *out16) {
    for (int i = 0; i < 4; ++i) {
        out16[i] = ip.x[i];
        out16[4 + i] = ip.k[i];
    }
    out16[8] = ip.w;
    out16[9] = ip.e;
    out16[10] = ip.l;
    out16[11] = ip.n_e_0;
    out16[12] = ip.theta_e_0;
    out16[13] = ip.b_0;
    out16[14] = ip.e_0;
    out16[15] = ip.n_scatt;
}
int ref_make_super_photon(double *out16) {
    auto [ip, quit] = CALL(T_make_sp)();
    if (!quit)
        init_photon_to_flat(ip, out16);
    return quit ? 1 : 0;
}
/* sample n photons from zone (i,j) with the given dn_max (fresh zone => tetrad recomputed) */
void ref_sample_zone_photons(int i, int j, double dn_max, int n, double *out16n) {
    harm::Zone z{.x_1 = i, .x_2 = j, .num_to_gen = n, .dn_max = dn_max, .first_photon = true};
    for (int m = 0; m < n; ++m) {
        photon::InitPhoton ip = CALL(T_sample_zone)(z);
        init_photon_to_flat(ip, out16n + 16 * m);
    }
}
void ref_track_super_photon(double *flat) {
    photon::Photon ph;
    photon_from_flat(flat, ph);
    CALL(T_track)(ph);
    photon_to_flat(ph, flat);
}
void ref_record_super_photon(const double *flat) {
    photon::Photon ph;
    photon_from_flat(flat, ph);
    CALL(T_record)(ph, 0);
}
/* returns wall seconds of run_simulation (harm_model.cpp:340-414) */
double ref_run_simulation() {
    auto t0 = std::chrono::steady_clock::now();
    MM.run_simulation();
    std::chrono::duration<double> dt = std::chrono::steady_clock::now() - t0;
    return dt.count();
}
/* A BOUNDED SAMPLE of a run at the model's own photon_n (bench.py --impl reference / cpu_baseline): every zone is
 * visited in the reference's order and handled exactly as the reference's CPU loop handles it -- init_zone +
 * stochastic rounding (get_zone, harm_model.cpp:673-704), sample_zone_photon per primary (make_super_photon
 * :794-811), the InitPhoton -> Photon copy and track_super_photon (run_simulation :366-393) -- except that the
 * zone's expected photon count is divided by `thin` before the rounding, so the sample holds 1/thin of the run's
 * primaries spread over all zones like the run itself.  Weights and tables are those of the full-photon_n run;
 * nothing is re-implemented, every call is a reference function.  Returns wall seconds; out[0] = primaries created,
 * out[1] = zones that emitted. */
double ref_run_thinned(double thin, uint64_t *out) {
    const harm::Header &h = FIELD(D_header);
    const int n0 = static_cast<int>(h.n[0]), n1 = static_cast<int>(h.n[1]);
    uint64_t created = 0, emitting = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < n0; ++i)
        for (int j = 0; j < n1; ++j) {
            auto [d_full, dn_max] = CALL(T_init_zone)(i, j);
            const double d_num = d_full / thin;
            int num = static_cast<int>(d_num);
            if (std::fmod(d_num, 1.0) > monty_rand::uniform())
                ++num;
            if (num > 0)
                ++emitting;
            harm::Zone z{.x_1 = i, .x_2 = j, .num_to_gen = num, .dn_max = dn_max, .first_photon = true};
            for (; z.num_to_gen > 0; --z.num_to_gen) {
                photon::InitPhoton ip = CALL(T_sample_zone)(z);
                photon::Photon ph;
                for (int m = 0; m < consts::n_dim; ++m) {
                    ph.x[m] = ip.x[m];
                    ph.k[m] = ip.k[m];
                }
                ph.w = ip.w;
                ph.e = ip.e;
                ph.e_0 = ip.e_0;
                ph.e_0_s = ip.e;
                ph.l = ip.l;
                ph.tau_scatt = 0.0;
                ph.tau_abs = 0.0;
                ph.x1i = ip.x[1];
                ph.x2i = ip.x[2];
                ph.n_e_0 = ip.n_e_0;
                ph.b_0 = ip.b_0;
                ph.theta_e_0 = ip.theta_e_0;
                ph.n_scatt = 0;
                CALL(T_track)(ph);
                ++created;
            }
        }
    FIELD(D_ncreated) += created;
    out[0] = created;
    out[1] = emitting;
    std::chrono::duration<double> dt = std::chrono::steady_clock::now() - t0;
    return dt.count();
}
int ref_report_spectrum(const char *path) {
    MM.report_spectrum(path);
    return 0;
}

} /* extern "C" */

#ifdef REF_HARNESS_MAIN
/* CLI twin of main.cpp:43-53 without abseil; adds --seed, --hotcross_cache, --spectrum_bin. */
static const char *arg_value(int argc, char **argv, const char *name, const char *dflt) {
    std::string a1 = std::string("--") + name, a2 = std::string("-") + name;
    for (int i = 1; i < argc; ++i) {
        std::string s = argv[i];
        if ((s == a1 || s == a2) && i + 1 < argc)
            return argv[i + 1];
        if (s.rfind(a1 + "=", 0) == 0)
            return argv[i] + a1.size() + 1;
        if (s.rfind(a2 + "=", 0) == 0)
            return argv[i] + a2.size() + 1;
    }
    return dflt;
}
int main(int argc, char **argv) {
    int photon_n = std::atoi(arg_value(argc, argv, "photon_n", "5000000"));
    double mass_unit = std::atof(arg_value(argc, argv, "mass_unit", "4e19"));
    std::string dump = arg_value(argc, argv, "harm_dump_path", "");
    std::string spec = arg_value(argc, argv, "spectrum_path", "");
    std::string spec_bin = arg_value(argc, argv, "spectrum_bin", "");
    std::string cache = arg_value(argc, argv, "hotcross_cache", "");
    int seed = std::atoi(arg_value(argc, argv, "seed", "123"));
    int verbose = std::atoi(arg_value(argc, argv, "verbose", "0"));
    /* --thin T: a bounded sample of the run, 1/T of its primaries over all zones (ref_run_thinned) */
    double thin = std::atof(arg_value(argc, argv, "thin", "0"));
    /* --bias_stats MAX_TAU,N_SCATT,N_REC: start the running bias statistics (harm_model.cpp:1296-1320,1391-1404) from
     * the end state of a complete run instead of from an empty one -- a bounded sample then does the per-primary work
     * of the full run's late phase.  The two counts are subtracted again from the counters printed below. */
    std::string bias_stats = arg_value(argc, argv, "bias_stats", "");
    double bs_tau = 0.0;
    unsigned long long bs_scatt = 0, bs_rec = 0;
    const bool seeded = std::sscanf(bias_stats.c_str(), "%lf,%llu,%llu", &bs_tau, &bs_scatt, &bs_rec) == 3;
    ref_create(photon_n, mass_unit, verbose);
    if (ref_read_file(dump.c_str()) != 0)
        return 1;
    auto t0 = std::chrono::steady_clock::now();
    ref_init(cache.c_str());
    std::chrono::duration<double> t_init = std::chrono::steady_clock::now() - t0;
    ref_rng_init(seed);
    if (seeded)
        ref_set_bias_stats(bs_tau, bs_scatt, bs_rec);
    uint64_t thin_out[2] = {0, 0};
    double t_run = thin > 0.0 ? ref_run_thinned(thin, thin_out) : ref_run_simulation();
    if (!spec.empty())
        ref_report_spectrum(spec.c_str());
    uint64_t c[3];
    ref_get_counters(c);
    if (seeded) {
        c[1] -= bs_scatt;
        c[2] -= bs_rec;
    }
    double sc[5];
    ref_get_scalars(sc);
    if (!spec_bin.empty()) {
        std::vector<double> s(6 * 200 * 13);
        ref_get_spectrum(s.data());
        std::ofstream out(spec_bin, std::ios::binary);
        out.write(reinterpret_cast<const char *>(s.data()), s.size() * sizeof(double));
    }
    std::printf("{\"impl\": \"reference-cpu\", \"photon_n\": %d, \"mass_unit\": %.17g, \"seed\": %d, "
                "\"created\": %llu, \"scattered\": %llu, \"recorded\": %llu, \"max_tau_scatt\": %.17g, "
                "\"init_s\": %.6f, \"run_s\": %.6f, \"thin\": %g, \"seeded_bias_stats\": %d}\n",
                photon_n, mass_unit, seed, (unsigned long long)c[0], (unsigned long long)c[1],
                (unsigned long long)c[2], sc[2], t_init.count(), t_run, thin, seeded ? 1 : 0);
    return 0;
}
#endif
