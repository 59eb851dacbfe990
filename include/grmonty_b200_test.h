/*
 * grmonty_b200_test.h -- test-only batch entry points of the B200 transport library.
 *
 * NOT part of the product ABI: these symbols exist only in libgrmonty_b200_test.so, the same sources as
 * libgrmonty_b200.so compiled with -DGRMONTY_B200_TEST_EXPORTS (cuda-grmonty_b200/__init__.py build_cuda).  The parity
 * tests call the device functions of the path one at a time through them (host arrays in / out) and compare with the
 * oracle and the reference's golden vectors; a context created by the test library is used with both headers.
 */
#ifndef GRMONTY_B200_TEST_H
#define GRMONTY_B200_TEST_H

#include "grmonty_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define GRMONTY_B200_PHOTON_FLAT 25  /* reference photon.hpp:19-36 order, n_scatt last */


/* x: [n][4] -> gcov [n][16], gcon [n][16], conn [n][64] (full symmetric); any output may be NULL */
int grmonty_b200_test_geometry(grmonty_b200_ctx *ctx, int64_t n, const double *x, double *gcov, double *gcon,
                               double *conn);
/* x,k: [n][4] -> dkdlam [n][4] (init_dkdlam), step [n] (step_size) */
int grmonty_b200_test_dkdlam_step(grmonty_b200_ctx *ctx, int64_t n, const double *x, const double *k,
                                  double *dkdlam, double *step);
/* photons: [n][25] in/out, dl: [n]; one full push_photon (with halving) per photon; attempts [n] may be NULL */
int grmonty_b200_test_push_photon(grmonty_b200_ctx *ctx, int64_t n, double *photons, const double *dl,
                                  int32_t *attempts);
/* photons: [n][25] in/out; nsteps x (step_size + push_photon) or until the photon leaves [r_h, 100];
 * every `stride` steps x[4] k[4] e_0_s are written to trace [n][nsteps/stride][9] (NaN where not reached) */
int grmonty_b200_test_trajectory(grmonty_b200_ctx *ctx, int64_t n, double *photons, int32_t nsteps,
                                 int32_t stride, double *trace);
/* x: [n][4] -> out [n][19]: n_e theta_e b u_con[4] u_cov[4] b_con[4] b_cov[4] (zeros if outside the grid) */
int grmonty_b200_test_fluid_params(grmonty_b200_ctx *ctx, int64_t n, const double *x, double *out);
/* args: [n][5] = nu theta_e n_e b theta -> out [n][5] = alpha_inv_scatt alpha_inv_abs synch k2_eval f_eval */
int grmonty_b200_test_radiation(grmonty_b200_ctx *ctx, int64_t n, const double *args, double *out);
/* args: [n][2] = w theta_e -> sigma [n] (total_compton_cross_lkup) */
int grmonty_b200_test_hotcross(grmonty_b200_ctx *ctx, int64_t n, const double *args, double *sigma);
/* k: [n][4], fluid: [n][19] (layout of test_fluid_params) -> theta [n], nu [n] */
int grmonty_b200_test_angles(grmonty_b200_ctx *ctx, int64_t n, const double *k, const double *fluid, double *theta,
                             double *nu);
/* in: [n][24] = gcov[16] u_con[4] trial[4] -> e_con [n][16], e_cov [n][16] */
int grmonty_b200_test_tetrad(grmonty_b200_ctx *ctx, int64_t n, const double *in, double *e_con, double *e_cov);
/* per-zone emission data computed at create: nz [n0*n1] (may be NULL), dn_max, num_to_gen */
int grmonty_b200_test_zones(grmonty_b200_ctx *ctx, double *nz, double *dn_max, int64_t *num_to_gen);
/* bias_func(theta_e, w) with the given frozen statistics; args [n][2] */
int grmonty_b200_test_bias(grmonty_b200_ctx *ctx, int64_t n, const double *args, double max_tau_scatt,
                           double n_scatt, double n_recorded, double *out);
/* birth state of primaries idx[n] -> photons [n][25] (dkdlam zero), rng [n][4] = id0 id1 id2 ctr */
int grmonty_b200_test_make_primaries(grmonty_b200_ctx *ctx, int64_t n, const int64_t *idx, double *photons,
                                     uint32_t *rng);
/* Track n given photons (and all their descendants) to completion with frozen bias statistics, recording into
 * the context's accumulators.  photons [n][25] in -> end state of each given photon out;
 * rng [n][4] = id0 id1 id2 ctr in/out; status [n]: bit0 recorded, bit1 scattered at least once, bit2 absorbed
 * or dropped. */
int grmonty_b200_test_track(grmonty_b200_ctx *ctx, int64_t n, double *photons, uint32_t *rng,
                            double max_tau_scatt, double n_scatt, double n_recorded, int32_t *status);
/* which: 0 uniform, 3..6 chi_sq(dof), 10 sample_y(p0), 11 sample_mu(p0), 12 klein_nishina(p0), 13 thomson,
 * 20 electron gamma, 21 electron mu (p0 = k0, p1 = theta_e), 22 scattered energy ratio, 23 scattered cosine.
 * Stream of sample i: primary stream `first_stream + i`. */
int grmonty_b200_test_samplers(grmonty_b200_ctx *ctx, int32_t which, double p0, double p1, int64_t first_stream,
                               int64_t n, double *out);
/* raw Philox4x32-10 blocks: ctr [n][4], key [n][2] -> out [n][4] */
int grmonty_b200_test_philox(grmonty_b200_ctx *ctx, int64_t n, const uint32_t *ctr, const uint32_t *key,
                             uint32_t *out);

/* pool accesses outside the pool seen by the checked accessors of the test build (csrc/gm_transport.cuh chk_slot);
 * reading resets the count */
int grmonty_b200_test_bounds_violations(grmonty_b200_ctx *ctx, uint32_t *count);

#ifdef __cplusplus
}
#endif
#endif /* GRMONTY_B200_TEST_H */
