/*
 * gm_tables.cuh -- grid-dependent initialisation tables on the device.
 *
 * Reference: HARMModel::init_geometry harm_model.cpp:242-266 (sqrt|det g_cov| at zone centres),
 * init_weight_table :268-306 (201 frequencies x all zones of F(K)-weighted emissivity), init_nint_table :308-338
 * (20 001 values of B theta_e^2 x 200 frequencies).  On one host core these cost 0.2 s + 0.2 s at 192^2 and the
 * zone sum grows with the grid (x28 at 1024^2); here they are three launches.
 *
 * Sums are reproducible: a thread adds its zones in a fixed strided order, the block combines the partial sums in a
 * fixed tree; there are no floating-point atomics.
 */
#pragma once
#include "gm_fluid.cuh"
#include "gm_radiation.cuh"

namespace gm {

/* zone centre: sqrt|det g| and the weight-table prefactor  jcst n_e B theta_e^2 / K2(theta_e) * dV * sqrt|det g|
 * (0 where the zone does not emit), plus theta_e and B for f_eval.  The metric is block diagonal in x2, so
 * det g = g22 * det3(t, x1, phi). */
__global__ void zone_table_kernel(GmParams P, double s_fac, double *det, double *fac, double *te, double *bb) {
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    if (z >= P.n0 * P.n1)
        return;
    const int i = z / P.n1, j = z % P.n1;
    double x[4];
    MetricCov g;
    Fluid f;
    fluid_zone(P, i, j, x, g, f);
    const double d3 = g.g00 * (g.g11 * g.g33 - g.g13 * g.g13) - g.g01 * (g.g01 * g.g33 - g.g13 * g.g03) +
                      g.g03 * (g.g01 * g.g13 - g.g11 * g.g03);
    const double dz = sqrt(fabs(g.g22 * d3));
    det[z] = dz;
    double fz = 0.0, tz = 0.0, bz = 0.0;
    if (!(f.n_e == 0.0 || f.theta_e < kThetaEMin)) {
        const double k2 = k2_eval(P, f.theta_e);
        fz = (kJcst * f.n_e * f.b * f.theta_e * f.theta_e / k2) * s_fac * dz;
        tz = f.theta_e;
        bz = f.b;
    }
    fac[z] = fz;
    te[z] = tz;
    bb[z] = bz;
}

/* one block per frequency sample: weight[k] = ln( sum_z fac_z F(K(nu_k, z)) / (h photon_n) ) */
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) weight_table_kernel(GmParams P, const double *__restrict__ fac,
                                                            const double *__restrict__ te,
                                                            const double *__restrict__ bb, double *weight) {
    __shared__ double part[BLOCK];
    const int k = blockIdx.x;
    const double nu = exp(k * P.d_l_nu + P.l_nu_min);
    const int nz = P.n0 * P.n1;
    double sum = 0.0;
    for (int z = threadIdx.x; z < nz; z += BLOCK) {
        const double t = __ldg(te + z);
        if (t != 0.0)
            sum += __ldg(fac + z) * f_eval(P, t, __ldg(bb + z), nu);
    }
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int s = BLOCK / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s)
            part[threadIdx.x] += part[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0)
        weight[k] = log(part[0] / (kHPL * P.photon_n));
}

/* one thread per B theta_e^2 sample (reference :318-336) */
__global__ void nint_table_kernel(GmParams P, const double *__restrict__ weight, double n_fac, double *nint,
                                  double *dndlnu_max) {
    __shared__ double nu_j[kNESamp], ew_j[kNESamp];
    for (int j = threadIdx.x; j < kNESamp; j += blockDim.x) {
        nu_j[j] = exp(j * P.d_l_nu + P.l_nu_min);
        ew_j[j] = exp(weight[j]) + 1.0e-100;
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > kNint)
        return;
    const double b_mag = exp(i * P.d_l_b + P.l_b_min);
    double acc = 0.0, mx = 0.0;
    for (int j = 0; j < kNESamp; ++j) {
        const double dn = f_eval(P, 1.0, b_mag, nu_j[j]) / ew_j[j];
        if (dn > mx)
            mx = dn;
        acc += P.d_l_nu * dn;
    }
    nint[i] = log(acc * n_fac);
    dndlnu_max[i] = log(mx);
}

} /* namespace gm */
