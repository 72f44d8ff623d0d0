// K2 body: DerivativeIntegrator  f = x+ - x - dt*xdot  (derivative_integrator.jl:45-86): residual and the full
// d x 2z Jacobian block (zeros included) of EVERY derivative integrator for one interval.  Used by analytic_kernel
// (assemble.cu) and, fused, by the small-state bilinear kernel (bilinear_octet.cu): written together with the
// bilinear columns of the same interval, a Jacobian column leaves the L2 as whole lines instead of 16-byte pieces.
#pragma once
#include "dto_internal.h"

__device__ __forceinline__ void analytic_interval(const DProb& P, const double* __restrict__ Z, double* __restrict__ g,
                                                  double* __restrict__ jac, int b, int kl, int lane, int gs) {
    const int z = P.z;
    const double* zk = Z + (long long)b * P.n_vars_local + (long long)kl * z;
    const double* zk1 = zk + z;
    if (P.halo != nullptr && kl + 1 == P.nK - 1) zk1 = P.halo;
    const double dt = zk[P.dt_off];
    for (int ii = 0; ii < P.n_int; ++ii) {
        const DInt& I = P.in[ii];
        if (I.kind != DTO_INT_DERIVATIVE) continue;
        const int d = I.n;
        if (g != nullptr) {
            double* gp = g + (long long)b * P.n_cons_local + I.row_off + (long long)kl * d;
            for (int a = lane; a < d; a += gs) gp[a] = zk1[I.x_off + a] - zk[I.x_off + a] - dt * zk[I.u_off + a];
        }
        if (jac != nullptr) {
            double* jp = jac + (long long)b * P.nnz_jac_local;
            const long long own_off = jac_own_off(P, kl, I.doff, d);
            const long long prev_off = jac_prev_off(P, kl + 1, I.doff);
            for (int e = lane; e < 2 * z * d; e += gs) {
                const int l = e / d, a = e % d;
                double v = 0.0;
                long long pos;
                if (l < z) {
                    if (l == I.x_off + a) v += -1.0;
                    if (l == I.u_off + a) v += -dt;
                    if (l == P.dt_off) v += -zk[I.u_off + a];
                    pos = jac_col(P, kl, l) + own_off + a;
                } else {
                    if (l - z == I.x_off + a) v = 1.0;
                    pos = jac_col(P, (kl + 1), (l - z)) + prev_off + a;
                }
                jp[pos] = v;
            }
        }
    }
}
