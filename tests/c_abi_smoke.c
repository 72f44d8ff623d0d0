/* Plain-C consumer of libdto_b200.so: proves the ABI of include/dto_b200.h is language-neutral (no Python, no C++).
 * Builds the descriptor of the README quick example by hand (2-state BilinearIntegrator, 1 drive, N = 4,
 * QuadraticRegularizer; /root/reference README.md:75-95), evaluates constraint + Jacobian through the host-pointer
 * callbacks and checks the entries that are known in closed form:
 *   - structure: nnz = (N-1) * n * 2z, first column rows 1..n (1-based), reference order
 *   - d r_k / d x_{k+1} = I  (bilinear_integrator.jl:81: x_{k+1} enters linearly)
 *   - with u = 0, dt = 0.1 and G_drift = [-0.1 1; -1 -0.1]: exp(dt G) = e^{-0.01} [cos .1  sin .1; -sin .1  cos .1],
 *     so d r_k / d x_k = -exp(dt G) and r_k = x_{k+1} - exp(dt G) x_k to 1e-12.
 * Build + run (needs a CUDA device; exits 77 without one):
 *   gcc -std=c99 -I include tests/c_abi_smoke.c -L directtrajopt.jl_b200/lib -ldto_b200 -lm -Wl,-rpath,$PWD/directtrajopt.jl_b200/lib -o build/c_abi_smoke
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dto_b200.h"

#define NK 4
#define NX 2
#define ZD 4 /* x[2], u[1], dt[1] */

int main(void) {
    /* column-major generators: drift, drive */
    double G[2 * NX * NX] = {-0.1, -1.0, 1.0, -0.1, /* drive */ 0.0, 1.0, 1.0, 0.0};
    double R[1] = {1.0};
    int32_t u_off[1] = {2};
    int32_t times[NK] = {1, 2, 3, 4};
    double Z0[NK * ZD];
    int k;
    for (k = 0; k < NK; ++k) {
        Z0[k * ZD + 0] = 0.3 + 0.1 * k;
        Z0[k * ZD + 1] = -0.2 + 0.05 * k;
        Z0[k * ZD + 2] = 0.0; /* u = 0: closed-form propagator */
        Z0[k * ZD + 3] = 0.1;
    }
    dto_integrator_desc integ;
    memset(&integ, 0, sizeof(integ));
    integ.kind = DTO_INT_BILINEAR;
    integ.x_off = 0;
    integ.x_dim = NX;
    integ.u_off = 2;
    integ.u_dim = 1;
    integ.t_off = -1;
    integ.G = G;
    dto_objective_desc obj;
    memset(&obj, 0, sizeof(obj));
    obj.kind = DTO_OBJ_QUADREG;
    obj.weight = 1.0;
    obj.n_vars = 1;
    obj.var_offs = u_off;
    obj.n_times = NK;
    obj.times = times;
    obj.R = R;
    dto_problem_desc d;
    memset(&d, 0, sizeof(d));
    d.abi_version = DTO_B200_ABI_VERSION;
    d.N = NK;
    d.z = ZD;
    d.dt_off = 3;
    d.batch = 1;
    d.eval_hessian = 1;
    d.device = -1;
    d.n_integrators = 1;
    d.n_objectives = 1;
    d.integrators = &integ;
    d.objectives = &obj;
    d.Z0 = Z0;

    if (dto_abi_version() != DTO_B200_ABI_VERSION) {
        fprintf(stderr, "ABI mismatch\n");
        return 1;
    }
    dto_handle* h = NULL;
    int rc = dto_create(&d, &h);
    if (rc == DTO_ERR_CUDA) {
        fprintf(stderr, "no CUDA device: %s\n", dto_last_error(NULL));
        return 77;
    }
    if (rc != DTO_OK) {
        fprintf(stderr, "dto_create failed (%d): %s\n", rc, dto_last_error(NULL));
        return 1;
    }
    dto_size_info si;
    dto_sizes(h, &si);
    if (si.n_vars != NK * ZD || si.n_cons != (NK - 1) * NX || si.nnz_jac != (NK - 1) * NX * 2 * ZD || si.nnz_hess != NK * ZD * (ZD + 1) / 2 + (NK - 1) * ZD * ZD) {
        fprintf(stderr, "unexpected sizes\n");
        return 1;
    }
    int64_t* rows = (int64_t*)malloc(sizeof(int64_t) * si.nnz_jac);
    int64_t* cols = (int64_t*)malloc(sizeof(int64_t) * si.nnz_jac);
    double* jac = (double*)malloc(sizeof(double) * si.nnz_jac);
    double* g = (double*)malloc(sizeof(double) * si.n_cons);
    double* hess = (double*)malloc(sizeof(double) * si.nnz_hess);
    double* mu = (double*)malloc(sizeof(double) * si.n_cons);
    double J = 0.0;
    int64_t e;
    for (e = 0; e < si.n_cons; ++e) mu[e] = 0.5;
    if (dto_jac_structure(h, rows, cols) != DTO_OK || dto_eval_constraint(h, Z0, g) != DTO_OK || dto_eval_jacobian(h, Z0, jac) != DTO_OK ||
        dto_eval_objective(h, Z0, &J) != DTO_OK || dto_eval_hessian(h, Z0, 1.0, mu, hess) != DTO_OK) {
        fprintf(stderr, "evaluation failed: %s\n", dto_last_error(h));
        return 1;
    }
    if (rows[0] != 1 || rows[1] != 2 || cols[0] != 1 || cols[1] != 1) {
        fprintf(stderr, "structure order\n");
        return 1;
    }
    const double ea = exp(-0.01), c = cos(0.1), s = sin(0.1);
    const double E[2][2] = {{ea * c, ea * s}, {-ea * s, ea * c}};
    double worst = 0.0;
    int checked = 0;
    for (e = 0; e < si.nnz_jac; ++e) {
        const int r = (int)rows[e] - 1, cc = (int)cols[e] - 1; /* 0-based */
        const int kr = r / NX, a = r % NX, kc = cc / ZD, l = cc % ZD;
        double want;
        if (kc == kr && l < NX) want = -E[a][l];              /* d r_k / d x_k */
        else if (kc == kr + 1 && l < NX) want = a == l ? 1.0 : 0.0; /* d r_k / d x_{k+1} */
        else if (kc == kr + 1) want = 0.0;
        else continue; /* u and dt columns: not closed-form here */
        if (fabs(jac[e] - want) > worst) worst = fabs(jac[e] - want);
        ++checked;
    }
    for (k = 0; k < NK - 1; ++k) {
        int a;
        for (a = 0; a < NX; ++a) {
            const double want = Z0[(k + 1) * ZD + a] - (E[a][0] * Z0[k * ZD] + E[a][1] * Z0[k * ZD + 1]);
            if (fabs(g[k * NX + a] - want) > worst) worst = fabs(g[k * NX + a] - want);
        }
    }
    /* J = sum_k 1/2 (dt u)^2 = 0 at u = 0 */
    if (fabs(J) > worst) worst = fabs(J);
    printf("c_abi_smoke: %d Jacobian entries + %d residuals checked, worst abs error %.3e, variant %s, launches %lld\n", checked, (NK - 1) * NX,
           worst, dto_kernel_variant(h, 0), (long long)dto_launch_count(h));
    dto_destroy(h);
    free(rows); free(cols); free(jac); free(g); free(hess); free(mu);
    return worst < 1e-12 ? 0 : 1;
}
