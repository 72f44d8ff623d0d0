/*
 * dto_oracle.c -- C restatement of the reference's ALGORITHM for the dominant piece of the path.
 * TEST INFRASTRUCTURE ONLY (checker + CPU baseline); never linked into the product.
 * PARITY UNPINNED by the reference itself (no Julia here; see oracle/dto_oracle.py header).
 *
 * What it restates (citations into /root/reference):
 *   f(x+, x, u, dt) = x+ - expv(dt, G(u), x)                     src/integrators/bilinear_integrator.jl:81
 *   Jacobian  = ForwardDiff.jacobian of f over zz=[z_k; z_k+1]   src/integrators/bilinear_integrator.jl:111-131
 *   Hessian   = ForwardDiff.hessian of zz -> mu' f(zz)           src/integrators/bilinear_integrator.jl:135-161
 * expv is ExponentialAction 0.2's truncated, scaled Taylor action (Al-Mohy & Higham 2011):
 *   F = B;  for each of s stages: for j = 1.. : B <- (t/(s j)) A B; F += B, stop when two successive
 *   terms fall below tol*||F||; the package is generic Julia so the dual numbers flow through t, A, B
 *   and branch decisions use primal values.
 * ForwardDiff is restated as second-order forward-mode "jets" (value, gradient[nd], packed upper
 * Hessian[nd(nd+1)/2]) seeded on the same 2z-vector; the reference's nested duals carry the full
 * nd x nd second-order block, so this port does about half the reference's arithmetic.
 *
 * all_dirs = 1 differentiates w.r.t. all 2z entries like the reference (which does not know that only
 * x, u, dt and x+ are active); all_dirs = 0 seeds only the active n+m+1 entries (a stronger CPU
 * implementation, reported separately by bench.py).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

typedef struct {
    int nd, nh, order, len; /* len = 1 + (order>=1)*nd + (order>=2)*nh */
} jd_t;

static jd_t jd_make(int nd, int order) {
    jd_t d;
    d.nd = nd;
    d.nh = nd * (nd + 1) / 2;
    d.order = order;
    d.len = 1 + (order >= 1 ? nd : 0) + (order >= 2 ? d.nh : 0);
    return d;
}

/* y += a * b */
static void jet_fma(const jd_t* d, double* restrict y, const double* restrict a, const double* restrict b) {
    const int nd = d->nd;
    const double av = a[0], bv = b[0];
    y[0] += av * bv;
    if (d->order < 1) return;
    const double* ag = a + 1;
    const double* bg = b + 1;
    double* yg = y + 1;
    for (int i = 0; i < nd; ++i) yg[i] += av * bg[i] + bv * ag[i];
    if (d->order < 2) return;
    const double* aH = a + 1 + nd;
    const double* bH = b + 1 + nd;
    double* yH = y + 1 + nd;
    int p = 0;
    for (int i = 0; i < nd; ++i) {
        const double agi = ag[i], bgi = bg[i];
        for (int j = i; j < nd; ++j, ++p) yH[p] += av * bH[p] + bv * aH[p] + agi * bg[j] + ag[j] * bgi;
    }
}

/* y = a * b */
static void jet_mul(const jd_t* d, double* restrict y, const double* restrict a, const double* restrict b) {
    memset(y, 0, sizeof(double) * d->len);
    jet_fma(d, y, a, b);
}

/* One interval.  Outputs (any may be NULL): r[n], Jb[n x 2z] row-major, Hb[2z x 2z] row-major (of mu' f).
 * G: (m+1) matrices n x n ROW-major.  Returns the number of Taylor terms used. */
int dto_port_bilinear_interval(int n, int m, int z, int x_off, int u_off, int dt_off, const double* G, const double* zk,
                               const double* zk1, const double* mu, int order, int all_dirs, double* r, double* Jb, double* Hb) {
    const int nact = n + m + 1;
    const int nd = all_dirs ? 2 * z : nact;
    const jd_t d = jd_make(nd, order);
    /* direction index of a variable of zz (0..2z-1), or -1 */
    int* dir = (int*)malloc(sizeof(int) * 2 * z);
    for (int i = 0; i < 2 * z; ++i) dir[i] = all_dirs ? i : -1;
    if (!all_dirs) {
        for (int a = 0; a < n; ++a) dir[x_off + a] = a;
        for (int i = 0; i < m; ++i) dir[u_off + i] = n + i;
        dir[dt_off] = n + m;
    }
    const size_t L = (size_t)d.len;
    double* A = (double*)calloc((size_t)n * n * L, sizeof(double)); /* G(u) as jets */
    double* B = (double*)calloc((size_t)n * L, sizeof(double));
    double* Bn = (double*)calloc((size_t)n * L, sizeof(double));
    double* F = (double*)calloc((size_t)n * L, sizeof(double));
    double* t = (double*)calloc(L, sizeof(double));
    double* c = (double*)calloc(L, sizeof(double));
    double* tmp = (double*)calloc(L, sizeof(double));
    /* seed */
    t[0] = zk[dt_off];
    if (order >= 1 && dir[dt_off] >= 0) t[1 + dir[dt_off]] = 1.0;
    for (int rr = 0; rr < n; ++rr)
        for (int cc = 0; cc < n; ++cc) {
            double* a = A + ((size_t)rr * n + cc) * L;
            a[0] = G[rr * n + cc];
            for (int i = 0; i < m; ++i) {
                const double gi = G[(size_t)(1 + i) * n * n + rr * n + cc];
                a[0] += zk[u_off + i] * gi;
                if (order >= 1 && dir[u_off + i] >= 0) a[1 + dir[u_off + i]] = gi;
            }
        }
    for (int a = 0; a < n; ++a) {
        B[a * L] = zk[x_off + a];
        if (order >= 1 && dir[x_off + a] >= 0) B[a * L + 1 + dir[x_off + a]] = 1.0;
    }
    memcpy(F, B, sizeof(double) * n * L);
    /* scaling from the 1-norm of t*A (primal) */
    double nrm = 0.0;
    for (int cc = 0; cc < n; ++cc) {
        double s = 0.0;
        for (int rr = 0; rr < n; ++rr) s += fabs(A[((size_t)rr * n + cc) * L]);
        if (s > nrm) nrm = s;
    }
    nrm *= fabs(t[0]);
    int s_stages = nrm > 1.0 ? (int)ceil(nrm) : 1;
    const double tol = 2.220446049250313e-16;
    int terms = 0;
    for (int st = 0; st < s_stages; ++st) {
        double c1 = 0.0;
        for (int a = 0; a < n; ++a) c1 = fmax(c1, fabs(B[a * L]));
        for (int j = 1; j <= 60; ++j) {
            /* c = t / (s j) as a jet; Bn = c * (A * B) */
            for (size_t e = 0; e < L; ++e) c[e] = t[e] / ((double)s_stages * j);
            for (int rr = 0; rr < n; ++rr) {
                memset(tmp, 0, sizeof(double) * L);
                for (int cc = 0; cc < n; ++cc) jet_fma(&d, tmp, A + ((size_t)rr * n + cc) * L, B + (size_t)cc * L);
                jet_mul(&d, Bn + (size_t)rr * L, c, tmp);
            }
            double c2 = 0.0, fn = 0.0;
            for (int a = 0; a < n; ++a) {
                double* f = F + (size_t)a * L;
                const double* bn = Bn + (size_t)a * L;
                for (size_t e = 0; e < L; ++e) f[e] += bn[e];
                c2 = fmax(c2, fabs(bn[0]));
                fn = fmax(fn, fabs(f[0]));
            }
            double* sw = B;
            B = Bn;
            Bn = sw;
            ++terms;
            /* the derivative series lag the value series by up to two terms: keep the early exit of
             * expv but never before the partials' own terms are below tolerance */
            double dmax = 0.0;
            for (int a = 0; a < n; ++a)
                for (size_t e = 1; e < L; ++e) dmax = fmax(dmax, fabs(B[(size_t)a * L + e]));
            if (c1 + c2 <= tol * fn && dmax <= tol * fmax(fn, 1.0)) break;
            c1 = c2;
        }
        memcpy(B, F, sizeof(double) * n * L);
    }
    /* f = x+ - F */
    if (r)
        for (int a = 0; a < n; ++a) r[a] = zk1[x_off + a] - F[(size_t)a * L];
    if (Jb && order >= 1) {
        memset(Jb, 0, sizeof(double) * n * 2 * z);
        for (int a = 0; a < n; ++a) {
            for (int v = 0; v < 2 * z; ++v)
                if (dir[v] >= 0) Jb[(size_t)a * 2 * z + v] = -F[(size_t)a * L + 1 + dir[v]];
            Jb[(size_t)a * 2 * z + z + x_off + a] += 1.0;
        }
    }
    if (Hb && order >= 2 && mu) {
        memset(Hb, 0, sizeof(double) * 4 * z * z);
        for (int v = 0; v < 2 * z; ++v) {
            if (dir[v] < 0) continue;
            for (int w = v; w < 2 * z; ++w) {
                if (dir[w] < 0) continue;
                int i = dir[v], j = dir[w];
                if (i > j) {
                    int q = i;
                    i = j;
                    j = q;
                }
                const size_t p = (size_t)i * nd - (size_t)i * (i - 1) / 2 + (j - i);
                double s = 0.0;
                for (int a = 0; a < n; ++a) s -= mu[a] * F[(size_t)a * L + 1 + nd + p];
                Hb[(size_t)v * 2 * z + w] = s;
                Hb[(size_t)w * 2 * z + v] = s;
            }
        }
    }
    free(dir);
    free(A);
    free(B);
    free(Bn);
    free(F);
    free(t);
    free(c);
    free(tmp);
    return terms;
}

/* Timed loop over `count` intervals of a trajectory: residual (order 0) + Jacobian (order 1) + Hessian
 * (order 2) passes per interval -- the reference makes these three separate passes too
 * (eval_constraint / eval_constraint_jacobian / eval_hessian_lagrangian).  POSIX threads pull intervals
 * from a shared counter (the image's gcc cannot link libgomp).
 * Z: knot-major [N][z]; mu: [count][n].  Returns a checksum so the work cannot be optimised away. */
typedef struct {
    int n, m, z, x_off, u_off, dt_off, first, count, all_dirs;
    const double *G, *Z, *mu;
    atomic_int* next;
    double checksum;
} port_job;

static void* port_worker(void* arg) {
    port_job* J = (port_job*)arg;
    const int n = J->n, z = J->z;
    double* r = (double*)malloc(sizeof(double) * n);
    double* Jb = (double*)malloc(sizeof(double) * n * 2 * z);
    double* Hb = (double*)malloc(sizeof(double) * 4 * z * z);
    double s = 0.0;
    for (;;) {
        const int i = atomic_fetch_add(J->next, 1);
        if (i >= J->count) break;
        const int k = J->first + i;
        const double* zk = J->Z + (size_t)k * z;
        dto_port_bilinear_interval(n, J->m, z, J->x_off, J->u_off, J->dt_off, J->G, zk, zk + z, NULL, 0, J->all_dirs, r, NULL, NULL);
        dto_port_bilinear_interval(n, J->m, z, J->x_off, J->u_off, J->dt_off, J->G, zk, zk + z, NULL, 1, J->all_dirs, NULL, Jb, NULL);
        dto_port_bilinear_interval(n, J->m, z, J->x_off, J->u_off, J->dt_off, J->G, zk, zk + z, J->mu + (size_t)i * n, 2, J->all_dirs,
                                   NULL, NULL, Hb);
        for (int a = 0; a < n; ++a) s += r[a];
        for (int e = 0; e < n * 2 * z; ++e) s += Jb[e];
        for (int e = 0; e < 4 * z * z; ++e) s += Hb[e];
    }
    J->checksum = s;
    free(r);
    free(Jb);
    free(Hb);
    return NULL;
}

double dto_port_eval_intervals(int n, int m, int z, int x_off, int u_off, int dt_off, const double* G, const double* Z,
                               const double* mu, int first, int count, int all_dirs, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    atomic_int next = 0;
    port_job jobs[256];
    pthread_t tid[256];
    for (int t = 0; t < threads; ++t) {
        port_job j = {n, m, z, x_off, u_off, dt_off, first, count, all_dirs, G, Z, mu, &next, 0.0};
        jobs[t] = j;
        pthread_create(&tid[t], NULL, port_worker, &jobs[t]);
    }
    double checksum = 0.0;
    for (int t = 0; t < threads; ++t) {
        pthread_join(tid[t], NULL);
        checksum += jobs[t].checksum;
    }
    return checksum;
}

int dto_port_max_threads(void) {
    long v = sysconf(_SC_NPROCESSORS_ONLN);
    return v > 0 ? (int)v : 1;
}
