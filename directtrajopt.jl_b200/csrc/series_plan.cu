// Series plan of the bilinear interval kernels: how many Taylor terms exp(dt G(u_k)) needs, per interval.
//
// The reference's `expv` (ExponentialAction.jl: Al-Mohy & Higham 2011) does not size its truncated Taylor series from
// ||A||_1 but from norms of POWERS of A, which bound ||A^k||^(1/k) far better for oscillatory or non-normal generators.
// Here: d2 = ||A^2||_1^(1/2).  Since ||A^(2j)|| <= ||A^2||^j and ||A^(2j+1)|| <= ||A^2||^j ||A||, the tail of the series is
// bounded by (||A||_1 / d2) sum_k d2^k / k!; one term beyond T(d2) absorbs the factor whenever ||A||_1 <= T(d2) + 1.
// At config c2 (n = 32): ||A||_1 = 1.0, d2 = 0.50 -- 16 terms instead of 19 for a 2^-53 tail.
//
// A = dt (G_0 + sum_i u_i G_i) is affine in u, so A^2 = dt^2 sum_{a<=b} w_a w_b S_ab with w = (1, u) and the symmetrised
// pair products S_ab = G_a G_b + G_b G_a (S_aa = G_a^2) formed ONCE at construction: a plan costs (npairs + m + 1) n^2
// multiply-adds per interval (half of ONE n^3 product at c2; < 1 % of an interval's work) instead of matrix products.
// One CTA per SM stages the matrices in shared memory; a warp per (problem, interval), lane = column, sums |.| down its
// columns.  plan = { alpha, theta1 = ||A||_1 }: the kernels take the stages from theta1 (round-off of e^|A|) and the terms
// per stage from alpha / stages, where alpha is the argument that makes T(alpha) = min(T(theta1), T(d2) + 1).
// Every variant, pass, range, batch member and shard reads the same plan: their results stay bit-identical to each other.
#include "dto_internal.h"
#include "series_tables.cuh"

namespace {

// (per-problem generator sets get no plan: the kernels keep ||dt G||_1 for them)
// NM = m + 1 matrices; GS lanes share one interval (GS = n when n divides 32: 32 / n intervals per warp, else 32)
template <int NM>
__global__ void __launch_bounds__(512) series_plan_kernel(DProb P, int ii, const double* __restrict__ Z, double2* __restrict__ plan, int gs) {
    extern __shared__ __align__(16) double sm[];
    constexpr int NP = NM * (NM + 1) / 2;
    const DInt& I = P.in[ii];
    const int n = I.n, nn = n * n;
    const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    {   // stage the matrices: 16-byte loads, 8 in flight per thread (the copy is latency-bound: every CTA reads the same lines)
        const double2* g2 = reinterpret_cast<const double2*>(I.Grm);
        const double2* s2 = reinterpret_cast<const double2*>(I.Spair);
        double2* d2 = reinterpret_cast<double2*>(sm);
        const int ng = NM * nn / 2, ns = NP * nn / 2;  // n is even wherever plans are made
#pragma unroll 8
        for (int e = threadIdx.x; e < ng; e += blockDim.x) d2[e] = g2[e];
#pragma unroll 8
        for (int e = threadIdx.x; e < ns; e += blockDim.x) d2[ng + e] = s2[e];
    }
    __syncthreads();
    const double *G = sm, *S = sm + NM * nn;
    const int per_warp = 32 / gs, sub = lane / gs, c0 = lane % gs;
    const long long nIc = min(P.kc1, P.nI) - P.kc0, items = (long long)P.batch * nIc;
    const long long tasks = (items + per_warp - 1) / per_warp;
    for (long long task = (long long)blockIdx.x * wpc + wic; task < tasks; task += (long long)gridDim.x * wpc) {
        const long long item = task * per_warp + sub;
        const bool live = item < items;
        const long long it = live ? item : items - 1;
        const int b = (int)(it / nIc), kk = P.kc0 + (int)(it % nIc);
        const double* zk = Z + (long long)b * P.n_vars_local + (long long)kk * P.z;
        const double dt = zk[P.dt_off];
        double w[NM], wp[NP];  // (1, u) and its pair products
        w[0] = 1.0;
#pragma unroll
        for (int a = 1; a < NM; ++a) w[a] = zk[I.u_off + a - 1];
        {
            int p = 0;
#pragma unroll
            for (int a = 0; a < NM; ++a)
#pragma unroll
                for (int bb = a; bb < NM; ++bb, ++p) wp[p] = w[a] * w[bb];
        }
        double c1 = 0.0, c2 = 0.0;  // largest column sums of |G(u)| and |G(u)^2|
        for (int c = c0; c < n; c += gs) {
            double s1 = 0.0, s2 = 0.0;
#pragma unroll 2
            for (int r = 0; r < n; ++r) {
                const int e = r * n + c;
                double v1 = 0.0, v2 = 0.0, v3 = 0.0;
#pragma unroll
                for (int a = 0; a < NM; ++a) v1 = fma(w[a], G[a * nn + e], v1);
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    if (p & 1) v3 = fma(wp[p], S[p * nn + e], v3);
                    else v2 = fma(wp[p], S[p * nn + e], v2);
                }
                s1 += fabs(v1);
                s2 += fabs(v2 + v3);
            }
            c1 = fmax(c1, s1);
            c2 = fmax(c2, s2);
        }
        for (int o = gs >> 1; o > 0; o >>= 1) {
            c1 = fmax(c1, __shfl_xor_sync(0xffffffffu, c1, o));
            c2 = fmax(c2, __shfl_xor_sync(0xffffffffu, c2, o));
        }
        const double theta1 = fabs(dt) * c1, d2 = fabs(dt) * sqrt(c2);
        // alpha: an argument with T(alpha) = min(T(theta1), T(d2) + 1)
        double alpha = theta1;
        if (theta1 < 1e8 && d2 < theta1) {
            const double inv_st = theta1 > 4.0 ? 1.0 / ceil(theta1 * 0.25) : 1.0;  // as choose_series
            const int t1 = taylor_terms(theta1 * inv_st), t2 = taylor_terms(d2 * inv_st) + 1;
            if (t2 < t1 && theta1 * inv_st <= (double)t2) alpha = kTermThr[t2] / inv_st;
        }
        if (live && c0 == 0) plan[(long long)b * P.nI + kk] = make_double2(alpha, theta1);
    }
}

}  // namespace

// shared memory the staged form needs; 0: no plan (too many drives or the matrices do not fit: the kernels fall back to ||dt G||_1)
size_t series_plan_smem(int n, int m) {
    if (m > 4 || (n & 1)) return 0;
    const size_t nmat = m + 1, bytes = sizeof(double) * (nmat + nmat * (nmat + 1) / 2) * n * n;
    return bytes <= 200 * 1024 ? bytes : 0;
}

template <int NM>
static bool launch_plan_nm(const DProb& P, int ii, const double* Z, double2* out, long long items, int sms, cudaStream_t st) {
    const DInt& I = P.in[ii];
    const size_t smem = series_plan_smem(I.n, I.m);
    static PerDeviceOnce configured;
    if (configured.first()) {
        if (cudaFuncSetAttribute(series_plan_kernel<NM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return false;
    }
    const int gs = (I.n <= 32 && 32 % I.n == 0) ? I.n : 32, per_warp = 32 / gs;
    const long long tasks = (items + per_warp - 1) / per_warp;
    const int threads = 512, wpc = threads / 32;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(2048 / threads, (220 * 1024) / (smem + 1024)));
    const int grid = (int)std::min<long long>((tasks + wpc - 1) / wpc, (long long)sms * per_sm);
    series_plan_kernel<NM><<<grid, threads, smem, st>>>(P, ii, Z, out, gs);
    return true;
}

bool launch_series_plan(const DProb& P, int ii, const double* Z, cudaStream_t st, long long* launches) {
    const DInt& I = P.in[ii];
    if (I.plan == nullptr || I.Spair == nullptr || I.G_stride != 0 || series_plan_smem(I.n, I.m) == 0) return false;
    const long long items = (long long)P.batch * (std::min(P.kc1, P.nI) - P.kc0);
    if (items <= 0) return true;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double2* out = const_cast<double2*>(I.plan);
    bool ok = false;
    switch (I.m) {
        case 0: ok = launch_plan_nm<1>(P, ii, Z, out, items, sms, st); break;
        case 1: ok = launch_plan_nm<2>(P, ii, Z, out, items, sms, st); break;
        case 2: ok = launch_plan_nm<3>(P, ii, Z, out, items, sms, st); break;
        case 3: ok = launch_plan_nm<4>(P, ii, Z, out, items, sms, st); break;
        case 4: ok = launch_plan_nm<5>(P, ii, Z, out, items, sms, st); break;
    }
    if (ok) ++*launches;
    return ok;
}
