// FP64 peak microbenchmarks for B200 (sm_100a): DFMA saturation, DMMA (mma.sync f64) throughput
// and dependent-issue latency.  Used to establish the FP64 roofline denominator that
// MEASURED_PEAKS.json does not carry.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

template <int NACC>
__global__ void dmma884_kernel(double* out, int iters) {
    double c[NACC][2];
    double a = threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-4;
#pragma unroll
    for (int j = 0; j < NACC; ++j) { c[j][0] = j; c[j][1] = -j; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dmma1688_kernel(double* out, int iters) {
    double c[NACC][4];
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 1.0 - threadIdx.x * 1e-4, b1 = b0 * 0.5;
#pragma unroll
    for (int j = 0; j < NACC; ++j) { c[j][0] = j; c[j][1] = -j; c[j][2] = 2 * j; c[j][3] = 3 * j; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) {
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3])
                         : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dmma16816_kernel(double* out, int iters) {
    double c[NACC][4];
    double a[8], b[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = threadIdx.x * 1e-3 + j;
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = 1.0 - threadIdx.x * 1e-4 * (j + 1);
#pragma unroll
    for (int j = 0; j < NACC; ++j) { c[j][0] = j; c[j][1] = -j; c[j][2] = 2 * j; c[j][3] = 3 * j; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) {
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                         : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: DMMA fed by shared-memory B fragments (one LDS.64 per mma), A from registers
template <int NACC>
__global__ void dmma884_lds_kernel(double* out, int iters) {
    __shared__ double sm[32 * 32];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * 1e-3;
    __syncthreads();
    double c[NACC][2];
    double a = threadIdx.x * 1e-3;
    int lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < NACC; ++j) { c[j][0] = j; c[j][1] = -j; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) {
            double b = sm[((i + j) & 31) * 32 + lane];
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed issue: every warp interleaves NACC DMMA.8x8x4 with NF independent DFMA per iteration.  If the time equals
// max(DMMA-only, DFMA-only) the two are separate datapaths; if it equals the sum they share one.
template <int NACC, int NF>
__global__ void mixed_kernel(double* out, int iters, double fa, double fb) {
    double c[NACC > 0 ? NACC : 1][2];
    double x[NF > 0 ? NF : 1];
    double a = threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-4;
#pragma unroll
    for (int j = 0; j < NACC; ++j) { c[j][0] = j; c[j][1] = -j; }
#pragma unroll
    for (int j = 0; j < NF; ++j) x[j] = threadIdx.x + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < (NACC > NF ? NACC : NF); ++j) {
            if (j < NACC)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
            if (j < NF) asm volatile("fma.rn.f64 %0, %0, %1, %2;\n" : "+d"(x[j]) : "d"(fa), "d"(fb));
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += c[j][0] + c[j][1];
#pragma unroll
    for (int j = 0; j < NF; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s  SMs %d  clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    int nsm = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * nsm * 16 * 1024));
    const int iters = 20000;
    for (int bps = 1; bps <= 8; bps *= 2) {
        int blocks = nsm * bps, threads = 256;
        float ms = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
        double flops = 2.0 * 8 * iters * (double)blocks * threads;
        printf("DFMA    blocks/SM %d x256thr : %.3f ms  %.2f TFLOP/s\n", bps, ms, flops / ms * 1e-9);
    }
    for (int wps = 4; wps <= 32; wps *= 2) {
        int blocks = nsm, threads = wps * 32;
        {
            float ms = time_ms([&] { dmma884_kernel<8><<<blocks, threads>>>(out, iters); }, 5);
            double flops = 2.0 * 8 * 8 * 4 * 8.0 * iters * (double)blocks * wps;
            printf("DMMA m8n8k4   warps/SM %2d acc8 : %.3f ms  %.2f TFLOP/s\n", wps, ms, flops / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { dmma1688_kernel<4><<<blocks, threads>>>(out, iters); }, 5);
            double flops = 2.0 * 16 * 8 * 8 * 4.0 * iters * (double)blocks * wps;
            printf("DMMA m16n8k8  warps/SM %2d acc4 : %.3f ms  %.2f TFLOP/s\n", wps, ms, flops / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { dmma16816_kernel<4><<<blocks, threads>>>(out, iters); }, 5);
            double flops = 2.0 * 16 * 8 * 16 * 4.0 * iters * (double)blocks * wps;
            printf("DMMA m16n8k16 warps/SM %2d acc4 : %.3f ms  %.2f TFLOP/s\n", wps, ms, flops / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { dmma884_lds_kernel<8><<<blocks, threads>>>(out, iters); }, 5);
            double flops = 2.0 * 8 * 8 * 4 * 8.0 * iters * (double)blocks * wps;
            printf("DMMA m8n8k4+LDS warps/SM %2d acc8 : %.3f ms  %.2f TFLOP/s\n", wps, ms, flops / ms * 1e-9);
        }
    }
    // are DFMA and DMMA one datapath or two?  (8 warps/SM; per iteration 8 DMMA = 4096 flop/warp, NF DFMA = 64*NF flop/warp)
    {
        int blocks = nsm, threads = 8 * 32;
        float m0 = time_ms([&] { mixed_kernel<8, 0><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
        float f16 = time_ms([&] { mixed_kernel<0, 16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
        float f64 = time_ms([&] { mixed_kernel<0, 64><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
        float x16 = time_ms([&] { mixed_kernel<8, 16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
        float x64 = time_ms([&] { mixed_kernel<8, 64><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf("MIXED 8 warps/SM: 8 DMMA only %.3f ms | 16 DFMA only %.3f ms | 64 DFMA only %.3f ms | 8 DMMA + 16 DFMA %.3f ms | 8 DMMA + 64 DFMA %.3f ms\n",
               m0, f16, f64, x16, x64);
        double fl = (4096.0 + 64.0 * 64) * iters * (double)blocks * 8;
        printf("MIXED 8 DMMA + 64 DFMA combined rate: %.2f TFLOP/s\n", fl / x64 * 1e-9);
    }
    // dependent-chain latency: one warp per SM, single accumulator
    {
        float ms = time_ms([&] { dmma884_kernel<1><<<1, 32>>>(out, iters); }, 5);
        printf("DMMA m8n8k4 dependent chain: %.1f ns per mma (%.1f cycles @ %.0f MHz nominal)\n",
               ms * 1e6 / iters, ms * 1e6 / iters * p.clockRate * 1e-6, p.clockRate * 1e-3);
        float ms2 = time_ms([&] { dfma_kernel<<<1, 32>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf("DFMA 8-chain single warp: %.2f ns per 8 fma\n", ms2 * 1e6 / iters);
    }
    CK(cudaFree(out));
    return 0;
}
