// Host-link probe for the host-delivered path: what the box's PCIe complex gives
//   (a) the copy engine: one large cudaMemcpyAsync D2H, and the 2-D (strided) form with 256-byte rows
//   (b) SM-issued stores into MAPPED page-locked host memory (zero-copy stream-out kernel): contiguous and
//       256-of-512-byte strided, for several grid sizes
//   (c) host memset / memcpy rates of one thread (what the zero-fill threads can do)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/pcie_probe.cu -o build/pcie_probe
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

// copy `rows` pieces of `piece` doubles; source/destination row pitch in doubles; 16-byte stores
__global__ void stream_out(double2* __restrict__ dst, const double2* __restrict__ src, long long rows, int piece2, int spitch2, int dpitch2) {
    const long long total = rows * piece2;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / piece2;
        const int c = (int)(e % piece2);
        dst[r * dpitch2 + c] = src[r * spitch2 + c];
    }
}

static double time_ms(cudaStream_t st, int reps, void (*fn)(void*), void* ctx) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    fn(ctx);
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(a, st));
    for (int i = 0; i < reps; ++i) fn(ctx);
    CK(cudaEventRecord(b, st));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}

struct Ctx {
    cudaStream_t st;
    double *d, *h, *hdev;
    size_t bytes;
    long long rows;
    int grid, piece2, spitch2, dpitch2;
};

int main(int argc, char** argv) {
    const size_t MB = 1 << 20;
    const size_t bytes = (argc > 1 ? atoi(argv[1]) : 38) * MB;
    Ctx c{};
    CK(cudaStreamCreate(&c.st));
    c.bytes = bytes;
    CK(cudaMalloc(&c.d, 2 * bytes));
    CK(cudaMemset(c.d, 1, 2 * bytes));
    CK(cudaHostAlloc(&c.h, 2 * bytes, cudaHostAllocMapped | cudaHostAllocPortable));
    CK(cudaHostGetDevicePointer(&c.hdev, c.h, 0));
    memset(c.h, 0, 2 * bytes);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, buffer %zu MB\n", prop.name, bytes / MB);

    double ms = time_ms(c.st, 10, [](void* p) { Ctx* c = (Ctx*)p; CK(cudaMemcpyAsync(c->h, c->d, c->bytes, cudaMemcpyDeviceToHost, c->st)); }, &c);
    printf("copy engine D2H contiguous            : %.3f ms  %.1f GB/s\n", ms, bytes / ms * 1e-6);
    ms = time_ms(c.st, 10, [](void* p) { Ctx* c = (Ctx*)p; CK(cudaMemcpyAsync(c->d, c->h, c->bytes, cudaMemcpyHostToDevice, c->st)); }, &c);
    printf("copy engine H2D contiguous            : %.3f ms  %.1f GB/s\n", ms, bytes / ms * 1e-6);
    ms = time_ms(c.st, 10, [](void* p) { Ctx* c = (Ctx*)p; CK(cudaMemcpy2DAsync(c->h, 512, c->d, 512, 256, c->bytes / 256, cudaMemcpyDeviceToHost, c->st)); }, &c);
    printf("copy engine D2H 256 of 512 B rows     : %.3f ms  %.1f GB/s (payload)\n", ms, bytes / ms * 1e-6);
    ms = time_ms(c.st, 10, [](void* p) { Ctx* c = (Ctx*)p; CK(cudaMemcpy2DAsync(c->h, 8 * 1369, c->d, 8 * 1369, 8 * 222, 2 * c->bytes / (8 * 1369), cudaMemcpyDeviceToHost, c->st)); }, &c);
    printf("copy engine D2H 1776 of 10952 B rows  : %.3f ms  %.1f GB/s (payload)\n", ms, (double)(2 * bytes / (8 * 1369)) * 8 * 222 / ms * 1e-6);

    const int grids[] = {2, 4, 8, 16, 32, 74, 148, 296};
    for (int strided = 0; strided < 2; ++strided)
        for (int g : grids) {
            c.grid = g;
            c.piece2 = 16;  // 256 bytes
            c.spitch2 = strided ? 32 : 16;
            c.dpitch2 = strided ? 32 : 16;
            c.rows = (long long)(bytes / 256);
            ms = time_ms(c.st, 5, [](void* p) {
                Ctx* c = (Ctx*)p;
                stream_out<<<c->grid, 1024, 0, c->st>>>((double2*)c->hdev, (const double2*)c->d, c->rows, c->piece2, c->spitch2, c->dpitch2);
            }, &c);
            printf("SM stream-out to mapped host, %s, grid %3d x 1024 : %.3f ms  %.1f GB/s (payload)\n",
                   strided ? "256 of 512 B" : "contiguous  ", g, ms, bytes / ms * 1e-6);
        }
    // zero-copy reads (H2D by the SMs)
    for (int g : {8, 32, 148}) {
        c.grid = g;
        c.piece2 = 16; c.spitch2 = 16; c.dpitch2 = 16;
        c.rows = (long long)(bytes / 256);
        ms = time_ms(c.st, 5, [](void* p) {
            Ctx* c = (Ctx*)p;
            stream_out<<<c->grid, 1024, 0, c->st>>>((double2*)c->d, (const double2*)c->hdev, c->rows, c->piece2, c->spitch2, c->dpitch2);
        }, &c);
        printf("SM read from mapped host, grid %3d x 1024 : %.3f ms  %.1f GB/s\n", g, ms, bytes / ms * 1e-6);
    }
    // cudaHostRegister of ordinary memory (what a Julia Vector{Float64} is): cost and the same stream-out
    {
        double* plain = (double*)aligned_alloc(4096, bytes);
        memset(plain, 0, bytes);
        auto t0 = std::chrono::steady_clock::now();
        CK(cudaHostRegister(plain, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
        auto t1 = std::chrono::steady_clock::now();
        printf("cudaHostRegister(%zu MB, mapped)      : %.3f ms\n", bytes / MB, std::chrono::duration<double, std::milli>(t1 - t0).count());
        double* pdev = nullptr;
        CK(cudaHostGetDevicePointer(&pdev, plain, 0));
        Ctx c2 = c;
        c2.hdev = pdev;
        c2.grid = 32;
        ms = time_ms(c.st, 5, [](void* p) {
            Ctx* c = (Ctx*)p;
            stream_out<<<c->grid, 1024, 0, c->st>>>((double2*)c->hdev, (const double2*)c->d, c->rows, 16, 16, 16);
        }, &c2);
        printf("SM stream-out to REGISTERED memory, grid 32      : %.3f ms  %.1f GB/s\n", ms, bytes / ms * 1e-6);
        c2.h = plain;
        ms = time_ms(c.st, 10, [](void* p) { Ctx* c = (Ctx*)p; CK(cudaMemcpyAsync(c->h, c->d, c->bytes, cudaMemcpyDeviceToHost, c->st)); }, &c2);
        printf("copy engine D2H to REGISTERED memory             : %.3f ms  %.1f GB/s\n", ms, bytes / ms * 1e-6);
        CK(cudaHostUnregister(plain));
        free(plain);
    }
    {
        double* plain = (double*)malloc(bytes);
        memset(plain, 1, bytes);
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < 5; ++i) memset(plain, 0, bytes);
        auto t1 = std::chrono::steady_clock::now();
        printf("host memset, one thread               : %.1f GB/s\n", 5.0 * bytes / std::chrono::duration<double>(t1 - t0).count() * 1e-9);
        t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < 5; ++i) memcpy(plain, c.h, bytes);
        t1 = std::chrono::steady_clock::now();
        printf("host memcpy pinned -> pageable, one thread : %.1f GB/s\n", 5.0 * bytes / std::chrono::duration<double>(t1 - t0).count() * 1e-9);
        t0 = std::chrono::steady_clock::now();
        volatile int r = 0;
        for (int i = 0; i < 5; ++i) r += memcmp(plain, c.h, bytes);
        t1 = std::chrono::steady_clock::now();
        printf("host memcmp (equal buffers), one thread    : %.1f GB/s\n", 5.0 * bytes / std::chrono::duration<double>(t1 - t0).count() * 1e-9);
        // pageable D2H (what dto_eval_jacobian does with a plain Julia array)
        ms = 0;
        t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < 5; ++i) CK(cudaMemcpy(plain, c.d, bytes, cudaMemcpyDeviceToHost));
        t1 = std::chrono::steady_clock::now();
        printf("cudaMemcpy D2H to PAGEABLE memory     : %.1f GB/s\n", 5.0 * bytes / std::chrono::duration<double>(t1 - t0).count() * 1e-9);
        free(plain);
    }
    return 0;
}
