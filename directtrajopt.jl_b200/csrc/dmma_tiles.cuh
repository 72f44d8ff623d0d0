// Warp-level FP64 tensor-core tiles (mma.sync m8n8k4 -> SASS DMMA.8x8x4) shared by the interval kernels.
//
// Vectors are the ROWS of the A operand: out[vec][s] += sum_k V[vec][k] * M[s][k].  A lane holds row lane/4
// and columns 2*(lane%4)+{0,1} of each 8-wide tile, so the C fragment of one product is the A fragment of the
// next when k-step 2t contracts over the even and k-step 2t+1 over the odd states of tile t; the matching B
// fragments are then one 16-byte load from the row-major matrix.  Matrices are n x n, n = 8*NT, unpadded;
// for even NT the 8-double blocks of odd rows are swapped pairwise (XOR swizzle) so the 8 rows x 64 bytes of
// one LDS.128 phase fall into distinct banks.  The same physical layout is used in shared and global memory.
#pragma once

namespace dmma_tiles {

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ double quad_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

// offset of element (row s, column k)
template <int NT>
__host__ __device__ __forceinline__ int sw(int s, int k) {
    if constexpr (NT % 2 == 0) return s * (8 * NT) + (k ^ ((s & 1) << 3));
    else return s * (8 * NT) + k;
}

// out[mt][nt] += V[mt] * M'   (MT = tiles used, MD = tiles the source array is declared with)
template <int MT, int NT, int MD>
__device__ __forceinline__ void mma_apply(double (&out)[MT][NT][2], const double (&v)[MD][NT][2], const double* __restrict__ M,
                                          int lane) {
    constexpr int n = 8 * NT;
    const int row8 = lane >> 2;
    const int d = (NT % 2 == 0) ? ((row8 & 1) << 3) : 0;
    const double* base = M + row8 * n + 2 * (lane & 3);
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        const int off = 8 * t + ((t & 1) ? -d : d);  // 8 * (t ^ (row & 1))
        double2 b[NT];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) b[nt] = *reinterpret_cast<const double2*>(base + 8 * nt * n + off);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) dmma(out[mt][nt][0], out[mt][nt][1], v[mt][t][0], b[nt].x);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) dmma(out[mt][nt][0], out[mt][nt][1], v[mt][t][1], b[nt].y);
    }
}

template <int MT, int NT>
__device__ __forceinline__ void frag_zero(double (&f)[MT][NT][2]) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) f[mt][nt][0] = f[mt][nt][1] = 0.0;
}

}  // namespace dmma_tiles
