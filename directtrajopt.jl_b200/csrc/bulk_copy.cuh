// Staging a read-only table from global into shared memory with the bulk-copy engine (TMA, cp.async.bulk -> SASS UBLKCP):
// one thread issues 16 KB pieces that complete on an mbarrier, every thread of the CTA waits on it.  Per-thread loads make
// such a prologue a chain of L2 latencies (every CTA of the grid reads the same lines at the same time).
#pragma once

// all threads of the CTA call this; src and dst 16-byte aligned, bytes a multiple of 16 and < 2^20; `bar` is an 8-byte
// aligned word in shared memory used ONCE per kernel (phase 0)
__device__ __forceinline__ void bulk_stage(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    const unsigned bar_a = (unsigned)__cvta_generic_to_shared(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        const char* src = reinterpret_cast<const char*>(src_gmem);
        const unsigned dst = (unsigned)__cvta_generic_to_shared(dst_smem);
        for (unsigned off = 0; off < bytes; off += 16384u) {
            const unsigned sz = min(16384u, bytes - off);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + off),
                         "l"(src + off), "r"(sz), "r"(bar_a)
                         : "memory");
        }
    }
    unsigned done = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(bar_a)
                     : "memory");
    }
}
