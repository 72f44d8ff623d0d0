// Knot-range shards of one long trajectory (partitioning B): the cross-rank traffic of an
// evaluation, all of it through peer memory (NVLink P2P stores into windows mapped with CUDA IPC), no collective library
// on the path:
//   * the one-knot halo: after uploading iterate e a shard PUSHES its first knot into the left neighbour's exchange
//     window (slot e & 1) and then publishes e there; the left neighbour's first kernel of iterate e waits for that flag.
//     Flow control is an acknowledgement the reader writes back into the pusher's window when it moves on to its next
//     iterate, so a fast rank can run at most one iterate ahead of the neighbour that still reads its knot.
//   * the two scalars every rank needs (objective: sum, constraint violation: max): one kernel writes {J, viol, seq} into
//     this rank's slot of EVERY peer's window and spins until all slots of its own window carry seq; the sum runs in
//     rank order on every rank (bit-identical results everywhere).
// A wait that sees no progress for kTimeoutNs sets XWin::err and gives up (reported by the host as DTO_ERR_CUDA).
#include "dto_internal.h"

namespace {

constexpr unsigned long long kTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// spin until *flag >= want; false on timeout
__device__ __forceinline__ bool wait_geq(const volatile unsigned long long* flag, unsigned long long want) {
    const unsigned long long t0 = now_ns();
    while (*flag < want) {
        if (now_ns() - t0 > kTimeoutNs) return false;
        __nanosleep(200);
    }
    return true;
}

__global__ void shard_publish_kernel(XWin* own, XWin* left, XWin* right, const double* __restrict__ first_knot, int z,
                                     unsigned long long epoch) {
    const int lane = threadIdx.x;
    if (right != nullptr && lane == 0) {  // everything of iterate epoch - 1 on this stream is done: its halo slot may be reused
        *(volatile unsigned long long*)&right->ack = epoch - 1;
        __threadfence_system();
    }
    if (left != nullptr) {
        bool ok = true;
        if (lane == 0 && epoch >= 2) ok = wait_geq(&own->ack, epoch - 2);  // the reader left iterate epoch - 2 (same slot) behind
        ok = __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;
        if (!ok) {
            if (lane == 0) own->err = 1;
            return;
        }
        double* slot = left->halo + (size_t)(epoch & 1) * z;
        for (int i = lane; i < z; i += 32) slot[i] = first_knot[i];
        __threadfence_system();
        __syncwarp();
        if (lane == 0) {
            *(volatile unsigned long long*)&left->halo_epoch = epoch;
            __threadfence_system();
        }
    }
}

__global__ void shard_wait_kernel(XWin* own, unsigned long long epoch) {
    if (!wait_geq(&own->halo_epoch, epoch)) own->err = 2;
    __threadfence_system();
}

__global__ void scalar_exchange_kernel(XWin* own, XWin* const* __restrict__ peers, int rank, int world, unsigned long long seq,
                                       double* __restrict__ J, double* __restrict__ viol) {
    const int lane = threadIdx.x;
    const int par = (int)(seq & 1);
    const double myJ = *J, myV = *viol;
    if (lane < world) {
        volatile double* s = peers[lane]->scal[par][rank];
        s[0] = myJ;
        s[1] = myV;
        __threadfence_system();
        *(volatile unsigned long long*)&s[2] = seq;  // the tag goes last
        __threadfence_system();
    }
    bool ok = true;
    if (lane < world) ok = wait_geq((const volatile unsigned long long*)&own->scal[par][lane][2], seq);
    ok = __all_sync(0xffffffffu, ok);
    if (!ok) {
        if (lane == 0) own->err = 3;
        return;
    }
    __threadfence_system();
    if (lane == 0) {
        double sum = 0.0, mx = 0.0;
        for (int r = 0; r < world; ++r) {  // rank order on every rank: identical bits everywhere
            sum += ((volatile double*)own->scal[par][r])[0];
            mx = fmax(mx, ((volatile double*)own->scal[par][r])[1]);
        }
        *J = sum;
        *viol = mx;
    }
}

}  // namespace

void launch_shard_publish(XWin* own, XWin* left, XWin* right, const double* first_knot, int z, unsigned long long epoch, cudaStream_t st,
                          long long* launches) {
    shard_publish_kernel<<<1, 32, 0, st>>>(own, left, right, first_knot, z, epoch);
    ++*launches;
}

void launch_shard_wait(XWin* own, unsigned long long epoch, cudaStream_t st, long long* launches) {
    shard_wait_kernel<<<1, 1, 0, st>>>(own, epoch);
    ++*launches;
}

void launch_scalar_exchange(XWin* own, XWin* const* peers, int rank, int world, unsigned long long seq, double* J, double* viol,
                            cudaStream_t st, long long* launches) {
    scalar_exchange_kernel<<<1, 32, 0, st>>>(own, peers, rank, world, seq, J, viol);
    ++*launches;
}
