// Cross-GPU flag barrier over peer-mapped memory (one process per GPU, NVLink): the step that
// closes a replicated-result launch (mbqc_psr_grad_batch_push) -- every rank tells every other rank
// "my rows are stored" and waits for the same from all of them, on the stream, without a host
// round trip and without a collective library call.
#include "host_util.h"

namespace {

struct BarrierArgs {
    unsigned long long* flags[8];  // flags[d]: the flag array of rank d (peer-mapped; [16] words each)
    int n, me;
};

// The epoch is word 8 of the rank's own flag array, advanced by the kernel itself: the launch has no
// changing argument, so it can sit in a CUDA graph.  thread d: release-store the epoch into slot
// `me` of rank d's flags, then wait for slot d of our own.  The kernel before this one on the stream
// has completed, so its peer stores are performed; the system-scope release / acquire pair orders
// them before anything the waiting rank runs next.
__global__ void peer_barrier_kernel(const __grid_constant__ BarrierArgs a) {
    __shared__ unsigned long long s_epoch;
    const int d = threadIdx.x;
    if (d == 0) {
        unsigned long long* counter = a.flags[a.me] + 8;
        s_epoch = *counter + 1ull;
        *counter = s_epoch;
    }
    __syncthreads();
    if (d >= a.n) return;
    const unsigned long long epoch = s_epoch;
    unsigned long long* theirs = a.flags[d] + a.me;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(epoch) : "memory");
    const unsigned long long* mine = a.flags[a.me] + d;
    unsigned long long v;
    const long long t0 = clock64();
    do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
        // a rank that never arrives (crashed process, mismatched call sequence) must not park the GPU for
        // ever: after ~30 s of SM clocks the kernel faults, which the next CUDA call reports
        if (v < epoch && clock64() - t0 > 60000000000ll) __trap();
    } while (v < epoch);
}

}  // namespace

extern "C" {

// d_flags[r]: rank r's flag array (16 x uint64, zero-initialised once, in IPC-shared memory; entry
// `rank` is this process's own allocation; words 0..7 arrival flags, word 8 the rank's call count).
// Every rank must make the same sequence of calls.
int mbqc_peer_barrier(void* const* d_flags, int32_t n_ranks, int32_t rank, void* stream) {
    if (!d_flags || n_ranks < 1 || n_ranks > 8 || rank < 0 || rank >= n_ranks) return mbqc_set_error(MBQC_E_ARG, "bad barrier arguments");
    BarrierArgs a;
    for (int d = 0; d < 8; ++d) {
        a.flags[d] = d < n_ranks ? (unsigned long long*)d_flags[d] : nullptr;
        if (d < n_ranks && !d_flags[d]) return mbqc_set_error(MBQC_E_ARG, "NULL flag array");
    }
    a.n = n_ranks;
    a.me = rank;
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a);
    return mbqc_after_launch("peer_barrier_kernel");
}

}  // extern "C"
