// Cross-GPU hand-over flags for the exact multi-GPU commitment (sharding.py): the encoder of rank g stores column
// slab h of its codeword tile straight into rank h's memory over NVLink (CodewordSink, kernels.h); these two
// kernels are the only synchronisation -- no collective on the data path.
//   signal: after the encode kernel (stream order => its peer stores are complete), publish `value` into one u64
//           slot per peer (system-scope release)
//   wait  : spin (system-scope acquire) until every slot of the local flag array has reached `value`; bounded by a
//           timeout so that a lost peer can never hang the GPU: on expiry *err is set and the kernel returns
// Flags are monotone counters (round numbers); nobody ever resets them while a commitment is in flight.
#include "kernels.h"

namespace lgr {

__global__ void peer_signal_kernel(PeerSlots slots, unsigned long long value) {
    const int i = threadIdx.x;
    if (i < slots.n) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(slots.p[i]), "l"(value) : "memory");
    }
}

__global__ void peer_wait_kernel(const unsigned long long *flags, int n, unsigned long long value, unsigned long long timeout_ns,
                                 unsigned int *err) {
    const int i = threadIdx.x;
    if (i < n) {
        unsigned long long t0, t1, v;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + i) : "memory");
            if (v >= value) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) { atomicExch(err, 1u + (unsigned)i); break; }
            __nanosleep(200);
        }
    }
    __syncthreads();
    __threadfence_system();
}

cudaError_t launch_peer_signal(const PeerSlots &slots, unsigned long long value, cudaStream_t st) {
    if (slots.n <= 0) return cudaSuccess;
    peer_signal_kernel<<<1, 32, 0, st>>>(slots, value);
    return cudaGetLastError();
}
cudaError_t launch_peer_wait(const unsigned long long *flags, int n, unsigned long long value, unsigned long long timeout_ns,
                             unsigned int *err, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    peer_wait_kernel<<<1, 32, 0, st>>>(flags, n, value, timeout_ns, err);
    return cudaGetLastError();
}

}  // namespace lgr
