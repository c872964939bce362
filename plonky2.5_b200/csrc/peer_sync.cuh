// Stream-ordered signals between ranks (one process per GPU) through CUDA-IPC peer memory — the only synchronisation the streamed
// coset-sharded commit (gl_commit.cu · gl_commit_coset_stream) needs between its waves.
//
// Every rank owns a flag array [n_peers][n_waves] at the end of its exported buffer.  After the iNTT of its wave-w column group a rank
// STORES a monotonically increasing ticket into slot [self][w] of every peer's array (one kernel, release stores over NVLink); a peer's
// pull stream polls ITS OWN copy of the slot (acquire loads from local memory, no NVLink traffic) before the copy engine fetches the
// block.  Tickets only grow (epoch * n_waves + w + 1), so a slot never has to be reset and a stale value can never satisfy a wait.
// The wait gives up after `timeout_ns` and raises an error word instead of spinning forever on a peer that died.
#pragma once
#include <stdint.h>
#include "ntt.cuh"

namespace peersync {

struct Slots {
    uint64_t* p[ntt::MAX_PEERS];
};

__global__ void signal_kernel(Slots s, uint32_t n, uint64_t ticket) {
    if (threadIdx.x < n && s.p[threadIdx.x]) {
        __threadfence_system();   // the preceding kernels of this stream have completed; this orders the store behind their writes for the peer
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(s.p[threadIdx.x]), "l"(ticket) : "memory");
    }
}

__global__ void wait_kernel(const uint64_t* flag, uint64_t ticket, uint32_t* err, uint64_t timeout_ns) {
    uint64_t t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        uint64_t v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
        if (v >= ticket) return;
        uint64_t t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > timeout_ns) {
            atomicExch(err, 1u);
            return;
        }
        __nanosleep(256);
    }
}

}  // namespace peersync
