// Peer-memory exchange helpers shared by bounds_keys.cu and peer_comm.cu (see peer_comm.cu).
#pragma once
#include "bh_internal.h"

namespace bh {

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Spin until *flag == seq.  A wall-clock timeout (PeerComm::timeout_ns, env BH_PEER_TIMEOUT_MS, default 4 s) raises
// the error word instead of hanging the GPU.  The error is STICKY: once set, later waits return at once (the
// sequence protocol is desynchronised for good), and every host entry point that synchronises returns BH_ERR_NCCL
// until the peers are re-attached (api.cu: peer_error).
__device__ __forceinline__ void wait_flag(const uint32_t* flag, uint32_t seq, uint32_t* err, unsigned long long timeout_ns) {
    if (*reinterpret_cast<volatile uint32_t*>(err)) return;
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys(flag) != seq) {
        if (global_timer_ns() - t0 > timeout_ns) { atomicExch(err, 1u); break; }
    }
}

// Called by ONE block (>= kMaxPeers threads) after its block-wide reduction; thread 0 holds the local
// values on entry and the global (all ranks) values on return.
__device__ __forceinline__ void peer_bbox_exchange(const PeerComm& pc, uint32_t seq, double& xmin, double& xmax,
                                                   double& ymin, double& ymax) {
    __shared__ double s_loc[4];
    const int tid = threadIdx.x;
    if (tid == 0) { s_loc[0] = xmin; s_loc[1] = xmax; s_loc[2] = ymin; s_loc[3] = ymax; }
    __syncthreads();
    if (tid < pc.n_ranks) {
        double* dst = reinterpret_cast<double*>(pc.peer_base[tid] + pc.off_bbox) + 4 * pc.rank;
        dst[0] = s_loc[0]; dst[1] = s_loc[1]; dst[2] = s_loc[2]; dst[3] = s_loc[3];
        __threadfence_system();
        st_release_sys(reinterpret_cast<uint32_t*>(pc.peer_base[tid] + pc.off_bbox_flag) + pc.rank, seq);
        uint8_t* own = pc.peer_base[pc.rank];
        wait_flag(reinterpret_cast<const uint32_t*>(own + pc.off_bbox_flag) + tid, seq,
                  reinterpret_cast<uint32_t*>(own + pc.off_err), pc.timeout_ns);
    }
    __syncthreads();
    if (tid == 0) {
        const volatile double* in = reinterpret_cast<const volatile double*>(pc.peer_base[pc.rank] + pc.off_bbox);
        double a = INFINITY, b = -INFINITY, c = INFINITY, d = -INFINITY;
        for (int r = 0; r < pc.n_ranks; ++r) {   // same comparison form as ComputeRootBounds (project.cu:547-550)
            const double x0 = in[4 * r], x1 = in[4 * r + 1], y0 = in[4 * r + 2], y1 = in[4 * r + 3];
            a = (x0 < a) ? x0 : a; b = (b < x1) ? x1 : b;
            c = (y0 < c) ? y0 : c; d = (d < y1) ? y1 : d;
        }
        xmin = a; xmax = b; ymin = c; ymax = d;
    }
}

}  // namespace bh
