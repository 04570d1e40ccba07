// Multi-GPU exchange over NVLink peer memory, fused with the kernels that produce / consume the data.
//
// No reference counterpart (the reference is single-GPU).  The sharded build needs two reductions
// per step: the bounding box (4 doubles) and the per-finest-cell sums (4 doubles per cell, 8.4 MB at
// the default cap).  Instead of two NCCL calls they are done with plain stores into the peers'
// buffers (cudaIpc-mapped, one process per GPU, every peer one NVSwitch hop away) plus release /
// acquire flags carrying the step's sequence number:
//
//   * bounds_kernel's last block stores its rank's min/max into every peer's slot, waits for the
//     other ranks' slots and finalises the root box — no extra launch, no host involvement;
//   * rs_push_kernel pushes each 1/R slice of this rank's partial cell sums to the slice's owner
//     (reduce-scatter by direct stores), rs_reduce_ag_kernel waits for the R contributions, adds
//     them IN RANK ORDER (deterministic, identical on every rank) and stores the reduced slice into
//     every peer's final array (all-gather by direct stores);
//   * wait_ag_kernel lets the level pass start once all reduced slices have landed.
//
// Buffer reuse across steps is safe without double buffering: a rank can only run ahead of a peer
// by the distance the data dependencies allow (it needs the peer's contribution of step s to finish
// step s), see DESIGN.md §7.  Every spin has a wall-clock timeout (default 4 s, env BH_PEER_TIMEOUT_MS) that
// raises a sticky error word instead of hanging the GPU (peer_comm.cuh: wait_flag).
#include "peer_comm.cuh"

namespace bh {

namespace {

// reduce-scatter by direct stores: element (k, c) of this rank's partial sums goes to the owner of cell c
__global__ void __launch_bounds__(256)
rs_push_kernel(PeerComm pc, const uint32_t* __restrict__ seq_dev, const double* __restrict__ local_sums,
               uint32_t* __restrict__ ticket) {
    const uint64_t total = 4 * pc.ncells;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t k = e / pc.ncells, c = e - k * pc.ncells;
        const uint64_t o = c / pc.slice, j = c - o * pc.slice;
        double* dst = reinterpret_cast<double*>(pc.peer_base[o] + pc.off_rs) + ((uint64_t)pc.rank * 4 + k) * pc.slice + j;
        *dst = local_sums[e];
    }
    __shared__ bool last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0) *ticket = 0;
        const uint32_t seq = *seq_dev;
        __threadfence_system();
        if ((int)threadIdx.x < pc.n_ranks)
            st_release_sys(reinterpret_cast<uint32_t*>(pc.peer_base[threadIdx.x] + pc.off_rs_flag) + pc.rank, seq);
    }
}

// wait for all contributions to my slice, add them in rank order, all-gather the result by direct stores
__global__ void __launch_bounds__(256)
rs_reduce_ag_kernel(PeerComm pc, const uint32_t* __restrict__ seq_dev, uint32_t* __restrict__ ticket) {
    uint8_t* own = pc.peer_base[pc.rank];
    const uint32_t seq = *seq_dev;
    if ((int)threadIdx.x < pc.n_ranks)
        wait_flag(reinterpret_cast<const uint32_t*>(own + pc.off_rs_flag) + threadIdx.x, seq,
                  reinterpret_cast<uint32_t*>(own + pc.off_err), pc.timeout_ns);
    __syncthreads();
    const uint64_t c0 = (uint64_t)pc.rank * pc.slice;
    const uint64_t len = c0 >= pc.ncells ? 0 : (pc.ncells - c0 < pc.slice ? pc.ncells - c0 : pc.slice);
    const double* rs = reinterpret_cast<const double*>(own + pc.off_rs);
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < 4 * len; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t k = e / len, j = e - k * len;
        double v = 0.0;
        for (int r = 0; r < pc.n_ranks; ++r) v += rs[((uint64_t)r * 4 + k) * pc.slice + j];   // fixed rank order
        for (int r = 0; r < pc.n_ranks; ++r)
            reinterpret_cast<double*>(pc.peer_base[r] + pc.off_sums)[k * pc.ncells + c0 + j] = v;
    }
    __shared__ bool last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0) *ticket = 0;
        __threadfence_system();
        if ((int)threadIdx.x < pc.n_ranks)
            st_release_sys(reinterpret_cast<uint32_t*>(pc.peer_base[threadIdx.x] + pc.off_ag_flag) + pc.rank, seq);
    }
}

__global__ void wait_ag_kernel(PeerComm pc, const uint32_t* __restrict__ seq_dev) {
    uint8_t* own = pc.peer_base[pc.rank];
    if ((int)threadIdx.x < pc.n_ranks)
        wait_flag(reinterpret_cast<const uint32_t*>(own + pc.off_ag_flag) + threadIdx.x, *seq_dev,
                  reinterpret_cast<uint32_t*>(own + pc.off_err), pc.timeout_ns);
}

}  // namespace

void peer_comm_layout(PeerComm& pc, int rank, int n_ranks, uint64_t ncells, size_t* total_bytes) {
    pc.rank = rank; pc.n_ranks = n_ranks; pc.ncells = ncells;
    pc.slice = (ncells + n_ranks - 1) / n_ranks;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t at = off; off += (bytes + 255) & ~size_t(255); return at; };
    pc.off_bbox = take(sizeof(double) * 4 * kMaxPeers);
    pc.off_bbox_flag = take(sizeof(uint32_t) * kMaxPeers);
    pc.off_rs_flag = take(sizeof(uint32_t) * kMaxPeers);
    pc.off_ag_flag = take(sizeof(uint32_t) * kMaxPeers);
    pc.off_err = take(sizeof(uint32_t) * 4);     // [0] timeout flag, [1] step sequence number, [2], [3] tickets
    pc.off_rs = take(sizeof(double) * 4 * pc.slice * n_ranks);
    pc.off_sums = take(sizeof(double) * 4 * ncells);
    *total_bytes = off;
}

// local_sums -> (peer exchange) -> final sums in this rank's comm buffer (pc.off_sums)
void launch_peer_allreduce_cells(const PeerComm& pc, const double* local_sums, cudaStream_t st) {
    uint8_t* own = pc.peer_base[pc.rank];
    uint32_t* seq_dev = reinterpret_cast<uint32_t*>(own + pc.off_err) + 1;
    uint32_t* ticket = reinterpret_cast<uint32_t*>(own + pc.off_err) + 2;
    rs_push_kernel<<<148 * 2, 256, 0, st>>>(pc, seq_dev, local_sums, ticket);
    rs_reduce_ag_kernel<<<148, 256, 0, st>>>(pc, seq_dev, ticket + 1);
    wait_ag_kernel<<<1, 32, 0, st>>>(pc, seq_dev);
    g_launches += 3;
}

}  // namespace bh
