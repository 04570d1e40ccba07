// Multi-GPU exchange over NVLink peer memory, fused with the kernels that produce / consume the data.
//
// No reference counterpart (the reference is single-GPU).  The sharded build needs two reductions per step: the
// bounding box (4 doubles) and the per-finest-cell sums (count, m, m x, m y).  Instead of two NCCL calls they are done
// with plain stores into the peers' buffers (cudaIpc-mapped, one process per GPU, every peer one NVSwitch hop away)
// plus release / acquire flags carrying the step's sequence number:
//
//   * bounds_kernel's last block stores its rank's min/max into every peer's slot, waits for the other ranks' slots
//     and finalises the root box — no extra launch, no host involvement (bounds_keys.cu, peer_comm.cuh);
//   * cell_partial_kernel stores the sums of this rank's NON-EMPTY finest cells into its slot of EVERY rank's inbox
//     ([source rank][4][cells]) and raises one flag per peer (tree_build.cu);
//   * tree_bottom_kernel waits for the peers' flags, adds the ranks' contributions IN RANK ORDER (deterministic,
//     identical on every rank) and zeroes what it consumed, so the inbox is clean for the next step.
//
// Round 1 did a dense reduce-scatter + all-gather (3 launches, 2 flag waits, 15 MB of remote stores per rank and
// step whatever the occupancy: 62-75 us at 8 GPUs).  A rank's Morton slice touches ~1/R of the occupied cells, so
// the sparse all-to-all moves ~5 MB with one launch and one wait.
//
// Buffer reuse across steps is safe without double buffering: a rank raises the bounding-box flag of step s + 1 only
// after its own level pass of step s has consumed (and zeroed) its inbox, and nobody pushes cell sums of step s + 1
// before it has seen every rank's box of step s + 1.  Every spin has a wall-clock timeout (default 4 s, env
// BH_PEER_TIMEOUT_MS) that raises a sticky error word instead of hanging the GPU (peer_comm.cuh: wait_flag).
#include "peer_comm.cuh"

namespace bh {

void peer_comm_layout(PeerComm& pc, int rank, int n_ranks, uint64_t ncells, size_t* total_bytes) {
    pc.rank = rank; pc.n_ranks = n_ranks; pc.ncells = ncells;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t at = off; off += (bytes + 255) & ~size_t(255); return at; };
    pc.off_bbox = take(sizeof(double) * 4 * kMaxPeers);
    pc.off_bbox_flag = take(sizeof(uint32_t) * kMaxPeers);
    pc.off_in_flag = take(sizeof(uint32_t) * kMaxPeers);
    pc.off_err = take(sizeof(uint32_t) * 4);     // [0] timeout flag, [1] step sequence number, [2], [3] tickets
    pc.off_inbox = take(sizeof(double) * 4 * ncells * n_ranks);
    *total_bytes = off;
}

}  // namespace bh
