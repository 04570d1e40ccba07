// Internal declarations shared by the CUDA translation units of libbh.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <utility>

#include "../../include/bh.h"

namespace bh {

constexpr int kMaxLevels = 16;       // levels 0..max_depth-1 (max_depth <= 13 in the dense pyramid)
constexpr int kMaxDepthDense = 13;   // 4^12 finest cells * 32 B = 512 MiB of records
constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;       // keys per thread and tile
constexpr int kSortItemsSmall = 8;   // A/B (round 2): 489 tiles of 2048 keys at 1M bodies instead of 245 of 4096 — measured
constexpr int64_t kSortSmallN = 0;   // SLOWER (47.1 vs 39.6 us for the two passes: twice the look-back), so never selected
constexpr int kMaxSortPasses = 4;
constexpr int kMaxBins = 512;
constexpr uint32_t kHugeCellMin = 8192;    // finest cells above this many bodies are summed by kHugeParts blocks
constexpr int kHugeParts = 64;
constexpr int64_t kTwoBodiesPerLaneMin = 500000;   // traversal: 2 bodies per lane from this many bodies on (bh_params.reserved[0] overrides)

// node flags (packed FP32 traversal record and FP64 verification path share them)
constexpr uint32_t kNodeNonZero = 1u;  // mass > mass_eps         (project.cu:617 / :731)
constexpr uint32_t kNodeLeaf = 2u;     // all four children == -1 (project.cu:623-626)
constexpr uint32_t kNodeSingle = 4u;   // exactly one body inside (PARTICLE_INDEX == idx or -idx-2)

// 32-byte traversal record; the four children of a cell are one 128-byte line.
// Pyramid cells are indexed heap-style: root 0, children of cell p are 4p+1 .. 4p+4 (this equals
// level_off[l] + Morton code).  `rec` is allocated with a 96-byte lead-in so that every sibling
// group starts on a 128-byte line.
struct __align__(32) NodeRec {
    float chx, chy;   // scaled COM, high float of the double-float pair
    float clx, cly;   // scaled COM, low float   (com = ch + cl to ~2^-48)
    float gm;         // (float)(G * mass * scale^2); 0 for nodes the reference skips (mass <= mass_eps)
    float thr;        // lane accepts iff !(d2 <= thr): level threshold, or -1 for leaves / skipped nodes
    uint32_t count;   // bodies in the cell
    uint32_t first;   // sorted position of the cell's first body
};

// Written by the bounds kernel on the device every step; read by every later kernel.
struct StepConsts {
    double xmin, xmax, ymin, ymax;     // padded root box (project.cu:553-572)
    double size[kMaxLevels];           // max(width, height) of a level-l cell (project.cu:637-639)
    float thr2[kMaxLevels];            // FP32 mode: accept iff (S d)^2 > thr2[l]  (== size/(d+eps) < theta)
    double thr[kMaxLevels];            // size/theta - eps, unscaled (diagnostic)
    // FP32 traversal works on coordinates multiplied by the power of two S = `scale`, chosen so that the
    // root box spans ~2^20: keeps d^2 (d + eps) inside the FP32 exponent range from the box diagonal
    // down to separations of ~1e-19 of it.  Exact (power of two); it cancels in the accumulated force.
    double scale;
    float feps;                        // (float)(dist_eps * scale)
};

struct TreeArrays {
    // dense pyramid: level l occupies [level_off[l], level_off[l] + 4^l), Morton order inside
    double* mass;
    double* comx;
    double* comy;
    uint32_t* count;   // bodies per cell
    uint32_t* first;   // sorted position of first body
    uint32_t* flags;   // kNode* (verification / counting paths)
    NodeRec* rec;
    uint32_t* self_node;  // per body: pyramid index of the leaf holding only that body, else 0xffffffff
    uint32_t* tile_queue; // [2] work queue of the list traversal kernel (self-resetting, zero between launches)
};

struct Scratch {
    // zeroed by ONE memset at the start of every step:
    uint8_t* zero_base;
    size_t zero_bytes;
    uint32_t* digit_hist;     // [kMaxSortPasses][kMaxBins]
    uint32_t* tile_state;     // [passes][ntiles][nbins]
    uint32_t* tickets;        // [kMaxSortPasses] + misc counters
    uint32_t* heavy_count;    // 1
    unsigned long long* counters;  // [8] interactions, visits, opens, warp_steps, nodes, heavy
    uint32_t* bbox_ticket;    // 1
    uint32_t* scan_ticket;    // 1   cell_scan_kernel: tile order
    uint32_t* scan_state;     // [ceil(finest cells / 4096)] decoupled look-back states of the cell scan
    uint32_t* finest_count;   // alias of tree.count at the finest level (inside the zeroed block)
    // not zeroed:
    double* bbox_partial;     // [grid][4]
    uint32_t* heavy_list;     // finest cell ids
    uint32_t* huge_list;      // heavy cells with more than kHugeCellMin bodies (summed by many blocks)
    uint32_t* huge_count;     // 1 (zeroed)
    uint32_t* huge_tickets;   // per huge cell (zeroed)
    double* huge_partial;     // [huge cells][kHugeParts][3]
    double* cell_bnd;         // [2][2^F + 1] finest-level column / row boundaries (bounds kernel -> keys kernel)
    int64_t max_huge;
};

constexpr int kMaxHostChunks = 8;
struct ChunkBounds {
    uint32_t lo[kMaxHostChunks + 1];   // chunk k = bodies [lo[k], lo[k+1])
    int n_chunks;
};

constexpr int kMaxPeers = 8;   // GPUs of one NVSwitch box

// Peer-memory exchange (peer_comm.cu): every rank owns one cudaIpc-shared buffer with this layout.
struct PeerComm {
    int rank, n_ranks;
    uint8_t* peer_base[kMaxPeers];   // peer_base[rank] is this rank's own buffer
    size_t off_bbox, off_bbox_flag, off_in_flag, off_err, off_inbox;
    uint64_t ncells;                 // finest cells; inbox = [source rank][4][ncells] doubles (count, m, m x, m y)
    unsigned long long timeout_ns;   // wall-clock bound of every flag wait (env BH_PEER_TIMEOUT_MS, default 4000)
};

struct SortPlan {
    int key_bits, passes, bits_per_pass, nbins_log2;  // nbins_log2 in {8, 9}
    int ntiles;
    int items;         // keys per thread and tile (kSortItems or kSortItemsSmall); tile = kSortThreads * items keys
};

struct Dims {
    int64_t n;
    int max_depth;     // D
    int finest;        // F = D-1
    uint64_t level_off[kMaxLevels + 1];
    uint64_t ncells_finest;
    uint64_t npyramid;
};

// ---- kernel launchers (each in its own .cu) ------------------------------------------------
void launch_bounds(const double2* pos, int64_t n, const bh_params& p, const Dims& d, Scratch& s,
                   StepConsts* consts, int grid, cudaStream_t st, double* raw_out = nullptr,
                   const PeerComm* pc = nullptr);
void peer_comm_layout(PeerComm& pc, int rank, int n_ranks, uint64_t ncells, size_t* total_bytes);
void launch_bounds_finalize(const double* raw, const bh_params& p, const Dims& d, Scratch& s, StepConsts* consts,
                            cudaStream_t st);
void launch_keys(const double2* pos, int64_t n, const Dims& d, const SortPlan& sp, const StepConsts* consts,
                 uint32_t* keys, uint32_t* cell_count, uint32_t* digit_hist, cudaStream_t st, uint32_t idx_base,
                 const double* cell_bnd);   // cell_count: finest level of TreeArrays::count (zeroed); cell_bnd: table or nullptr
// pc != nullptr (sharded build with the peer-memory exchange): the partial sums of this rank's non-empty cells go
// straight into every rank's inbox (launch_tree_runs), and the level pass adds the ranks' contributions itself
void launch_tree_runs(const uint32_t* skeys, const uint32_t* sidx, const double2* pos, const double* mass,
                      int64_t n, const bh_params& p, const Dims& d, TreeArrays& t, Scratch& s, double* sums,
                      cudaStream_t st, const PeerComm* pc = nullptr);
void launch_tree_levels(const uint32_t* sidx, const double2* pos, const double* mass, const bh_params& p,
                        const Dims& d, TreeArrays& t, Scratch& s, const StepConsts* consts, const double* sums,
                        cudaStream_t st, const PeerComm* pc = nullptr);
void launch_sort(uint32_t* keys[2], uint32_t* vals[2], int64_t n, const SortPlan& sp, Scratch& s,
                 int* result_buf, cudaStream_t st, uint32_t val_base);   // values of pass 0 = val_base + input position
void launch_tree(const uint32_t* skeys, const uint32_t* sidx, const double2* pos, const double* mass,
                 int64_t n, const bh_params& p, const Dims& d, TreeArrays& t, Scratch& s,
                 const StepConsts* consts, cudaStream_t st);
void launch_traverse(const uint32_t* skeys, const uint32_t* sidx, const double2* pos_in, const double2* vel_in,
                     double2* pos, double2* vel, double2* acc,
                     double2* force, const double* mass, int64_t n, int64_t own_lo, int64_t own_hi,
                     const uint32_t* own_list, const uint32_t* own_count_dev, int64_t own_n,
                     const bh_params& p, const Dims& d, const TreeArrays& t, const StepConsts* consts,
                     unsigned long long* counters, bool integrate, cudaStream_t st);
void launch_integrate(double2* pos, double2* vel, double2* acc, const double2* force, const double* mass,
                      int64_t lo, int64_t hi, double dt, cudaStream_t st);
// stable partition of the sorted positions by ranges of original body index (pipelined bh_step_host):
// counts = scratch [n_chunks][ceil(n / 256)], lists = [n]; list k starts at lists[cb.lo[k]]
void launch_chunk_lists(const uint32_t* sidx, int64_t n, const ChunkBounds& cb, uint32_t* counts, uint32_t* lists,
                        cudaStream_t st);
void launch_direct(const double2* pos, const double* mass, int64_t n, double G, float4* packed,
                   double2* force, cudaStream_t st);
int measure_fp32_peak(int device, double* tflops, double* mhz);
// generate.cu: seeded initial conditions; bodies [i0, i1) into host arrays / bodies [lo, hi) into device arrays
int generate_host(int kind, uint64_t seed, int64_t i0, int64_t i1, double* pos, double* vel, double* mass);
void philox_host(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]);
void launch_generate(int kind, uint64_t seed, int64_t lo, int64_t hi, double2* pos, double2* vel, double* mass,
                     cudaStream_t st);

int traverse_launch_count();
extern thread_local uint64_t g_launches;  // kernels launched by this library on this thread

// Programmatic dependent launch (env BH_PDL=1, set per call by the API layer): the kernels of the
// single-GPU step chain (keys -> sort passes -> cell runs -> heavy / huge cells -> levels -> traversal) are
// launched with cudaLaunchAttributeProgrammaticStreamSerialization.  Each of them starts with
// pdl_entry(): `griddepcontrol.launch_dependents` lets the NEXT kernel's blocks be scheduled as soon as
// all blocks of this one are resident, `griddepcontrol.wait` then holds them until the PREVIOUS kernel
// has completed and flushed — so launch latency and block scheduling overlap the predecessor's tail
// while every memory dependency is kept (completion is transitive along the chain).  Both instructions
// are no-ops in a kernel launched without the attribute.
extern thread_local bool g_pdl;
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_entry() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
// Launch `kernel`; after_kernel = the previous operation on `st` is one of this library's chain kernels.
template <typename... KArgs, typename... Args>
inline void launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, bool after_kernel,
                         Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (g_pdl && after_kernel) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif

void set_error(const char* fmt, ...);
#define BH_CUDA_OK(expr)                                                                          \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            ::bh::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
            return BH_ERR_CUDA;                                                                   \
        }                                                                                         \
    } while (0)

}  // namespace bh
