// C-ABI of libbh.so: context life cycle, the per-step kernel sequence (optionally replayed as a
// CUDA graph), the multi-GPU exchange (NVLink peer memory or NCCL), the pipelined host step, getters in
// original body order.
// See include/bh.h for the contract and the reference lines each entry point replaces.
#include <dlfcn.h>
#include <nccl.h>   // types only; the library is dlopen'ed so that libbh.so has no hard NCCL dependency

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

#include "bh_internal.h"
#include "host_io.h"

namespace bh {

thread_local uint64_t g_launches = 0;
thread_local bool g_pdl = false;
static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

// ---- NCCL through dlopen ---------------------------------------------------------------------
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;
static std::once_flag g_nccl_once;
static void nccl_load();
static NcclApi* nccl_api() {
    std::call_once(g_nccl_once, nccl_load);     // contexts may live on several threads
    if (!g_nccl.handle) { set_error("libnccl not loadable or incomplete"); return nullptr; }
    return &g_nccl;
}
static void nccl_load() {
    NcclApi& api = g_nccl;
    // Order: a libnccl the process has ALREADY loaded (a python host's torch brings its own copy; a second copy of the
    // same SONAME cannot coexist), then BH_NCCL_LIB (explicit path for C++ hosts), then the system library.
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names) { h = dlopen(nm, RTLD_NOW | RTLD_NOLOAD); if (h) break; }
    if (!h) { const char* env = getenv("BH_NCCL_LIB"); if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_LOCAL); }
    if (!h) for (const char* nm : names) { h = dlopen(nm, RTLD_NOW | RTLD_LOCAL); if (h) break; }
    if (!h) return;
#define BH_SYM(field, name)                                                        \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));            \
    if (!api.field) return;
    BH_SYM(GetUniqueId, "ncclGetUniqueId");
    BH_SYM(CommInitRank, "ncclCommInitRank");
    BH_SYM(CommDestroy, "ncclCommDestroy");
    BH_SYM(Broadcast, "ncclBroadcast");
    BH_SYM(AllReduce, "ncclAllReduce");
    BH_SYM(GroupStart, "ncclGroupStart");
    BH_SYM(GroupEnd, "ncclGroupEnd");
    BH_SYM(GetErrorString, "ncclGetErrorString");
#undef BH_SYM
    api.handle = h;
}

#define BH_NCCL_OK(api, expr)                                                                   \
    do {                                                                                        \
        ncclResult_t r__ = (expr);                                                              \
        if (r__ != ncclSuccess) {                                                               \
            set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, (api)->GetErrorString(r__)); \
            return BH_ERR_NCCL;                                                                 \
        }                                                                                       \
    } while (0)

// ---- resident Morton order: kernels (see bh_ctx::perm) ----
__global__ void __launch_bounds__(256)
reorder_gather_kernel(const uint32_t* __restrict__ sidx, int64_t n, const double2* __restrict__ pos,
                      const double2* __restrict__ vel, const double2* __restrict__ acc, const double2* __restrict__ force,
                      const double* __restrict__ mass, const uint32_t* __restrict__ perm, double2* __restrict__ o2,
                      double* __restrict__ o1, uint32_t* __restrict__ ou) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t b = sidx[j];
    o2[j] = pos[b]; o2[n + j] = vel[b]; o2[2 * n + j] = acc[b]; o2[3 * n + j] = force[b];
    o1[j] = mass[b];
    ou[j] = perm ? perm[b] : b;
}
// out[perm[j]] = in[j]  (internal -> original order)
__global__ void __launch_bounds__(256)
unpermute2_kernel(const double2* __restrict__ in, const uint32_t* __restrict__ perm, int64_t n, double2* __restrict__ out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) out[perm[j]] = in[j];
}
// out[j] = in[perm[j]]  (original -> internal order)
__global__ void __launch_bounds__(256)
permute2_kernel(const double2* __restrict__ in, const uint32_t* __restrict__ perm, int64_t n, double2* __restrict__ out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) out[j] = in[perm[j]];
}

}  // namespace bh

using namespace bh;

struct bh_ctx {
    bh_params p;
    Dims d;
    SortPlan sp;
    int device = 0;
    cudaStream_t stream = nullptr;
    double2 *pos = nullptr, *vel = nullptr, *acc = nullptr, *force = nullptr, *snap_pos = nullptr,
            *snap_vel = nullptr;
    double* mass = nullptr;
    uint32_t* keys[2] = {nullptr, nullptr};
    uint32_t* idx[2] = {nullptr, nullptr};
    int sorted = 0;               // which of keys[]/idx[] holds the sorted result
    StepConsts* consts = nullptr;
    TreeArrays tree{};
    NodeRec* rec_alloc = nullptr;
    Scratch s{};
    float4* packed = nullptr;     // direct-sum kernel input
    // multi-GPU
    int64_t own_lo = 0, own_hi = 0;
    // pipelined host step (bh_step_host): bodies in `host_chunks` ranges of original index, each with its own
    // traversal launch, integrator launch and download, so that downloads overlap the remaining traversals
    int host_chunks = 1;             // env BH_HOST_CHUNKS (default 1 = one traversal, one download)
    uint32_t* chunk_lists = nullptr; // [n] sorted positions, stably partitioned by chunk
    uint32_t* chunk_counts = nullptr;
    cudaStream_t dl_stream = nullptr;
    cudaEvent_t ev_trav[kMaxHostChunks] = {}, ev_vel[kMaxHostChunks] = {}, ev_dl = nullptr, ev_fork = nullptr;
    cudaGraphExec_t host_graph = nullptr;   // bh_step_host captured for one set of (pinned) host pointers
    const void* host_graph_key[4] = {nullptr, nullptr, nullptr, nullptr};
    uint64_t host_graph_kernels = 0;
    bool host_graph_failed = false;
    bool host_pipeline_multi = true;    // multi-rank bh_step_host pipelined like the single-rank one (BH_HOST_PIPELINE_MULTI=0: plain sequence)
    // BH_HOST_TRACE=1: timeline of one bh_step_host call (direct submission), printed to stderr
    bool host_trace = false;
    std::vector<std::pair<const char*, cudaEvent_t>> trace_ev;
    double* cell_sums = nullptr;   // [4][finest cells]: count, m, m x, m y — all-reduced every step
    double* bbox_raw = nullptr;    // [4]: xmin, -xmax, ymin, -ymax — all-reduced (min) every step
    SortPlan sp_own;               // sort plan for this rank's slice
    bool tree_full = false;        // the tree on the device was built from ALL bodies (diagnostic getters)
    bool mass_complete = true;     // multi-rank: masses of the other ranks' slices have been gathered
    PeerComm pc{};                 // NVLink peer-memory exchange (peer_comm.cu)
    uint8_t* comm_buf = nullptr;   // this rank's cudaIpc-shared buffer
    size_t comm_bytes = 0;
    bool p2p_ready = false;
    ncclComm_t comm = nullptr;
    // graphs
    cudaGraphExec_t graph[2] = {nullptr, nullptr};   // [0] plain step, [1] step from snapshot
    uint64_t graph_kernels[2] = {0, 0};
    // timing
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t copy_stream = nullptr;           // uploads of bh_step_host overlap the build on `stream`
    cudaEvent_t ev_up[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t pev[8] = {};
    bool profiling = false;
    bh_timers timers{};
    uint64_t launches_base = 0;
    bool bodies_set = false, tree_valid = false, have_snapshot = false, timed = false;
    bool tree_built = false;         // some tree has been built (node count / root box of the last build stay readable after a step)
    uint64_t zero_mass_bodies = 0;   // counted on the host by bh_set_bodies (see bh_counters)
    // Resident Morton order (single rank): the body arrays are physically re-sorted into the order of the last sort
    // from time to time, so that "body index" ~ "sorted position": the per-body gathers / scatters of the build
    // and of the traversal prologue / epilogue become (nearly) coalesced.  `perm` maps the internal index to the
    // caller's original index; setters and getters go through it, so the C-ABI keeps its ORIGINAL-order contract.
    // Exact leaves on a multi-rank context: a leaf's bodies may live on other ranks, so the sharded build does not
    // do.  Every step the ranks all-gather the positions (NCCL broadcasts of the slices, 16 B/body), every rank
    // builds the FULL tree redundantly (north_star's original scheme) and walks only its own bodies (the sorted
    // positions whose body lies in its index slice: chunk_lists).  No CUDA graph (NCCL calls inside the step).
    bool exact_multi = false;
    uint32_t* perm = nullptr;        // [n]; meaningful only while !perm_identity
    bool perm_identity = true;
    bool auto_reorder = true;        // env BH_REORDER=0 disables; off in FP64 mode unless BH_REORDER=1
    bool force_reorder = false;      // env BH_REORDER=1: also in FP64 verification mode
    int steps_since_reorder = 0;
    uint64_t n_reorders = 0;
    double2* ro_tmp2 = nullptr;      // [4][n] double2: gather targets of pos, vel, acc, force
    double* ro_tmp1 = nullptr;       // [n]
    uint32_t* ro_tmpu = nullptr;     // [n]
    // trajectory output (bh_trajectory_*): two device snapshots + two pinned host buffers, background writer
    FrameWriter* traj = nullptr;
    double2* traj_dev[2] = {nullptr, nullptr};
    double* traj_host[2] = {nullptr, nullptr};
    cudaEvent_t traj_snap[2] = {nullptr, nullptr}, traj_ready[2] = {nullptr, nullptr};
    int traj_stride = 1, traj_slot = 0;
    int64_t traj_calls = 0;
    size_t step_zero_bytes = 0;      // see zero_scratch
    bool pdl = false;                // env BH_PDL=1: programmatic dependent launch along the single-GPU chain
    bool keys_table = false;         // env BH_KEYS_TABLE=1: cell keys from the boundary table instead of per-body bisection
    bool snapshot_by_copy = false;   // env BH_SNAPSHOT_COPY=1 (A/B switch, see enqueue_step)
    int bounds_grid = 1;
};

namespace {

int compute_dims(const bh_params& p, Dims& d, SortPlan& sp) {
    d.n = p.n_bodies;
    d.max_depth = p.max_depth;
    d.finest = p.max_depth - 1;
    uint64_t off = 0;
    for (int l = 0; l <= kMaxLevels; ++l) {
        d.level_off[l] = off;
        if (l <= d.finest) off += 1ull << (2 * l);
    }
    d.ncells_finest = 1ull << (2 * d.finest);
    d.npyramid = off;
    sp.key_bits = 2 * d.finest;
    int bits = sp.key_bits < 1 ? 1 : sp.key_bits;
    sp.passes = (bits + 8) / 9;
    sp.bits_per_pass = (bits + sp.passes - 1) / sp.passes;
    sp.nbins_log2 = sp.bits_per_pass <= 8 ? 8 : 9;
    sp.items = p.n_bodies < kSortSmallN ? kSortItemsSmall : kSortItems;
    const int64_t tile = (int64_t)kSortThreads * sp.items;
    sp.ntiles = (int)((p.n_bodies + tile - 1) / tile);
    if (sp.ntiles < 1) sp.ntiles = 1;
    return BH_OK;
}

template <typename T>
int dev_alloc(T** ptr, size_t count) {
    if (count == 0) count = 1;
    cudaError_t e = cudaMalloc((void**)ptr, count * sizeof(T));
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? BH_ERR_NOMEM : BH_ERR_CUDA;
    }
    return BH_OK;
}

#define BH_TRY(expr) do { int rc__ = (expr); if (rc__ != BH_OK) return rc__; } while (0)

constexpr int kReorderEvery = 16;       // steps between two physical re-sorts of the body arrays (resident Morton order)
constexpr int kRepartitionEvery = 64;   // the same on a multi-rank context (it costs an all-gather of the state + a full build)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int check_launch() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("kernel launch: %s", cudaGetErrorString(e)); return BH_ERR_CUDA; }
    return BH_OK;
}

// Peer-memory exchange: the device-side timeout word is sticky and fatal (peer_comm.cuh: wait_flag).  Called by every
// entry point that has just synchronised the stream; the context stays in error until bh_attach_peers is called again.
int peer_error(bh_ctx* c) {
    if (!c->p2p_ready) return BH_OK;
    uint32_t err = 0;
    BH_CUDA_OK(cudaMemcpy(&err, c->comm_buf + c->pc.off_err, sizeof err, cudaMemcpyDeviceToHost));
    if (err) {
        set_error("peer-memory exchange timed out (a rank did not reach the step); results are invalid — re-attach the peers");
        return BH_ERR_NCCL;
    }
    return BH_OK;
}

// All-gather of one owned slice per rank (ragged sizes allowed): grouped broadcasts.
int exchange_slices(bh_ctx* c, void* base, size_t elem_bytes) {
    if (c->p.n_ranks <= 1) return BH_OK;
    // (the communicator is checked BEFORE libnccl is touched: loading the system's libnccl into a process that imports
    // torch — with its own bundled libnccl of the same SONAME — afterwards breaks that import)
    if (!c->comm) { set_error("multi-rank context without an attached NCCL communicator"); return BH_ERR_NCCL; }
    NcclApi* api = nccl_api();
    if (!api) return BH_ERR_NCCL;
    BH_NCCL_OK(api, api->GroupStart());
    for (int r = 0; r < c->p.n_ranks; ++r) {
        int64_t lo, hi;
        bh_shard_range(c->p.n_bodies, c->p.n_ranks, r, &lo, &hi);
        char* ptr = (char*)base + (size_t)lo * elem_bytes;
        BH_NCCL_OK(api, api->Broadcast(ptr, ptr, (size_t)(hi - lo) * elem_bytes, ncclUint8, r, c->comm, c->stream));
    }
    BH_NCCL_OK(api, api->GroupEnd());
    return BH_OK;
}

// Per-step zeroing: ONE memset over [finest-level body counts | sort histograms, tile states, tickets,
// counters | huge-cell tickets] (laid out back to back by bh_create) and one 0xff fill of self_node.
void zero_scratch(bh_ctx* c) {
    cudaMemsetAsync(c->tree.count + c->d.level_off[c->d.finest], 0, c->step_zero_bytes, c->stream);
    cudaMemsetAsync(c->tree.self_node, 0xff, c->d.n * sizeof(uint32_t), c->stream);
}

void prof_mark(bh_ctx* c, int i) { if (c->profiling) cudaEventRecord(c->pev[i], c->stream); }

int allreduce_f64(bh_ctx* c, double* buf, size_t count, ncclRedOp_t op) {
    if (!c->comm) { set_error("multi-rank context without an attached NCCL communicator"); return BH_ERR_NCCL; }
    NcclApi* api = nccl_api();
    if (!api) return BH_ERR_NCCL;
    BH_NCCL_OK(api, api->AllReduce(buf, buf, count, ncclDouble, op, c->comm, c->stream));
    return BH_OK;
}

// bounds -> keys -> sort -> tree, on the context's stream.
//
// Single GPU (or full = true: diagnostic getters of a multi-rank context after an exchange of
// positions): everything over all N bodies.
// Multi GPU (sharded build): every rank keys and sorts ONLY its own index slice and sums its bodies
// per finest cell; two all-reduces make the tree global — 4 doubles (min of xmin, -xmax, ymin, -ymax)
// and 4 doubles per finest cell (count, m, m x, m y: 8.4 MB at the default cap, independent of N).
// The level pass then runs redundantly on the identical reduced sums.  No body data crosses NVLink.
// `src` = the positions the tree is built from: c->pos, or the snapshot when a step restarts from it
// (out-of-place step: every kernel reads the snapshot, only the fused integrator writes c->pos / c->vel,
// so no restore copy is needed).
// `mass_ready` (optional, pipelined host step): event to wait for before the first kernel that reads masses.
int enqueue_build(bh_ctx* c, bool full = false, const double2* src = nullptr, cudaEvent_t mass_ready = nullptr);
int enqueue_build(bh_ctx* c, bool full, const double2* src, cudaEvent_t mass_ready) {
    if (!src) src = c->pos;
    g_pdl = c->pdl;
    const bool own_list_needed = c->exact_multi && !full;
    if (own_list_needed) {      // exact leaves, several ranks: everybody needs everybody's bodies (see bh_ctx::exact_multi)
        if (!c->mass_complete) { BH_TRY(exchange_slices(c, c->mass, sizeof(double))); c->mass_complete = true; }
        BH_TRY(exchange_slices(c, (void*)src, sizeof(double2)));
        full = true;
    }
    zero_scratch(c);
    prof_mark(c, 0);
    const bool sharded = c->p.n_ranks > 1 && !full;
    if (!sharded) {
        launch_bounds(src, c->d.n, c->p, c->d, c->s, c->consts, c->bounds_grid, c->stream);
        launch_keys(src, c->d.n, c->d, c->sp, c->consts, c->keys[0], c->tree.count + c->d.level_off[c->d.finest], c->s.digit_hist, c->stream, 0,
                    c->s.cell_bnd);
        prof_mark(c, 1);
        launch_sort(c->keys, c->idx, c->d.n, c->sp, c->s, &c->sorted, c->stream, 0u);
        prof_mark(c, 2);
        launch_tree(c->keys[c->sorted], c->idx[c->sorted], src, c->mass, c->d.n, c->p, c->d, c->tree, c->s, c->consts,
                    c->stream);
        c->tree_full = true;
        if (own_list_needed) {   // sorted positions of this rank's bodies, in sorted order: [0, lo) | [lo, hi) | [hi, n)
            if (!c->chunk_lists) {
                BH_TRY(dev_alloc(&c->chunk_lists, (size_t)c->d.n));
                BH_TRY(dev_alloc(&c->chunk_counts, (size_t)kMaxHostChunks * ((c->d.n + 255) / 256)));
            }
            ChunkBounds cb{};
            cb.n_chunks = 3;
            cb.lo[0] = 0; cb.lo[1] = (uint32_t)c->own_lo; cb.lo[2] = (uint32_t)c->own_hi; cb.lo[3] = (uint32_t)c->d.n;
            launch_chunk_lists(c->idx[c->sorted], c->d.n, cb, c->chunk_counts, c->chunk_lists, c->stream);
        }
    } else {
        const int64_t lo = c->own_lo, n_own = c->own_hi - c->own_lo;
        if (c->p2p_ready) {   // box exchange fused into the bounds kernel (peer stores + flags)
            launch_bounds(src + lo, n_own, c->p, c->d, c->s, c->consts, c->bounds_grid, c->stream, nullptr, &c->pc);
        } else {
            launch_bounds(src + lo, n_own, c->p, c->d, c->s, c->consts, c->bounds_grid, c->stream, c->bbox_raw);
            BH_TRY(allreduce_f64(c, c->bbox_raw, 4, ncclMin));
            launch_bounds_finalize(c->bbox_raw, c->p, c->d, c->s, c->consts, c->stream);
        }
        launch_keys(src + lo, n_own, c->d, c->sp_own, c->consts, c->keys[0], c->tree.count + c->d.level_off[c->d.finest], c->s.digit_hist, c->stream,
                    (uint32_t)lo, c->s.cell_bnd);
        prof_mark(c, 1);
        launch_sort(c->keys, c->idx, n_own, c->sp_own, c->s, &c->sorted, c->stream, (uint32_t)lo);
        prof_mark(c, 2);
        if (mass_ready) cudaStreamWaitEvent(c->stream, mass_ready, 0);
        const PeerComm* pc = c->p2p_ready ? &c->pc : nullptr;
        // peer exchange: the partial sums of the non-empty cells go straight into every rank's inbox, the level pass
        // adds the contributions in rank order; otherwise one NCCL all-reduce of the dense sums
        launch_tree_runs(c->keys[c->sorted], c->idx[c->sorted], src, c->mass, n_own, c->p, c->d, c->tree, c->s,
                         c->cell_sums, c->stream, pc);
        prof_mark(c, 6);
        if (!pc) BH_TRY(allreduce_f64(c, c->cell_sums, 4 * c->d.ncells_finest, ncclSum));
        prof_mark(c, 7);
        launch_tree_levels(c->idx[c->sorted], src, c->mass, c->p, c->d, c->tree, c->s, c->consts, c->cell_sums, c->stream, pc);
        c->tree_full = false;
    }
    prof_mark(c, 3);
    return check_launch();
}

// `list` / `list_n` (optional): evaluate only the bodies at these sorted positions (pipelined host step);
// the kernel variant is still chosen from the TOTAL body count so that the bits equal a whole-set launch.
int enqueue_forces(bh_ctx* c, bool integrate, const double2* src_pos = nullptr, const double2* src_vel = nullptr,
                   const uint32_t* list = nullptr, int64_t list_n = 0) {
    if (c->exact_multi && !list) { list = c->chunk_lists + c->own_lo; list_n = c->own_hi - c->own_lo; }
    // after a sharded build the sorted list holds exactly this rank's bodies
    const int64_t n_all = (c->p.n_ranks > 1 && !c->tree_full) ? (c->own_hi - c->own_lo) : c->d.n;
    g_pdl = c->pdl;
    bh_params p = c->p;
    // (9 = exact leaves in the list kernel: validated for whole-set launches only, so an own-list launch falls back to
    // the pair / generic kernel like the default does)
    if (list && (p.reserved[0] == 0 || p.reserved[0] == 9)) p.reserved[0] = n_all >= kTwoBodiesPerLaneMin ? 2 : 1;
    // Exact leaves read OTHER bodies' positions (the members of a shared cap leaf) during the walk; a fused in-place
    // integrator in a warp that finished earlier would already have moved them.  Forces first, then one integrator
    // launch over the same bodies — unless the step is out of place (reads the snapshot, writes pos / vel).
    const bool split = integrate && (c->p.flags & BH_FLAG_EXACT_LEAVES) && !src_pos;
    launch_traverse(c->keys[c->sorted], c->idx[c->sorted], src_pos ? src_pos : c->pos, src_vel ? src_vel : c->vel,
                    c->pos, c->vel, c->acc, c->force, c->mass, c->d.n,
                    c->own_lo, c->own_hi, list, nullptr, list ? list_n : n_all, p, c->d, c->tree, c->consts, c->s.counters,
                    integrate && !split, c->stream);
    if (split) launch_integrate(c->pos, c->vel, c->acc, c->force, c->mass, c->own_lo, c->own_hi, c->p.dt, c->stream);
    prof_mark(c, 4);
    return check_launch();
}

int enqueue_step(bh_ctx* c, bool from_snapshot) {
    const double2 *src_pos = nullptr, *src_vel = nullptr;
    if (from_snapshot) {
        if (c->snapshot_by_copy) {   // BH_SNAPSHOT_COPY=1: restore by device-to-device copies, then step in place
            const int64_t lo = c->own_lo, cnt = c->own_hi - c->own_lo;   // a rank only touches its own slice
            cudaMemcpyAsync(c->pos + lo, c->snap_pos + lo, sizeof(double2) * cnt, cudaMemcpyDeviceToDevice, c->stream);
            cudaMemcpyAsync(c->vel + lo, c->snap_vel + lo, sizeof(double2) * cnt, cudaMemcpyDeviceToDevice, c->stream);
        } else {                     // default: the step reads the snapshot and writes pos / vel (no copies)
            src_pos = c->snap_pos; src_vel = c->snap_vel;
        }
    }
    BH_TRY(enqueue_build(c, false, src_pos));
    BH_TRY(enqueue_forces(c, true, src_pos, src_vel));
    prof_mark(c, 5);
    return BH_OK;
}

void accumulate_profile(bh_ctx* c) {
    if (!c->profiling) return;
    cudaEventSynchronize(c->pev[5]);
    float ms[5];
    for (int i = 0; i < 5; ++i) cudaEventElapsedTime(&ms[i], c->pev[i], c->pev[i + 1]);
    c->timers.bounds_keys_us += ms[0] * 1e3;
    c->timers.sort_us += ms[1] * 1e3;
    c->timers.build_us += ms[2] * 1e3;
    c->timers.traverse_us += ms[3] * 1e3;   // traversal with the fused integrator
    if (c->p.n_ranks > 1 && !c->tree_full) {   // sharded build: the per-cell sum exchange is the middle part of the build phase
        float ex = 0.f;
        cudaEventElapsedTime(&ex, c->pev[6], c->pev[7]);
        c->timers.exchange_us += ex * 1e3;
        c->timers.build_us -= ex * 1e3;
    }
    c->timers.total_us += (ms[0] + ms[1] + ms[2] + ms[3] + ms[4]) * 1e3;
    c->timers.steps += 1;
}

bool reorder_allowed(const bh_ctx* c);
int reorder_period(const bh_ctx* c);
int enqueue_reorder(bh_ctx* c);

int run_steps(bh_ctx* c, int nsteps, bool from_snapshot) {
    if (!c->bodies_set) { set_error("bh_step before bh_set_bodies"); return BH_ERR_INVALID; }
    if (from_snapshot && !c->have_snapshot) { set_error("no snapshot taken"); return BH_ERR_INVALID; }
    if (nsteps < 0) { set_error("nsteps < 0"); return BH_ERR_INVALID; }
    DeviceGuard g(c->device);
    // NCCL calls are not captured; the peer-memory exchange is plain kernels, so multi-rank steps replay too
    const bool use_graph = !(c->p.flags & BH_FLAG_NO_GRAPH) && !c->profiling && (c->p.n_ranks == 1 || c->p2p_ready) &&
                           !c->exact_multi;
    const int gi = from_snapshot ? 1 : 0;
    BH_CUDA_OK(cudaEventRecord(c->ev0, c->stream));
    if (use_graph && nsteps > 0) {
        if (!c->graph[gi]) {
            cudaGraph_t graph = nullptr;
            uint64_t before = g_launches;
            BH_CUDA_OK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
            int rc = enqueue_step(c, from_snapshot);
            cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
            if (rc != BH_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess) { set_error("stream capture: %s", cudaGetErrorString(e)); return BH_ERR_CUDA; }
            c->graph_kernels[gi] = g_launches - before;
            g_launches = before;   // capture launched nothing
            BH_CUDA_OK(cudaGraphInstantiate(&c->graph[gi], graph, 0));
            cudaGraphDestroy(graph);
        }
        for (int s = 0; s < nsteps; ++s) {
            if (!from_snapshot && reorder_allowed(c) && ++c->steps_since_reorder > reorder_period(c)) BH_TRY(enqueue_reorder(c));
            BH_CUDA_OK(cudaGraphLaunch(c->graph[gi], c->stream));
            g_launches += c->graph_kernels[gi];
        }
    } else {
        for (int s = 0; s < nsteps; ++s) {
            if (!from_snapshot && reorder_allowed(c) && !c->profiling && ++c->steps_since_reorder > reorder_period(c))
                BH_TRY(enqueue_reorder(c));
            BH_TRY(enqueue_step(c, from_snapshot));
            accumulate_profile(c);
        }
    }
    BH_CUDA_OK(cudaEventRecord(c->ev1, c->stream));
    c->timed = true;
    // the fused integrator has moved the bodies: the tree on the device describes the positions BEFORE the last step
    // (or the snapshot), so bh_compute_forces and the diagnostic getters must not pair it with the new positions
    if (nsteps > 0) { c->tree_valid = false; c->tree_built = true; }
    return BH_OK;
}

// ---- resident Morton order -----------------------------------------------------------------------------------------
int reorder_alloc(bh_ctx* c) {
    const size_t n = (size_t)c->d.n;
    if (!c->ro_tmp2) BH_TRY(dev_alloc(&c->ro_tmp2, 4 * n));
    if (!c->ro_tmp1) BH_TRY(dev_alloc(&c->ro_tmp1, n));
    if (!c->ro_tmpu) BH_TRY(dev_alloc(&c->ro_tmpu, n));
    if (!c->perm) BH_TRY(dev_alloc(&c->perm, n));
    return BH_OK;
}

// single rank: on unless BH_REORDER=0 (or FP64 verification mode without BH_REORDER=1).  Multi rank: the same, and it
// needs the NCCL communicator (the re-partition gathers the slices); collective — every rank takes the same decision.
bool reorder_allowed(const bh_ctx* c) {
    if (!c->auto_reorder) return false;
    if ((c->p.flags & BH_FLAG_FP64_TRAVERSAL) && !c->force_reorder) return false;
    if (c->p.flags & BH_FLAG_EXACT_LEAVES) return c->p.n_ranks == 1;
    return c->p.n_ranks == 1 || c->comm != nullptr;
}
int reorder_period(const bh_ctx* c) { return c->p.n_ranks > 1 ? kRepartitionEvery : kReorderEvery; }

// Sort the bodies by their current cell keys and move the state arrays into that order (asynchronous on the stream).
// Multi-rank contexts (RE-PARTITIONING): every rank holds full-size arrays of which only its index slice is
// current, so the slices are gathered first (NCCL broadcasts), every rank then does the SAME full build and the SAME
// permutation — after which the fixed index slices [N r / R, N (r + 1) / R) are contiguous Morton ranges again,
// whatever order the application handed the bodies over in and however far they have drifted since.
int enqueue_reorder(bh_ctx* c) {
    BH_TRY(reorder_alloc(c));
    const bool multi = c->p.n_ranks > 1;
    if (multi) {
        if (!c->mass_complete) { BH_TRY(exchange_slices(c, c->mass, sizeof(double))); c->mass_complete = true; }
        BH_TRY(exchange_slices(c, c->pos, sizeof(double2)));
        BH_TRY(exchange_slices(c, c->vel, sizeof(double2)));
        BH_TRY(exchange_slices(c, c->acc, sizeof(double2)));
        BH_TRY(exchange_slices(c, c->force, sizeof(double2)));
        if (c->have_snapshot) {
            BH_TRY(exchange_slices(c, c->snap_pos, sizeof(double2)));
            BH_TRY(exchange_slices(c, c->snap_vel, sizeof(double2)));
        }
    }
    BH_TRY(enqueue_build(c, multi, nullptr, nullptr));
    const int64_t n = c->d.n;
    reorder_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(
        c->idx[c->sorted], n, c->pos, c->vel, c->acc, c->force, c->mass, c->perm_identity ? nullptr : c->perm, c->ro_tmp2,
        c->ro_tmp1, c->ro_tmpu);
    ++g_launches;
    BH_TRY(check_launch());
    const size_t b2 = sizeof(double2) * (size_t)n;
    BH_CUDA_OK(cudaMemcpyAsync(c->pos, c->ro_tmp2, b2, cudaMemcpyDeviceToDevice, c->stream));
    BH_CUDA_OK(cudaMemcpyAsync(c->vel, c->ro_tmp2 + n, b2, cudaMemcpyDeviceToDevice, c->stream));
    BH_CUDA_OK(cudaMemcpyAsync(c->acc, c->ro_tmp2 + 2 * n, b2, cudaMemcpyDeviceToDevice, c->stream));
    BH_CUDA_OK(cudaMemcpyAsync(c->force, c->ro_tmp2 + 3 * n, b2, cudaMemcpyDeviceToDevice, c->stream));
    BH_CUDA_OK(cudaMemcpyAsync(c->mass, c->ro_tmp1, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
    BH_CUDA_OK(cudaMemcpyAsync(c->perm, c->ro_tmpu, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
    if (c->have_snapshot) {         // keep a snapshot taken earlier consistent with the new internal order
        permute2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->snap_pos, c->idx[c->sorted], n, c->ro_tmp2);
        permute2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->snap_vel, c->idx[c->sorted], n, c->ro_tmp2 + n);
        BH_TRY(check_launch());
        BH_CUDA_OK(cudaMemcpyAsync(c->snap_pos, c->ro_tmp2, b2, cudaMemcpyDeviceToDevice, c->stream));
        BH_CUDA_OK(cudaMemcpyAsync(c->snap_vel, c->ro_tmp2 + n, b2, cudaMemcpyDeviceToDevice, c->stream));
    }
    c->perm_identity = false;
    c->steps_since_reorder = 0;
    c->n_reorders += 1;
    c->tree_valid = false;          // sorted positions / self_node refer to the old internal indices
    if (multi) c->mass_complete = true;
    return BH_OK;
}

int copy_out2(bh_ctx* c, const double2* dev, double* host, bool gather_ranks) {
    if (!host) { set_error("null output buffer"); return BH_ERR_INVALID; }
    DeviceGuard g(c->device);
    // multi-rank: every rank only keeps its own slice current; getters are collective and gather first
    if (gather_ranks && c->p.n_ranks > 1) BH_TRY(exchange_slices(c, (void*)dev, sizeof(double2)));
    if (!c->perm_identity) {     // internal (resident Morton) order -> the caller's original order
        BH_TRY(reorder_alloc(c));
        unpermute2_kernel<<<(unsigned)((c->d.n + 255) / 256), 256, 0, c->stream>>>(dev, c->perm, c->d.n, c->ro_tmp2);
        BH_TRY(check_launch());
        dev = c->ro_tmp2;
    }
    BH_CUDA_OK(cudaMemcpyAsync(host, dev, sizeof(double2) * c->d.n, cudaMemcpyDeviceToHost, c->stream));
    BH_CUDA_OK(cudaStreamSynchronize(c->stream));
    return peer_error(c);
}

// Diagnostic getters (keys, order, node table, dump) of a multi-rank context need the tree over
// ALL bodies with global sorted positions: gather positions, then build like a single GPU.
int ensure_full_tree(bh_ctx* c) {
    if (c->p.n_ranks <= 1 || c->tree_full) return BH_OK;
    if (!c->mass_complete) { BH_TRY(exchange_slices(c, c->mass, sizeof(double))); c->mass_complete = true; }
    BH_TRY(exchange_slices(c, c->pos, sizeof(double2)));
    BH_TRY(enqueue_build(c, true));
    return BH_OK;
}

void trace_mark(bh_ctx* c, const char* label, cudaStream_t st) {
    if (!c->host_trace) return;
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    cudaEventRecord(ev, st);
    c->trace_ev.emplace_back(label, ev);
}

void trace_dump(bh_ctx* c) {
    if (!c->host_trace) return;
    for (auto& le : c->trace_ev) {
        float ms = 0.f;
        cudaEventSynchronize(le.second);
        cudaEventElapsedTime(&ms, c->ev0, le.second);
        fprintf(stderr, "[bh_step_host trace] %-22s %8.3f ms\n", le.first, ms);
        cudaEventDestroy(le.second);
    }
    c->trace_ev.clear();
}

bool host_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// One step with host buffers on three streams (all forked from / joined to c->stream, so the sequence can
// be captured into a graph):
//   copy_stream : positions, masses, then the velocities chunk by chunk (the order the step needs them)
//   stream      : bounds, keys, sort, chunk lists, tree, then one forces-only traversal per index chunk
//   dl_stream   : (high priority) per chunk: integrator as soon as the chunk's forces and velocities are
//                 there, then the contiguous download of the chunk's new positions — the downloads overlap
//                 the traversal of the remaining chunks.
// Chunks are ranges of ORIGINAL body index (contiguous in the caller's arrays); chunk lists keep Morton
// order inside a chunk (launch_chunk_lists), and per-body results do not depend on the grouping.
int enqueue_host_step(bh_ctx* c, const double* pos, const double* vel, const double* mass, double* out_pos, int nch) {
    const int64_t n = c->d.n;
    ChunkBounds cb{};
    cb.n_chunks = nch;
    for (int k = 0; k <= nch; ++k) cb.lo[k] = (uint32_t)(((__int128)n * k) / nch);
    g_pdl = false;   // event waits sit between the kernels of this path: plain launches
    BH_CUDA_OK(cudaEventRecord(c->ev_fork, c->stream));
    BH_CUDA_OK(cudaStreamWaitEvent(c->copy_stream, c->ev_fork, 0));
    BH_CUDA_OK(cudaMemcpyAsync(c->pos, pos, sizeof(double2) * n, cudaMemcpyHostToDevice, c->copy_stream));
    BH_CUDA_OK(cudaEventRecord(c->ev_up[0], c->copy_stream));
    trace_mark(c, "up: pos done", c->copy_stream);
    BH_CUDA_OK(cudaMemcpyAsync(c->mass, mass, sizeof(double) * n, cudaMemcpyHostToDevice, c->copy_stream));
    BH_CUDA_OK(cudaEventRecord(c->ev_up[1], c->copy_stream));
    trace_mark(c, "up: mass done", c->copy_stream);
    for (int k = 0; k < nch; ++k) {
        const int64_t lo = cb.lo[k], cnt = (int64_t)cb.lo[k + 1] - lo;
        BH_CUDA_OK(cudaMemcpyAsync(c->vel + lo, vel + 2 * lo, sizeof(double2) * cnt, cudaMemcpyHostToDevice, c->copy_stream));
        BH_CUDA_OK(cudaEventRecord(c->ev_vel[k], c->copy_stream));
        trace_mark(c, "up: vel chunk done", c->copy_stream);
    }
    zero_scratch(c);
    BH_CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_up[0], 0));
    launch_bounds(c->pos, n, c->p, c->d, c->s, c->consts, c->bounds_grid, c->stream);
    launch_keys(c->pos, n, c->d, c->sp, c->consts, c->keys[0], c->tree.count + c->d.level_off[c->d.finest], c->s.digit_hist, c->stream, 0, c->s.cell_bnd);
    launch_sort(c->keys, c->idx, n, c->sp, c->s, &c->sorted, c->stream, 0u);
    if (nch > 1) launch_chunk_lists(c->idx[c->sorted], n, cb, c->chunk_counts, c->chunk_lists, c->stream);
    trace_mark(c, "gpu: sort+lists done", c->stream);
    BH_CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_up[1], 0));
    launch_tree(c->keys[c->sorted], c->idx[c->sorted], c->pos, c->mass, n, c->p, c->d, c->tree, c->s, c->consts, c->stream);
    c->tree_full = true;
    BH_TRY(check_launch());
    trace_mark(c, "gpu: tree done", c->stream);
    for (int k = 0; k < nch; ++k) {
        const int64_t lo = cb.lo[k], hi = cb.lo[k + 1];
        if (nch > 1) BH_TRY(enqueue_forces(c, false, nullptr, nullptr, c->chunk_lists + lo, hi - lo));
        else BH_TRY(enqueue_forces(c, false));
        BH_CUDA_OK(cudaEventRecord(c->ev_trav[k], c->stream));
        trace_mark(c, "gpu: traverse chunk", c->stream);
        BH_CUDA_OK(cudaStreamWaitEvent(c->dl_stream, c->ev_trav[k], 0));
        BH_CUDA_OK(cudaStreamWaitEvent(c->dl_stream, c->ev_vel[k], 0));
        launch_integrate(c->pos, c->vel, c->acc, c->force, c->mass, lo, hi, c->p.dt, c->dl_stream);
        BH_TRY(check_launch());
        trace_mark(c, "dl: integrate chunk", c->dl_stream);
        BH_CUDA_OK(cudaMemcpyAsync(out_pos + 2 * lo, c->pos + lo, sizeof(double2) * (hi - lo), cudaMemcpyDeviceToHost,
                                   c->dl_stream));
        trace_mark(c, "dl: download chunk", c->dl_stream);
    }
    BH_CUDA_OK(cudaEventRecord(c->ev_dl, c->dl_stream));
    BH_CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_dl, 0));
    return BH_OK;
}

}  // namespace

extern "C" {

const char* bh_last_error(void) { return g_err; }
int bh_abi_version(void) { return BH_ABI_VERSION; }

void bh_default_params(bh_params* p) {
    memset(p, 0, sizeof *p);
    p->n_bodies = 40000;
    p->G = 6.67e-11; p->dt = 1.0; p->theta = 5e-1; p->dist_eps = 1e-15; p->mass_eps = 1e-15;
    p->pad_frac = 0.1; p->pad_fallback = 1e-6; p->max_depth = 10; p->device = -1; p->flags = 0;
    p->exact_leaf_max = 64; p->rank = 0; p->n_ranks = 1;
}

int bh_shard_range(int64_t n, int32_t n_ranks, int32_t rank, int64_t* lo, int64_t* hi) {
    if (n < 0 || n_ranks < 1 || rank < 0 || rank >= n_ranks || !lo || !hi) {
        set_error("bh_shard_range: bad arguments");
        return BH_ERR_INVALID;
    }
    *lo = (int64_t)(((__int128)n * rank) / n_ranks);
    *hi = (int64_t)(((__int128)n * (rank + 1)) / n_ranks);
    return BH_OK;
}

int bh_create(const bh_params* p, bh_ctx** out) {
    if (!p || !out) { set_error("bh_create: null argument"); return BH_ERR_INVALID; }
    if (p->n_bodies < 1 || p->n_bodies >= (1ll << 30)) { set_error("n_bodies must be in [1, 2^30)"); return BH_ERR_INVALID; }
    if (p->max_depth < 1 || p->max_depth > kMaxDepthDense) {
        set_error("max_depth must be in [1, %d] (dense pyramid)", kMaxDepthDense);
        return BH_ERR_INVALID;
    }
    if (p->n_ranks < 1 || p->rank < 0 || p->rank >= p->n_ranks) { set_error("bad rank / n_ranks"); return BH_ERR_INVALID; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
        return BH_ERR_CUDA;
    }
    bh_ctx* c = new bh_ctx();
    c->p = *p;
    // A/B switches (environment, read once per context)
    { const char* e = getenv("BH_SNAPSHOT_COPY"); c->snapshot_by_copy = e && e[0] == '1'; }
    // cell keys from the boundary table (same bits as the per-body bisection): 220 -> 159 us for bounds + keys at 16M
    // bodies (the bisection is 415 instructions per body), +2 us at 1M; BH_KEYS_TABLE=0 / 1 forces either
    { const char* e = getenv("BH_KEYS_TABLE"); c->keys_table = e ? e[0] == '1' : (p->n_bodies / p->n_ranks >= 1500000); }
    { const char* e = getenv("BH_PDL"); c->pdl = e && e[0] == '1' && p->n_ranks == 1; }
    { const char* e = getenv("BH_HOST_TRACE"); c->host_trace = e && e[0] == '1'; }
    { const char* e = getenv("BH_REORDER"); c->auto_reorder = !(e && e[0] == '0'); c->force_reorder = e && e[0] == '1'; }
    { const char* e = getenv("BH_HOST_PIPELINE_MULTI"); c->host_pipeline_multi = !(e && e[0] == '0'); }   // default on
    { const char* e = getenv("BH_HOST_CHUNKS"); if (e && atoi(e) >= 1) c->host_chunks = std::min(atoi(e), kMaxHostChunks); }
    if (p->device >= 0) c->device = p->device;
    else if ((e = cudaGetDevice(&c->device)) != cudaSuccess) { set_error("cudaGetDevice: %s", cudaGetErrorString(e)); delete c; return BH_ERR_CUDA; }
    if (c->device >= ndev) { set_error("device %d of %d", c->device, ndev); delete c; return BH_ERR_INVALID; }
    DeviceGuard g(c->device);
    compute_dims(c->p, c->d, c->sp);
    c->exact_multi = (p->flags & BH_FLAG_EXACT_LEAVES) && p->n_ranks > 1;
    const int64_t n = p->n_bodies;
    int rc = BH_OK;
    auto fail = [&](int code) { bh_destroy(c); return code; };
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        set_error("cudaStreamCreate: %s", cudaGetErrorString(e));
        return fail(BH_ERR_CUDA);
    }
#define BH_ALLOC(ptr, count) if ((rc = dev_alloc(&(ptr), (size_t)(count))) != BH_OK) return fail(rc)
    BH_ALLOC(c->pos, n); BH_ALLOC(c->vel, n); BH_ALLOC(c->acc, n); BH_ALLOC(c->force, n); BH_ALLOC(c->mass, n);
    BH_ALLOC(c->keys[0], n); BH_ALLOC(c->keys[1], n); BH_ALLOC(c->idx[0], n); BH_ALLOC(c->idx[1], n);
    BH_ALLOC(c->consts, 1);
    const uint64_t np = c->d.npyramid;
    BH_ALLOC(c->tree.mass, np); BH_ALLOC(c->tree.comx, np); BH_ALLOC(c->tree.comy, np);
    BH_ALLOC(c->tree.first, np); BH_ALLOC(c->tree.flags, np);
    BH_ALLOC(c->rec_alloc, np + 4);
    c->tree.rec = c->rec_alloc + 3;   // 96-byte lead-in: sibling groups (4p+1 .. 4p+4) start on 128-byte lines
    BH_ALLOC(c->tree.self_node, n);
    BH_ALLOC(c->tree.tile_queue, 2);
    // One allocation: [tree.count: all levels, finest level last][zeroed scratch block][huge-cell tickets].
    // Everything from the finest level's counts to the end is zeroed by ONE memset per step (zero_scratch).
    const int nbins = 1 << c->sp.nbins_log2;
    const size_t np_pad = ((size_t)np + 3) & ~(size_t)3;   // keeps the 64-bit counters 8-byte aligned
    const size_t scan_tiles = (size_t)((c->d.ncells_finest + 4095) / 4096);
    size_t words = (size_t)kMaxSortPasses * kMaxBins + 16 /*tickets, heavy, bbox ticket*/ + 16 /*8 x u64 counters*/ +
                   (size_t)c->sp.passes * c->sp.ntiles * nbins + scan_tiles;
    words = (words + 3) & ~(size_t)3;
    c->s.max_huge = n / kHugeCellMin + 1;
    const size_t huge_words = (size_t)c->s.max_huge + 1;    // [max_huge] = huge_count
    BH_ALLOC(c->tree.count, np_pad + words + huge_words);
    uint32_t* zb = c->tree.count + np_pad;
    c->step_zero_bytes = (np_pad - (size_t)c->d.level_off[c->d.finest] + words + huge_words) * sizeof(uint32_t);
    c->s.zero_base = (uint8_t*)zb;
    c->s.zero_bytes = words * sizeof(uint32_t);
    c->s.digit_hist = zb;
    c->s.tickets = zb + (size_t)kMaxSortPasses * kMaxBins;
    c->s.heavy_count = c->s.tickets + 8;
    c->s.bbox_ticket = c->s.tickets + 9;
    c->s.counters = (unsigned long long*)(c->s.tickets + 16);
    c->s.tile_state = c->s.tickets + 32;
    c->s.scan_ticket = c->s.tickets + 10;
    c->s.scan_state = c->s.tile_state + (size_t)c->sp.passes * c->sp.ntiles * nbins;
    c->s.huge_tickets = zb + words;
    c->s.huge_count = c->s.huge_tickets + c->s.max_huge;
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, c->device)) != cudaSuccess) { set_error("%s", cudaGetErrorString(e)); return fail(BH_ERR_CUDA); }
    int64_t bg = (n + 256 * 8 - 1) / (256 * 8);
    c->bounds_grid = (int)std::max<int64_t>(1, std::min<int64_t>(bg, prop.multiProcessorCount * 4));
    BH_ALLOC(c->s.bbox_partial, (size_t)c->bounds_grid * 4);
    BH_ALLOC(c->s.heavy_list, c->d.ncells_finest);
    BH_ALLOC(c->s.huge_list, c->s.max_huge);
    BH_ALLOC(c->s.huge_partial, (size_t)c->s.max_huge * kHugeParts * 3);
    // Cell keys: per-body FP64 bisection below 1.5M bodies per rank, the boundary-table variant (same bits) above:
    // at 1M bodies the two are equal (latency-bound), at 16M the table saves 60 us.
    if (c->keys_table) BH_ALLOC(c->s.cell_bnd, 2 * (((size_t)1 << c->d.finest) + 1));
    if (p->n_ranks > 1) {
        BH_ALLOC(c->cell_sums, 4 * c->d.ncells_finest); BH_ALLOC(c->bbox_raw, 4);
    }
#undef BH_ALLOC
    bh_shard_range(n, p->n_ranks, p->rank, &c->own_lo, &c->own_hi);
    c->sp_own = c->sp;
    {   // a rank sorts only its slice: keep the allocation of the full plan (it has at least as many tile states)
        const int64_t n_own = c->own_hi - c->own_lo;
        c->sp_own.items = c->sp.items;
        const int64_t tile = (int64_t)kSortThreads * c->sp_own.items;
        c->sp_own.ntiles = (int)std::max<int64_t>(1, (n_own + tile - 1) / tile);
    }
    cudaEventCreate(&c->ev0); cudaEventCreate(&c->ev1);
    cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    for (auto& ev : c->ev_up) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    for (auto& ev : c->pev) cudaEventCreate(&ev);
    {   // integrator + downloads of the pipelined host step must not queue behind the next chunk's traversal blocks
        int least = 0, greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        cudaStreamCreateWithPriority(&c->dl_stream, cudaStreamNonBlocking, greatest);
    }
    cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    for (auto& ev : c->ev_trav) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    for (auto& ev : c->ev_vel) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_dl, cudaEventDisableTiming);
    cudaMemsetAsync(c->tree.tile_queue, 0, 2 * sizeof(uint32_t), c->stream);
    cudaMemsetAsync(c->acc, 0, sizeof(double2) * n, c->stream);
    cudaMemsetAsync(c->force, 0, sizeof(double2) * n, c->stream);
    cudaMemsetAsync(c->tree.count, 0, sizeof(uint32_t) * (np_pad + words + huge_words), c->stream);
    if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) { set_error("%s", cudaGetErrorString(e)); return fail(BH_ERR_CUDA); }
    c->launches_base = g_launches;
    *out = c;
    return BH_OK;
}

int bh_destroy(bh_ctx* c) {
    if (!c) return BH_OK;
    DeviceGuard g(c->device);
    if (c->traj) bh_trajectory_end(c);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto& gr : c->graph) if (gr) cudaGraphExecDestroy(gr);
    if (c->comm) { NcclApi* api = nccl_api(); if (api) api->CommDestroy(c->comm); }
    if (c->p2p_ready)
        for (int r = 0; r < c->p.n_ranks; ++r)
            if (r != c->p.rank && c->pc.peer_base[r]) cudaIpcCloseMemHandle(c->pc.peer_base[r]);
    if (c->comm_buf) cudaFree(c->comm_buf);
    void* ptrs[] = {c->pos, c->vel, c->acc, c->force, c->snap_pos, c->snap_vel, c->mass, c->keys[0],
                    c->keys[1], c->idx[0], c->idx[1], c->consts, c->tree.mass, c->tree.comx, c->tree.comy,
                    c->tree.count /* + scratch + tickets */, c->tree.first, c->tree.flags, c->rec_alloc, c->tree.self_node, c->tree.tile_queue,
                    c->s.bbox_partial, c->s.heavy_list, c->s.huge_list, c->s.huge_partial, c->s.cell_bnd,
                    c->packed, c->chunk_lists, c->chunk_counts, c->cell_sums, c->bbox_raw, c->perm, c->ro_tmp2, c->ro_tmp1,
                    c->ro_tmpu};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (auto& ev : c->pev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : c->ev_up) if (ev) cudaEventDestroy(ev);
    for (auto& ev : c->ev_trav) if (ev) cudaEventDestroy(ev);
    for (auto& ev : c->ev_vel) if (ev) cudaEventDestroy(ev);
    if (c->ev_dl) cudaEventDestroy(c->ev_dl);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->host_graph) cudaGraphExecDestroy(c->host_graph);
    if (c->dl_stream) cudaStreamDestroy(c->dl_stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return BH_OK;
}

int bh_nccl_unique_id(void* id128) {
    NcclApi* api = nccl_api();
    if (!api) return BH_ERR_NCCL;
    ncclUniqueId id;
    BH_NCCL_OK(api, api->GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return BH_OK;
}

int bh_attach_nccl(bh_ctx* c, const void* id128) {
    if (!c || !id128) { set_error("null argument"); return BH_ERR_INVALID; }
    NcclApi* api = nccl_api();
    if (!api) return BH_ERR_NCCL;
    DeviceGuard g(c->device);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    BH_NCCL_OK(api, api->CommInitRank(&c->comm, c->p.n_ranks, id, c->p.rank));
    return BH_OK;
}

int bh_comm_handle(bh_ctx* c, void* handle64) {
    if (!c || !handle64) { set_error("null argument"); return BH_ERR_INVALID; }
    if (c->p.n_ranks < 2 || c->p.n_ranks > kMaxPeers) { set_error("peer exchange needs 2..%d ranks", kMaxPeers); return BH_ERR_INVALID; }
    DeviceGuard g(c->device);
    if (!c->comm_buf) {
        peer_comm_layout(c->pc, c->p.rank, c->p.n_ranks, c->d.ncells_finest, &c->comm_bytes);
        {   const char* e = getenv("BH_PEER_TIMEOUT_MS");
            const double ms = e && atof(e) > 0 ? atof(e) : 4000.0;
            c->pc.timeout_ns = (unsigned long long)(ms * 1e6); }
        BH_CUDA_OK(cudaMalloc((void**)&c->comm_buf, c->comm_bytes));
        BH_CUDA_OK(cudaMemset(c->comm_buf, 0, c->comm_bytes));
        for (int r = 0; r < kMaxPeers; ++r) c->pc.peer_base[r] = nullptr;
        c->pc.peer_base[c->p.rank] = c->comm_buf;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    BH_CUDA_OK(cudaIpcGetMemHandle(&h, c->comm_buf));
    memcpy(handle64, &h, sizeof h);
    return BH_OK;
}

int bh_attach_peers(bh_ctx* c, const void* handles, int32_t n_handles) {
    if (!c || !handles) { set_error("null argument"); return BH_ERR_INVALID; }
    if (n_handles != c->p.n_ranks || !c->comm_buf) { set_error("bh_attach_peers: call bh_comm_handle on every rank first and pass n_ranks handles"); return BH_ERR_INVALID; }
    DeviceGuard g(c->device);
    BH_CUDA_OK(cudaStreamSynchronize(c->stream));
    if (c->p2p_ready) {                      // re-attach: drop the old mappings and graphs, clear the sticky error + flags
        for (int r = 0; r < c->p.n_ranks; ++r)
            if (r != c->p.rank && c->pc.peer_base[r]) { cudaIpcCloseMemHandle(c->pc.peer_base[r]); c->pc.peer_base[r] = nullptr; }
        for (auto& gr : c->graph) if (gr) { cudaGraphExecDestroy(gr); gr = nullptr; }
        c->p2p_ready = false;
    }
    BH_CUDA_OK(cudaMemset(c->comm_buf, 0, c->comm_bytes));   // box slots, flags, error word, sequence number, tickets, inbox
    for (int r = 0; r < c->p.n_ranks; ++r) {
        if (r == c->p.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + 64 * (size_t)r, sizeof h);
        void* ptr = nullptr;
        BH_CUDA_OK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        c->pc.peer_base[r] = (uint8_t*)ptr;
    }
    c->p2p_ready = true;
    return BH_OK;
}

int bh_set_bodies(bh_ctx* c, const double* pos, const double* vel, const double* mass) {
    if (!c || !pos || !vel || !mass) { set_error("null argument"); return BH_ERR_INVALID; }
    DeviceGuard g(c->device);
    const int64_t n = c->d.n;
    // A rank of a multi-rank context uploads only ITS slice (the host buffers still hold all bodies,
    // original order): the sharded build needs nothing else.  Diagnostic getters gather on demand.
    (void)n;
    const int64_t lo = c->own_lo, cnt = c->own_hi - c->own_lo;
    BH_CUDA_OK(cudaMemcpyAsync(c->pos + lo, pos + 2 * lo, sizeof(double2) * cnt, cudaMemcpyHostToDevice, c->stream));
    BH_CUDA_OK(cudaMemcpyAsync(c->vel + lo, vel + 2 * lo, sizeof(double2) * cnt, cudaMemcpyHostToDevice, c->stream));
    BH_CUDA_OK(cudaMemcpyAsync(c->mass + lo, mass + lo, sizeof(double) * cnt, cudaMemcpyHostToDevice, c->stream));
    c->mass_complete = c->p.n_ranks == 1;
    uint64_t zeros = 0;                      // overlaps the copies
    for (int64_t i = lo; i < lo + cnt; ++i) zeros += mass[i] == 0.0;
    c->zero_mass_bodies = zeros;
    BH_CUDA_OK(cudaStreamSynchronize(c->stream));
    BH_TRY(check_launch());
    c->bodies_set = true;
    c->tree_valid = false;
    c->perm_identity = true;                       // caller's order
    c->steps_since_reorder = reorder_period(c) - 1;   // a multi-step run re-sorts the arrays before its second step
    return BH_OK;
}

static int set_vec(bh_ctx* c, double2* dst, const double* src) {
    if (!c || !src) { set_error("null argument"); return BH_ERR_INVALID; }
    if (!c->bodies_set) { set_error("bh_set_bodies first"); return BH_ERR_INVALID; }
    DeviceGuard g(c->device);
    const int64_t n = c->d.n;
    if (!c->perm_identity) {    // the arrays are in resident order: upload in the caller's order, then permute
        BH_TRY(reorder_alloc(c));
        BH_CUDA_OK(cudaMemcpyAsync(c->ro_tmp2, src, sizeof(double2) * n, cudaMemcpyHostToDevice, c->stream));
        permute2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->ro_tmp2, c->perm, n, dst);
        BH_TRY(check_launch());
    } else {
        BH_CUDA_OK(cudaMemcpyAsync(dst + c->own_lo, src + 2 * c->own_lo, sizeof(double2) * (c->own_hi - c->own_lo),
                                   cudaMemcpyHostToDevice, c->stream));
    }
    BH_CUDA_OK(cudaStreamSynchronize(c->stream));
    c->tree_valid = false;
    return BH_OK;
}
int bh_set_positions(bh_ctx* c, const double* pos) { return set_vec(c, c ? c->pos : nullptr, pos); }
int bh_set_velocities(bh_ctx* c, const double* vel) { return set_vec(c, c ? c->vel : nullptr, vel); }

int bh_snapshot(bh_ctx* c) {
    if (!c || !c->bodies_set) { set_error("bh_snapshot: no bodies"); return BH_ERR_INVALID; }
    DeviceGuard g(c->device);
    const int64_t n = c->d.n;
    // steps that restart from the snapshot cannot re-sort the arrays in between: take the snapshot in resident order
    if (reorder_allowed(c)) BH_TRY(enqueue_reorder(c));
    if (!c->snap_pos) { BH_TRY(dev_alloc(&c->snap_pos, (size_t)n)); BH_TRY(dev_alloc(&c->snap_vel, (size_t)n)); }
    BH_CUDA_OK(cudaMemcpyAsync(c->snap_pos, c->pos, sizeof(double2) * n, cudaMemcpyDeviceToDevice, c->stream));
    BH_CUDA_OK(cudaMemcpyAsync(c->snap_vel, c->vel, sizeof(double2) * n, cudaMemcpyDeviceToDevice, c->stream));
    BH_CUDA_OK(cudaStreamSynchronize(c->stream));
    c->have_snapshot = true;
    return BH_OK;
}

int bh_restore(bh_ctx* c) {
    if (!c || !c->have_snapshot) { set_error("bh_restore: no snapshot"); return BH_ERR_INVALID; }
    DeviceGuard g(c->device);
    const int64_t n = c->d.n;
    BH_CUDA_OK(cudaMemcpyAsync(c->pos, c->snap_pos, sizeof(double2) * n, cudaMemcpyDeviceToDevice, c->stream));
    BH_CUDA_OK(cudaMemcpyAsync(c->vel, c->snap_vel, sizeof(double2) * n, cudaMemcpyDeviceToDevice, c->stream));
    c->tree_valid = false;
    return BH_OK;
}

int bh_step(bh_ctx* c, int32_t nsteps) {
    if (!c) { set_error("null context"); return BH_ERR_INVALID; }
    return run_steps(c, nsteps, false);
}
int bh_step_from_snapshot(bh_ctx* c, int32_t nsteps) {
    if (!c) { set_error("null context"); return BH_ERR_INVALID; }
    return run_steps(c, nsteps, true);
}

// One step with HOST buffers, pipelined: positions go up first and bounds / keys / sort start as soon
// as they land; masses are only needed by the tree pass and velocities only by the integrator, so
// their uploads overlap the build and the traversal.  out_pos_host receives the new positions.
int bh_step_host(bh_ctx* c, const double* pos, const double* vel, const double* mass, double* out_pos) {
    if (!c || !pos || !vel || !mass || !out_pos) { set_error("null argument"); return BH_ERR_INVALID; }
    if (c->p.n_ranks > 1 && c->host_pipeline_multi && c->p2p_ready && !c->profiling && !c->exact_multi) {
        // the single-rank pipeline for one rank's slice — positions up, then bounds / keys / sort while masses
        // and velocities are still in flight; integrator + download on the high-priority stream.  Bit-identical to
        // the plain sequence below (tests/multi_gpu_check.py --host-step, profiles/r02_multi_gpu_check_g2.log);
        // 2 GPUs, 1M bodies each: 1.72 -> 1.56 ms per step.
        DeviceGuard g(c->device);
        c->perm_identity = true;     // the call overwrites the rank's whole slice in the caller's order
        const int64_t lo = c->own_lo, cnt = c->own_hi - c->own_lo;
        BH_CUDA_OK(cudaEventRecord(c->ev0, c->stream));
        BH_CUDA_OK(cudaEventRecord(c->ev_fork, c->stream));
        BH_CUDA_OK(cudaStreamWaitEvent(c->copy_stream, c->ev_fork, 0));
        BH_CUDA_OK(cudaMemcpyAsync(c->pos + lo, pos + 2 * lo, sizeof(double2) * cnt, cudaMemcpyHostToDevice, c->copy_stream));
        BH_CUDA_OK(cudaEventRecord(c->ev_up[0], c->copy_stream));
        BH_CUDA_OK(cudaMemcpyAsync(c->mass + lo, mass + lo, sizeof(double) * cnt, cudaMemcpyHostToDevice, c->copy_stream));
        BH_CUDA_OK(cudaEventRecord(c->ev_up[1], c->copy_stream));
        BH_CUDA_OK(cudaMemcpyAsync(c->vel + lo, vel + 2 * lo, sizeof(double2) * cnt, cudaMemcpyHostToDevice, c->copy_stream));
        BH_CUDA_OK(cudaEventRecord(c->ev_vel[0], c->copy_stream));
        c->mass_complete = false;
        BH_CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_up[0], 0));
        BH_TRY(enqueue_build(c, false, nullptr, c->ev_up[1]));
        BH_TRY(enqueue_forces(c, false));
        BH_CUDA_OK(cudaEventRecord(c->ev_trav[0], c->stream));
        BH_CUDA_OK(cudaStreamWaitEvent(c->dl_stream, c->ev_trav[0], 0));
        BH_CUDA_OK(cudaStreamWaitEvent(c->dl_stream, c->ev_vel[0], 0));
        launch_integrate(c->pos, c->vel, c->acc, c->force, c->mass, c->own_lo, c->own_hi, c->p.dt, c->dl_stream);
        BH_TRY(check_launch());
        BH_CUDA_OK(cudaMemcpyAsync(out_pos + 2 * lo, c->pos + lo, sizeof(double2) * cnt, cudaMemcpyDeviceToHost, c->dl_stream));
        BH_CUDA_OK(cudaEventRecord(c->ev_dl, c->dl_stream));
        BH_CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_dl, 0));
        BH_CUDA_OK(cudaEventRecord(c->ev1, c->stream));
        BH_CUDA_OK(cudaStreamSynchronize(c->stream));
        c->bodies_set = true; c->tree_valid = false; c->tree_built = true; c->timed = true;
        return peer_error(c);
    }
    if (c->p.n_ranks > 1) {
        // multi-rank: every rank moves only its own slice both ways (out_pos gets this rank's slice; the
        // other entries are left untouched — one process per GPU owns one slice of the host arrays)
        BH_TRY(bh_set_bodies(c, pos, vel, mass));
        BH_TRY(bh_step(c, 1));
        DeviceGuard g(c->device);
        const int64_t lo = c->own_lo, cnt = c->own_hi - c->own_lo;
        BH_CUDA_OK(cudaMemcpyAsync(out_pos + 2 * lo, c->pos + lo, sizeof(double2) * cnt, cudaMemcpyDeviceToHost, c->stream));
        BH_CUDA_OK(cudaStreamSynchronize(c->stream));
        return peer_error(c);
    }
    if (c->profiling) {
        BH_TRY(bh_set_bodies(c, pos, vel, mass));
        BH_TRY(bh_step(c, 1));
        return bh_get_positions(c, out_pos);
    }
    DeviceGuard g(c->device);
    const int64_t n = c->d.n;
    c->perm_identity = true;     // the call overwrites the whole state in the caller's order
    // exact leaves read other bodies' positions during the walk: chunk k's integrator must not overlap chunk k+1's walk
    const int nch = (c->p.flags & BH_FLAG_EXACT_LEAVES) ? 1 : (int)std::max<int64_t>(1, std::min<int64_t>(c->host_chunks, n / 4096));
    if (nch > 1 && !c->chunk_lists) {
        BH_TRY(dev_alloc(&c->chunk_lists, (size_t)n));
        BH_TRY(dev_alloc(&c->chunk_counts, (size_t)kMaxHostChunks * ((n + 255) / 256)));
    }
    // With pinned host buffers the whole call (3 streams: uploads, compute, integrate + downloads) is captured
    // once into a CUDA graph keyed by the four host pointers and replayed: one launch instead of ~40 API calls.
    const void* key[4] = {pos, vel, mass, out_pos};
    bool use_graph = !(c->p.flags & BH_FLAG_NO_GRAPH) && !c->host_graph_failed && !c->host_trace && host_pinned(pos) && host_pinned(vel) &&
                     host_pinned(mass) && host_pinned(out_pos);
    if (use_graph && (!c->host_graph || memcmp(key, c->host_graph_key, sizeof key) != 0)) {
        if (c->host_graph) { cudaGraphExecDestroy(c->host_graph); c->host_graph = nullptr; }
        cudaGraph_t graph = nullptr;
        const uint64_t before = g_launches;
        int rc = BH_ERR_CUDA;
        if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            rc = enqueue_host_step(c, pos, vel, mass, out_pos, nch);
            cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
            if (rc == BH_OK && e != cudaSuccess) rc = BH_ERR_CUDA;
        }
        c->host_graph_kernels = g_launches - before;
        g_launches = before;   // capture launched nothing
        if (rc == BH_OK && cudaGraphInstantiate(&c->host_graph, graph, 0) != cudaSuccess) rc = BH_ERR_CUDA;
        if (graph) cudaGraphDestroy(graph);
        if (rc != BH_OK) {     // fall back to direct submission for the rest of this context's life
            cudaGetLastError();
            c->host_graph = nullptr;
            c->host_graph_failed = true;
            use_graph = false;
        } else {
            memcpy(c->host_graph_key, key, sizeof key);
        }
    }
    BH_CUDA_OK(cudaEventRecord(c->ev0, c->stream));
    if (use_graph) {
        BH_CUDA_OK(cudaGraphLaunch(c->host_graph, c->stream));
        g_launches += c->host_graph_kernels;
    } else {
        BH_TRY(enqueue_host_step(c, pos, vel, mass, out_pos, nch));
    }
    BH_CUDA_OK(cudaEventRecord(c->ev1, c->stream));
    BH_CUDA_OK(cudaStreamSynchronize(c->stream));
    trace_dump(c);
    c->bodies_set = true; c->tree_valid = false; c->tree_built = true; c->timed = true; c->tree_full = true;
    return BH_OK;
}

int bh_build_tree(bh_ctx* c) {
    if (!c || !c->bodies_set) { set_error("bh_build_tree: no bodies"); return BH_ERR_INVALID; }
    DeviceGuard g(c->device);
    BH_TRY(enqueue_build(c));
    c->tree_valid = true; c->tree_built = true;
    return BH_OK;
}

int bh_compute_forces(bh_ctx* c) {
    if (!c || !c->tree_valid) { set_error("bh_compute_forces: no tree for the current positions (call bh_build_tree; a step moves the bodies)"); return BH_ERR_INVALID; }
    DeviceGuard g(c->device);
    if (c->p.n_ranks > 1 && c->tree_full && !c->exact_multi) BH_TRY(enqueue_build(c));   // diagnostic getters left a full tree behind
    cudaMemsetAsync(c->s.counters, 0, 4 * sizeof(unsigned long long), c->stream);
    return enqueue_forces(c, false);
}

int bh_integrate(bh_ctx* c) {
    if (!c || !c->bodies_set) { set_error("bh_integrate: no bodies"); return BH_ERR_INVALID; }
    DeviceGuard g(c->device);
    launch_integrate(c->pos, c->vel, c->acc, c->force, c->mass, c->own_lo, c->own_hi, c->p.dt, c->stream);
    BH_TRY(check_launch());
    c->tree_valid = false;
    return BH_OK;
}

int bh_synchronize(bh_ctx* c) {
    if (!c) { set_error("null context"); return BH_ERR_INVALID; }
    DeviceGuard g(c->device);
    BH_CUDA_OK(cudaStreamSynchronize(c->stream));
    return peer_error(c);
}

int bh_get_positions(bh_ctx* c, double* out) { if (!c) return BH_ERR_INVALID; return copy_out2(c, c->pos, out, true); }
int bh_get_velocities(bh_ctx* c, double* out) { if (!c) return BH_ERR_INVALID; return copy_out2(c, c->vel, out, true); }
int bh_get_accelerations(bh_ctx* c, double* out) { if (!c) return BH_ERR_INVALID; return copy_out2(c, c->acc, out, true); }
int bh_get_forces(bh_ctx* c, double* out) { if (!c) return BH_ERR_INVALID; return copy_out2(c, c->force, out, true); }

int bh_get_bounds(bh_ctx* c, double out4[4]) {
    if (!c || !c->tree_built) { set_error("bh_get_bounds: no tree built"); return BH_ERR_INVALID; }
    DeviceGuard g(c->device);
    StepConsts h;
    BH_CUDA_OK(cudaMemcpyAsync(&h, c->consts, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    BH_CUDA_OK(cudaStreamSynchronize(c->stream));
    out4[0] = h.xmin; out4[1] = h.xmax; out4[2] = h.ymin; out4[3] = h.ymax;
    return BH_OK;
}

static int fetch_sorted(bh_ctx* c, std::vector<uint32_t>& keys, std::vector<uint32_t>& idx) {
    const int64_t n = c->d.n;
    keys.resize(n); idx.resize(n);
    BH_CUDA_OK(cudaMemcpyAsync(keys.data(), c->keys[c->sorted], 4 * n, cudaMemcpyDeviceToHost, c->stream));
    BH_CUDA_OK(cudaMemcpyAsync(idx.data(), c->idx[c->sorted], 4 * n, cudaMemcpyDeviceToHost, c->stream));
    BH_CUDA_OK(cudaStreamSynchronize(c->stream));
    return BH_OK;
}

// internal -> original index; empty = identity
static int fetch_perm(bh_ctx* c, std::vector<uint32_t>& perm) {
    perm.clear();
    if (c->perm_identity) return BH_OK;
    perm.resize(c->d.n);
    BH_CUDA_OK(cudaMemcpyAsync(perm.data(), c->perm, 4 * c->d.n, cudaMemcpyDeviceToHost, c->stream));
    BH_CUDA_OK(cudaStreamSynchronize(c->stream));
    return BH_OK;
}

int bh_get_body_keys(bh_ctx* c, uint32_t* out) {
    if (!c || !out || !c->tree_valid) { set_error("bh_get_body_keys: no tree for the current positions (call bh_build_tree; a step moves the bodies)"); return BH_ERR_INVALID; }
    DeviceGuard g(c->device);
    BH_TRY(ensure_full_tree(c));
    std::vector<uint32_t> keys, idx, perm;
    BH_TRY(fetch_sorted(c, keys, idx));
    BH_TRY(fetch_perm(c, perm));
    for (int64_t j = 0; j < c->d.n; ++j) out[perm.empty() ? idx[j] : perm[idx[j]]] = keys[j];
    return BH_OK;
}

int bh_get_sorted_order(bh_ctx* c, uint32_t* out) {
    if (!c || !out || !c->tree_valid) { set_error("bh_get_sorted_order: no tree for the current positions (call bh_build_tree; a step moves the bodies)"); return BH_ERR_INVALID; }
    DeviceGuard g(c->device);
    BH_TRY(ensure_full_tree(c));
    std::vector<uint32_t> keys, idx, perm;
    BH_TRY(fetch_sorted(c, keys, idx));
    BH_TRY(fetch_perm(c, perm));
    for (int64_t j = 0; j < c->d.n; ++j) out[j] = perm.empty() ? idx[j] : perm[idx[j]];
    return BH_OK;
}

int bh_get_tree_size(bh_ctx* c, int64_t* n_nodes) {
    if (!c || !n_nodes || !c->tree_built) { set_error("bh_get_tree_size: no tree built"); return BH_ERR_INVALID; }
    DeviceGuard g(c->device);
    unsigned long long v = 0;
    BH_CUDA_OK(cudaMemcpyAsync(&v, c->s.counters + 4, sizeof v, cudaMemcpyDeviceToHost, c->stream));
    BH_CUDA_OK(cudaStreamSynchronize(c->stream));
    *n_nodes = (int64_t)v;
    return BH_OK;
}

static int fetch_host_tree(bh_ctx* c, HostTree& ht) {
    const uint64_t np = c->d.npyramid;
    const int64_t n = c->d.n;
    ht.finest = c->d.finest;
    for (int l = 0; l <= kMaxLevels; ++l) ht.level_off[l] = c->d.level_off[l];
    ht.mass.resize(np); ht.comx.resize(np); ht.comy.resize(np); ht.count.resize(np); ht.first.resize(np);
    ht.sidx.resize(n); ht.pos.resize(2 * n);
    BH_CUDA_OK(cudaMemcpyAsync(ht.mass.data(), c->tree.mass, 8 * np, cudaMemcpyDeviceToHost, c->stream));
    BH_CUDA_OK(cudaMemcpyAsync(ht.comx.data(), c->tree.comx, 8 * np, cudaMemcpyDeviceToHost, c->stream));
    BH_CUDA_OK(cudaMemcpyAsync(ht.comy.data(), c->tree.comy, 8 * np, cudaMemcpyDeviceToHost, c->stream));
    BH_CUDA_OK(cudaMemcpyAsync(ht.count.data(), c->tree.count, 4 * np, cudaMemcpyDeviceToHost, c->stream));
    BH_CUDA_OK(cudaMemcpyAsync(ht.first.data(), c->tree.first, 4 * np, cudaMemcpyDeviceToHost, c->stream));
    BH_CUDA_OK(cudaMemcpyAsync(ht.sidx.data(), c->idx[c->sorted], 4 * n, cudaMemcpyDeviceToHost, c->stream));
    BH_CUDA_OK(cudaMemcpyAsync(ht.pos.data(), c->pos, 16 * n, cudaMemcpyDeviceToHost, c->stream));
    StepConsts h;
    BH_CUDA_OK(cudaMemcpyAsync(&h, c->consts, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    BH_CUDA_OK(cudaStreamSynchronize(c->stream));
    ht.bounds[0] = h.xmin; ht.bounds[1] = h.xmax; ht.bounds[2] = h.ymin; ht.bounds[3] = h.ymax;
    BH_TRY(fetch_perm(c, ht.perm));
    return BH_OK;
}

int bh_get_tree(bh_ctx* c, double* out_rows, int64_t cap_rows, int64_t* n_rows) {
    if (!c || !n_rows || !c->tree_valid) { set_error("bh_get_tree: no tree for the current positions (call bh_build_tree; a step moves the bodies)"); return BH_ERR_INVALID; }
    DeviceGuard g(c->device);
    BH_TRY(ensure_full_tree(c));
    HostTree ht;
    BH_TRY(fetch_host_tree(c, ht));
    *n_rows = canonical_rows(ht, out_rows, cap_rows);
    return BH_OK;
}

int bh_dump_quadtree(bh_ctx* c, const char* path) {
    if (!c || !path || !c->tree_valid) { set_error("bh_dump_quadtree: no tree for the current positions (call bh_build_tree; a step moves the bodies)"); return BH_ERR_INVALID; }
    DeviceGuard g(c->device);
    BH_TRY(ensure_full_tree(c));
    HostTree ht;
    BH_TRY(fetch_host_tree(c, ht));
    return dump_quadtree_txt(ht, path);
}

int bh_get_counters(bh_ctx* c, bh_counters* out) {
    if (!c || !out) { set_error("null argument"); return BH_ERR_INVALID; }
    DeviceGuard g(c->device);
    unsigned long long v[8];
    uint32_t heavy = 0, huge = 0;
    BH_CUDA_OK(cudaMemcpyAsync(v, c->s.counters, sizeof v, cudaMemcpyDeviceToHost, c->stream));
    BH_CUDA_OK(cudaMemcpyAsync(&heavy, c->s.heavy_count, 4, cudaMemcpyDeviceToHost, c->stream));
    BH_CUDA_OK(cudaMemcpyAsync(&huge, c->s.huge_count, 4, cudaMemcpyDeviceToHost, c->stream));
    BH_CUDA_OK(cudaStreamSynchronize(c->stream));
    memset(out, 0, sizeof *out);
    out->interactions = v[0]; out->visits = v[1]; out->opens = v[2]; out->warp_steps = v[3]; out->nodes = v[4];
    out->heavy_cells = (uint64_t)heavy + huge;   // cell_scan_kernel queues the two kinds separately
    out->zero_mass_bodies = c->zero_mass_bodies;
    out->reorders = c->n_reorders;
    return BH_OK;
}

int bh_set_profiling(bh_ctx* c, int32_t on) {
    if (!c) { set_error("null context"); return BH_ERR_INVALID; }
    c->profiling = on != 0;
    return BH_OK;
}

int bh_get_timers(bh_ctx* c, bh_timers* out) {
    if (!c || !out) { set_error("null argument"); return BH_ERR_INVALID; }
    *out = c->timers;
    out->kernel_launches = g_launches - c->launches_base;
    return BH_OK;
}

int bh_reset_timers(bh_ctx* c) {
    if (!c) { set_error("null context"); return BH_ERR_INVALID; }
    memset(&c->timers, 0, sizeof c->timers);
    c->launches_base = g_launches;
    return BH_OK;
}

int bh_last_step_ms(bh_ctx* c, float* ms) {
    if (!c || !ms || !c->timed) { set_error("bh_last_step_ms: no timed step"); return BH_ERR_INVALID; }
    DeviceGuard g(c->device);
    BH_CUDA_OK(cudaEventSynchronize(c->ev1));
    BH_CUDA_OK(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return peer_error(c);
}

int bh_direct_forces(bh_ctx* c, double* out, float* device_ms) {
    if (!c || !c->bodies_set) { set_error("bh_direct_forces: no bodies"); return BH_ERR_INVALID; }
    if (c->p.n_ranks > 1) { set_error("bh_direct_forces: single-rank contexts only (a rank holds only its slice)"); return BH_ERR_INVALID; }
    DeviceGuard g(c->device);
    if (!c->packed) BH_TRY(dev_alloc(&c->packed, (size_t)c->d.n));
    BH_CUDA_OK(cudaEventRecord(c->ev0, c->stream));
    launch_direct(c->pos, c->mass, c->d.n, c->p.G, c->packed, c->force, c->stream);
    BH_CUDA_OK(cudaEventRecord(c->ev1, c->stream));
    BH_TRY(check_launch());
    BH_CUDA_OK(cudaEventSynchronize(c->ev1));
    c->timed = true;
    if (device_ms) BH_CUDA_OK(cudaEventElapsedTime(device_ms, c->ev0, c->ev1));
    if (out) return copy_out2(c, c->force, out, false);
    return BH_OK;
}

int bh_generate(bh_ctx* c, int32_t kind, uint64_t seed) {
    if (!c) { set_error("null context"); return BH_ERR_INVALID; }
    if (kind < BH_GEN_UNIFORM_SQUARE || kind > BH_GEN_PLUMMER_2D) { set_error("unknown generator kind %d", kind); return BH_ERR_INVALID; }
    DeviceGuard g(c->device);
    launch_generate(kind, seed, c->own_lo, c->own_hi, c->pos, c->vel, c->mass, c->stream);
    BH_TRY(check_launch());
    BH_CUDA_OK(cudaStreamSynchronize(c->stream));
    c->mass_complete = c->p.n_ranks == 1;
    c->zero_mass_bodies = 0;                 // log-uniform in [0.1, 0.5]
    c->perm_identity = true;
    c->steps_since_reorder = reorder_period(c) - 1;
    c->bodies_set = true;
    c->tree_valid = false;
    return BH_OK;
}

void bh_philox4x32_10(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]) { philox_host(counter, key, out); }

int bh_generate_host(int32_t kind, uint64_t seed, int64_t first, int64_t count, int32_t r6, double* pos, double* vel,
                     double* mass) {
    if (!pos || !vel || !mass || first < 0 || count < 0) { set_error("bh_generate_host: bad arguments"); return BH_ERR_INVALID; }
    BH_TRY(generate_host(kind, seed, first, first + count, pos, vel, mass));
    if (r6) {
        for (int64_t i = 0; i < 2 * count; ++i) { pos[i] = round6(pos[i]); vel[i] = round6(vel[i]); }
        for (int64_t i = 0; i < count; ++i) mass[i] = round6(mass[i]);
    }
    return BH_OK;
}

int bh_write_init_files(const char* mf, const char* pf, const char* vf, int64_t n, const double* mass, const double* pos,
                        const double* vel) {
    if (!mf || !pf || !vf || !mass || !pos || !vel || n < 0) { set_error("bh_write_init_files: bad arguments"); return BH_ERR_INVALID; }
    return write_init_files(mf, pf, vf, n, mass, pos, vel);
}

int bh_load_text(const char* mf, const char* pf, const char* vf, int64_t n, double* mass, double* pos, double* vel) {
    return load_text(mf, pf, vf, n, mass, pos, vel);
}

int bh_append_positions_txt(const char* path, const double* pos, int64_t n, double time, int truncate) {
    return append_positions_txt(path, pos, n, time, truncate);
}

int bh_trajectory_begin(bh_ctx* c, const char* path, int32_t stride) {
    if (!c || !path || stride < 1) { set_error("bh_trajectory_begin: bad arguments"); return BH_ERR_INVALID; }
    if (c->p.n_ranks > 1) { set_error("bh_trajectory_begin: single-rank contexts only"); return BH_ERR_INVALID; }
    if (c->traj) { set_error("bh_trajectory_begin: a trajectory is already open"); return BH_ERR_INVALID; }
    DeviceGuard g(c->device);
    const size_t n = (size_t)c->d.n;
    for (int k = 0; k < 2; ++k) {
        if (!c->traj_dev[k]) BH_TRY(dev_alloc(&c->traj_dev[k], n));
        if (!c->traj_host[k]) BH_CUDA_OK(cudaMallocHost((void**)&c->traj_host[k], sizeof(double2) * n));
        if (!c->traj_snap[k]) BH_CUDA_OK(cudaEventCreateWithFlags(&c->traj_snap[k], cudaEventDisableTiming));
        if (!c->traj_ready[k]) BH_CUDA_OK(cudaEventCreateWithFlags(&c->traj_ready[k], cudaEventDisableTiming));
    }
    c->traj = new FrameWriter(path, c->d.n, c->device);
    if (!c->traj->ok()) { delete c->traj; c->traj = nullptr; return BH_ERR_IO; }
    c->traj_stride = stride; c->traj_slot = 0; c->traj_calls = 0;
    return BH_OK;
}

int bh_trajectory_record(bh_ctx* c, double time) {
    if (!c || !c->traj) { set_error("bh_trajectory_record: no open trajectory"); return BH_ERR_INVALID; }
    if (!c->bodies_set) { set_error("bh_trajectory_record: no bodies"); return BH_ERR_INVALID; }
    if ((c->traj_calls++ % c->traj_stride) != 0) return BH_OK;
    DeviceGuard g(c->device);
    const int k = c->traj_slot;
    c->traj_slot ^= 1;
    c->traj->acquire(k);        // the frame two records ago has left this buffer (back-pressure if the disk is slower)
    // device snapshot on the compute stream (the next step's integrator overwrites c->pos in place), then the
    // device-to-host copy on the side stream, overlapping the following steps
    if (c->perm_identity) {
        BH_CUDA_OK(cudaMemcpyAsync(c->traj_dev[k], c->pos, sizeof(double2) * c->d.n, cudaMemcpyDeviceToDevice, c->stream));
    } else {                    // resident order -> the caller's body order
        unpermute2_kernel<<<(unsigned)((c->d.n + 255) / 256), 256, 0, c->stream>>>(c->pos, c->perm, c->d.n, c->traj_dev[k]);
        BH_TRY(check_launch());
    }
    BH_CUDA_OK(cudaEventRecord(c->traj_snap[k], c->stream));
    BH_CUDA_OK(cudaStreamWaitEvent(c->copy_stream, c->traj_snap[k], 0));
    BH_CUDA_OK(cudaMemcpyAsync(c->traj_host[k], c->traj_dev[k], sizeof(double2) * c->d.n, cudaMemcpyDeviceToHost, c->copy_stream));
    BH_CUDA_OK(cudaEventRecord(c->traj_ready[k], c->copy_stream));
    c->traj->submit(k, c->traj_host[k], time, (void*)c->traj_ready[k]);
    return BH_OK;
}

int bh_trajectory_end(bh_ctx* c) {
    if (!c || !c->traj) { set_error("bh_trajectory_end: no open trajectory"); return BH_ERR_INVALID; }
    DeviceGuard g(c->device);
    c->traj->finish();
    const bool ok = c->traj->ok();
    delete c->traj;
    c->traj = nullptr;
    for (int k = 0; k < 2; ++k) {
        if (c->traj_dev[k]) { cudaFree(c->traj_dev[k]); c->traj_dev[k] = nullptr; }
        if (c->traj_host[k]) { cudaFreeHost(c->traj_host[k]); c->traj_host[k] = nullptr; }
        if (c->traj_snap[k]) { cudaEventDestroy(c->traj_snap[k]); c->traj_snap[k] = nullptr; }
        if (c->traj_ready[k]) { cudaEventDestroy(c->traj_ready[k]); c->traj_ready[k] = nullptr; }
    }
    if (!ok) { set_error("trajectory file: write failed"); return BH_ERR_IO; }
    return BH_OK;
}

int bh_measure_fp32_peak(int32_t device, double* tflops, double* mhz) {
    if (!tflops) { set_error("null argument"); return BH_ERR_INVALID; }
    return measure_fp32_peak(device, tflops, mhz);
}

}  // extern "C"
