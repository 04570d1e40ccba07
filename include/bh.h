/* bh.h — C-ABI of the B200-native 2-D Barnes-Hut engine (libbh.so).
 *
 * The reference (DavidSevic/gpu-nbody-simulation, implementation/project.cu) has no library
 * or FFI seam: its simulation path is inlined in one executable.  This header is the seam cut
 * through `runSimulationGpu` (project.cu:918-1024); every entry point names the reference
 * code it replaces.  Plain C types only (no torch, no C++ types); all buffers named `host`
 * are caller-owned host memory in the reference's own layouts (AoS double[N][2] for vectors,
 * double[N] for masses, project.cu:38-43), in ORIGINAL body order.
 *
 * Threading: one host thread per context.  Every call returns BH_OK (0) or a negative error
 * code; bh_last_error() returns the message of the last failure on the calling thread.  The
 * reference reads no CUDA return code at all; here every CUDA / NCCL call is checked.
 * There is no CPU fallback: without a CUDA device bh_create fails with BH_ERR_CUDA.
 */
#ifndef BH_H_
#define BH_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BH_ABI_VERSION 1

enum {
    BH_OK = 0,
    BH_ERR_INVALID = -1,  /* bad argument / bad state */
    BH_ERR_CUDA = -2,     /* CUDA runtime error (message holds cudaGetErrorString) */
    BH_ERR_NCCL = -3,     /* NCCL error or libnccl not loadable */
    BH_ERR_IO = -4,       /* file could not be opened / too few lines (project.cu:118-144) */
    BH_ERR_NOMEM = -5
};

/* bh_params.flags */
enum {
    BH_FLAG_FP64_TRAVERSAL = 1u << 0, /* evaluate forces in FP64 with the reference's expression
                                         order (verification mode); default: FP32 arithmetic on a
                                         double-float displacement */
    BH_FLAG_COUNTERS = 1u << 1,       /* keep per-step interaction / visit counters (cheap) */
    BH_FLAG_NO_GRAPH = 1u << 2,       /* launch kernels directly instead of replaying a CUDA graph */
    BH_FLAG_EXACT_EPS = 1u << 3,      /* FP32 traversal: evaluate 1/(d + dist_eps) exactly (2 SFU ops per
                                         interaction) instead of to first order in dist_eps/d (1 SFU op;
                                         relative error (dist_eps/d)^2, < 1e-6 for d > 1e-12) */
    BH_FLAG_EXACT_LEAVES = 1u << 4    /* EXTENSION, not reference behaviour: a multi-body leaf at
                                         the depth cap acts through its bodies one by one (self excluded) instead
                                         of through one monopole that contains the body itself (project.cu:360-382,
                                         :646).  Specified by the oracle's bho_compute_forces_exact_leaves; default
                                         off = reference semantics.  On a multi-rank context every step all-gathers
                                         the positions over NCCL and every rank builds the full tree (a leaf's bodies
                                         may live on other ranks); bh_attach_nccl is required. */
};

/* Runtime copy of the reference's compile-time macros and source-level constants.
 * bh_default_params() fills in the reference's values. */
typedef struct bh_params {
    int64_t n_bodies;     /* -DN_BODIES            project.cu:1-3    (40000) */
    double G;             /* const double G        project.cu:27     (6.67e-11) */
    double dt;            /* DELTA_T               project.cu:29     (1.0) */
    double theta;         /* THETA                 project.cu:60     (0.5) */
    double dist_eps;      /* "+ 1e-15" on distance project.cu:634, :748 */
    double mass_eps;      /* "<= 1e-15" node skip  project.cu:617, :731 */
    double pad_frac;      /* padFraction           project.cu:558    (0.1) */
    double pad_fallback;  /* pad if extent == 0    project.cu:564    (1e-6) */
    int32_t max_depth;    /* QUADTREE_MAX_DEPTH    project.cu:61     (10, root = depth 1); 1..13 */
    int32_t device;       /* CUDA device ordinal; -1 = current device */
    uint32_t flags;       /* BH_FLAG_* */
    int32_t exact_leaf_max; /* finest cells with <= this many bodies accumulate mass/COM with the
                               reference's sequential running average (project.cu:367-373, bit-exact);
                               fuller cells use a fixed-shape parallel sum.  default 64 */
    /* multi-GPU (no reference counterpart): rank r owns the contiguous index range bh_shard_range gives
     * it; every rank keys / sorts / sums only its own bodies, the ranks exchange the bounding box and the
     * per-cell sums (NVLink peer stores, or two NCCL all-reduces), never the bodies. */
    int32_t rank;         /* 0 .. n_ranks-1 */
    int32_t n_ranks;      /* 1 = single GPU */
    int32_t reserved[4];  /* reserved[0]: traversal tuning knob: 0 = default, 1 / 2 = bodies per lane (2 = packed pair
                             kernel), 3 = generic kernel with 2 bodies per lane, 9 = with BH_FLAG_EXACT_LEAVES: the
                             list kernel's member loop */
} bh_params;

typedef struct bh_ctx bh_ctx; /* opaque; owns device memory, stream, CUDA graph, NCCL comm */

/* Per-step work counters with the reference's per-body semantics (project.cu:608-670). */
typedef struct bh_counters {
    uint64_t interactions; /* executions of the force block project.cu:651-658 / :765-772 */
    uint64_t visits;       /* node pops as the reference would count them (per body) */
    uint64_t opens;        /* per-body opened nodes */
    uint64_t warp_steps;   /* warp-level traversal steps actually executed (parent expansions) */
    uint64_t nodes;        /* nodes of the reference-equivalent tree (quadtree.size()) */
    uint64_t heavy_cells;  /* finest cells summed with the parallel path (see exact_leaf_max) */
    uint64_t zero_mass_bodies; /* bodies of mass exactly 0 in the last bh_set_bodies: the reference treats a leaf holding
                                  one as EMPTY and overwrites it (project.cu:393-405), which makes its topology depend
                                  on the insertion order; the engine counts every body (order-independent tree) and
                                  reports the condition here instead of reproducing it */
    uint64_t reorders;     /* physical re-sorts of the body arrays (single rank) / re-partitions (multi rank) since bh_create */
} bh_counters;

/* Accumulated device time per phase in microseconds (cudaEvent, only while profiling is on). */
typedef struct bh_timers {
    double bounds_keys_us, sort_us, build_us, traverse_us, integrate_us, exchange_us, total_us;
    uint64_t steps;        /* steps accumulated */
    uint64_t kernel_launches; /* kernels of this library launched since bh_reset_timers */
} bh_timers;

const char* bh_last_error(void);
int bh_abi_version(void);
void bh_default_params(bh_params* p);

/* ---- life cycle (replaces the cudaMalloc/cudaFree block project.cu:932-940, :1014-1019) ---- */
int bh_create(const bh_params* p, bh_ctx** out);
int bh_destroy(bh_ctx* ctx);
/* multi-GPU: rank 0 calls bh_nccl_unique_id, the host application ships the 128 bytes to every
 * rank (e.g. torch.distributed broadcast), every rank calls bh_attach_nccl before bh_step. */
int bh_nccl_unique_id(void* id128);
int bh_attach_nccl(bh_ctx* ctx, const void* id128);
/* NVLink peer-memory exchange (optional, 2..8 ranks of one box; replaces the per-step NCCL calls by
 * stores into cudaIpc-mapped peer buffers fused with the kernels, and lets multi-rank steps replay as
 * CUDA graphs): every rank calls bh_comm_handle, the host application all-gathers the 64-byte handles
 * (rank order), every rank calls bh_attach_peers with the n_ranks * 64 bytes. */
int bh_comm_handle(bh_ctx* ctx, void* handle64);
int bh_attach_peers(bh_ctx* ctx, const void* handles, int32_t n_handles);
/* index range [lo, hi) of bodies owned by `rank` (pure integer logic, usable without a GPU) */
int bh_shard_range(int64_t n_bodies, int32_t n_ranks, int32_t rank, int64_t* lo, int64_t* hi);

/* ---- body state (replaces the H2D copies project.cu:943-945) ---- */
int bh_set_bodies(bh_ctx* ctx, const double* pos_xy_host, const double* vel_xy_host, const double* mass_host);
int bh_set_positions(bh_ctx* ctx, const double* pos_xy_host);   /* teacher forcing */
int bh_set_velocities(bh_ctx* ctx, const double* vel_xy_host);
/* device-resident snapshot / restore of (pos, vel): lets a benchmark restart every step from the
 * initial distribution without touching the host (the reference physics collapses the tree after
 * one step). */
int bh_snapshot(bh_ctx* ctx);
int bh_restore(bh_ctx* ctx);

/* ---- the hot path ---- */
/* nsteps iterations of the loop body project.cu:955-1011: bounds -> keys -> sort -> tree -> COM ->
 * traversal -> a=F/m, v+=a dt, x+=v dt.  Asynchronous on the context's stream. */
int bh_step(bh_ctx* ctx, int32_t nsteps);
/* same, but every step starts from the snapshot (read out of place; pos / vel receive the result) */
int bh_step_from_snapshot(bh_ctx* ctx, int32_t nsteps);
/* one step with HOST buffers (the reference-facing call: project.cu moves the tree H2D and the
 * positions D2H every step, :968, :1010): uploads pos / mass / vel, runs one step, downloads the new
 * positions into out_pos_xy_host.  Uploads are pipelined against the build; with pinned (page-locked)
 * buffers the whole call is captured once into a CUDA graph keyed by the four pointers and replayed, so
 * keep passing the same buffers.  Synchronous. */
int bh_step_host(bh_ctx* ctx, const double* pos_xy_host, const double* vel_xy_host, const double* mass_host,
                 double* out_pos_xy_host);
/* phase-split entry points for teacher-forced parity */
int bh_build_tree(bh_ctx* ctx);       /* replaces buildTree        project.cu:575-591 (+H2D :968) */
int bh_compute_forces(bh_ctx* ctx);   /* replaces computeForcesGpu project.cu:679-793; needs a built tree */
int bh_integrate(bh_ctx* ctx);        /* replaces updateAccVelPos  project.cu:819-836 */
int bh_synchronize(bh_ctx* ctx);      /* cudaDeviceSynchronize project.cu:989, :1003 */

/* ---- results, ORIGINAL body order (replaces the D2H copy project.cu:1010) ---- */
int bh_get_positions(bh_ctx* ctx, double* out_xy_host);
int bh_get_velocities(bh_ctx* ctx, double* out_xy_host);
int bh_get_accelerations(bh_ctx* ctx, double* out_xy_host);
int bh_get_forces(bh_ctx* ctx, double* out_xy_host);
int bh_get_bounds(bh_ctx* ctx, double out4[4]);                /* xmin xmax ymin ymax, project.cu:536-573 */
int bh_get_body_keys(bh_ctx* ctx, uint32_t* out_keys_host);    /* cell path of DetermineChild, project.cu:348-356 */
int bh_get_sorted_order(bh_ctx* ctx, uint32_t* out_idx_host);  /* body index at every sorted position */

/* Canonical node table of the reference-equivalent tree: DFS pre-order, children 0->3 (the order
 * of TraverseTreeToFile, project.cu:504-534).  Row = 10 doubles { depth (root 0), xmin, xmax,
 * ymin, ymax, mass, comx, comy, PARTICLE_INDEX as the reference stores it (idx | -idx-2 | -1),
 * is_internal }.  bh_get_tree_size returns the row count (== quadtree.size()). */
int bh_get_tree_size(bh_ctx* ctx, int64_t* n_nodes);
int bh_get_tree(bh_ctx* ctx, double* out_rows_host, int64_t cap_rows, int64_t* n_rows);
/* quadtree_{init,final}_gpu.txt writer, format of project.cu:509-526 (reads plot_quadtree.py) */
int bh_dump_quadtree(bh_ctx* ctx, const char* path);

int bh_get_counters(bh_ctx* ctx, bh_counters* out);
int bh_set_profiling(bh_ctx* ctx, int32_t on);   /* per-phase cudaEvent timers (forces NO_GRAPH path) */
int bh_get_timers(bh_ctx* ctx, bh_timers* out);
int bh_reset_timers(bh_ctx* ctx);
/* device time in ms between the start and the end of the last bh_step / bh_step_from_snapshot
 * call, measured with cudaEvents on the context's stream (synchronizes). */
int bh_last_step_ms(bh_ctx* ctx, float* ms);

/* ---- direct all-pairs kernel (BASELINE config 5; formula of main_approach_1.cpp:53-75) ---- */
int bh_direct_forces(bh_ctx* ctx, double* out_xy_host /* may be NULL */, float* device_ms);

/* ---- seeded initial conditions (replaces initializeGpu / initializeCpu, project.cu:298-341, whose cuRAND
 * states are seeded with time(0)): counter-based Philox4x32-10, body i is a pure function of (seed, i);
 * value ranges of project.cu:30-35; the disk and the projected Plummer sphere are BASELINE configs 2-4. ---- */
enum { BH_GEN_UNIFORM_SQUARE = 0, BH_GEN_UNIFORM_DISK = 1, BH_GEN_PLUMMER_2D = 2 };
/* fills the context's bodies on the device (a rank of a multi-rank context generates only its slice) */
int bh_generate(bh_ctx* ctx, int32_t kind, uint64_t seed);
/* the same bodies [first, first + count) into host arrays; needs no GPU.  round6 != 0 passes every value
 * through the reference writers' text format (6 significant digits) so that arrays and files agree. */
int bh_generate_host(int32_t kind, uint64_t seed, int64_t first, int64_t count, int32_t round6,
                     double* pos_xy_host, double* vel_xy_host, double* mass_host);
/* the raw generator (Philox4x32-10, Random123 conventions), for known-answer tests and reproducibility */
void bh_philox4x32_10(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]);

/* ---- text formats of the reference ---- */
/* loadSimulationDataFromText project.cu:103-161: first n lines of each file */
/* writers of the three initial-condition files, format of project.cu:236-246, :268-281 (default ostream
 * formatting = "%g" with 6 significant digits; masses one per line, vectors "x y" per line) */
int bh_write_init_files(const char* masses_file, const char* positions_file, const char* velocities_file,
                        int64_t n_bodies, const double* mass_host, const double* pos_xy_host,
                        const double* vel_xy_host);
int bh_load_text(const char* masses_file, const char* positions_file, const char* velocities_file,
                 int64_t n, double* mass_out, double* pos_out, double* vel_out);
/* savePositions project.cu:855-863: appends "time i x y \n" (std::to_string, 6 decimals) */
int bh_append_positions_txt(const char* path, const double* pos_xy_host, int64_t n, double time, int truncate);

/* ---- trajectory output: the file plot_2d.py reads, format of savePositions (project.cu:855-863:
 * one "time i x y \n" line per body and frame, std::to_string formatting).  The reference's CPU programs rebuild and
 * append the whole file synchronously; here ONE file stays open, every `stride`-th recorded state is snapshotted on
 * the device (10 us at 1M bodies), copied to one of two pinned host buffers on a side stream and formatted by a
 * background thread while the simulation keeps stepping (at 1M bodies x 1000 steps an unstrided file would be 40 GB,
 * SURVEY H7).  Single-rank contexts. ---- */
int bh_trajectory_begin(bh_ctx* ctx, const char* path, int32_t stride);
/* call once for the initial state (time 0, project.cu:879) and once after every step (project.cu:907): the 1st,
 * (1 + stride)-th, ... calls are written; returns as soon as the copies are enqueued */
int bh_trajectory_record(bh_ctx* ctx, double time);
int bh_trajectory_end(bh_ctx* ctx);   /* waits until every recorded frame is on disk, closes the file */

/* ---- measurement helpers ---- */
/* FP32 FMA peak of the device measured with a register-resident FMA loop, in TFLOP/s */
int bh_measure_fp32_peak(int32_t device, double* tflops, double* sm_clock_mhz_est);

#ifdef __cplusplus
}
#endif
#endif /* BH_H_ */
