// Kernel 3: quadtree build + upward mass / centre-of-mass pass.
//
// Replaces buildTree = InitializeRoot + N x QuadInsert + ComputeMass (project.cu:343-346,
// :358-453, :473-502, :575-591), which the reference runs sequentially on the host every step.
//
// Representation.  With the depth cap (QUADTREE_MAX_DEPTH, root = depth 1) every node of the
// reference's PR quadtree is a cell of the regular pyramid: level l (0 = root .. F = cap-1) is the
// 2^l x 2^l grid produced by l FP64 bisections of the padded root box, and a cell's Morton code is
// the concatenation of DetermineChild results (x = low bit).  The reference tree is exactly the
// subset { cell : every proper ancestor holds >= 2 bodies }: a node is internal iff it holds >= 2
// bodies and lies above the cap level, and a split always materialises all four children
// (project.cu:410-434).  So the tree is stored as the dense pyramid (4^(F+1)-1)/3 cells (349 525 at
// the default cap, 11 MB of 32-byte records: resident in B200's 126 MB L2) with NO child pointers;
// existence is implied by the parent's body count.  Topology is therefore a pure function of the
// sorted cell keys (Karras-style: a finest cell's run is delimited where adjacent sorted keys
// differ; ancestors' runs are unions of their children's).
//
// Arithmetic.  Mass / COM follow the reference operation by operation, without FMA contraction
// (the reference builds the tree in x86-64 host code):
//   * cap-level cell: running weighted average in ascending body index (project.cu:367-373);
//   * single body above the cap: the body's own mass and position (project.cu:400-403);
//   * internal cell: children 0..3, sums started from 0.0, then one division (project.cu:480-495).
// Hence the node table is bit-identical to the reference's, except finest cells holding more than
// `exact_leaf_max` bodies, whose (inherently sequential) running average is replaced by a
// fixed-shape parallel sum (relative difference ~1e-16; deterministic, identical on every rank).
#include "bh_internal.h"
#include "peer_comm.cuh"

namespace bh {

namespace {

struct Cell {
    double m, cx, cy;
    uint32_t cnt, first;
};

// Combine four children (in order 0..3) into their parent.  `single body` parents take the body
// itself; internal parents follow ComputeMass.
// When the parent is internal, every child holding exactly one body is that body's own leaf: its
// pyramid index goes to self_node[body] (the traversal's self-interaction test, project.cu:646/:760).
// SHARDED (multi-GPU): `cnt` is the GLOBAL body count of the cell (all ranks), `first` the sorted
// position of this rank's first body inside it (kNoFirst if the rank has none there); node values
// come from all-reduced sums, so a lone body's cell copies its only non-empty child instead of
// looking the body up (which another rank may own).
constexpr uint32_t kNoFirst = 0xffffffffu;

template <bool SHARDED>
__device__ __forceinline__ Cell combine4(const Cell c[4], const uint32_t* __restrict__ sidx,
                                         const double2* __restrict__ pos, const double* __restrict__ mass,
                                         uint32_t* __restrict__ self_node, uint64_t child_base, bool write_self) {
    Cell p;
    p.cnt = c[0].cnt + c[1].cnt + c[2].cnt + c[3].cnt;
    p.first = c[0].first != kNoFirst ? c[0].first : c[1].first != kNoFirst ? c[1].first
            : c[2].first != kNoFirst ? c[2].first : c[3].first;
    if (p.cnt == 0) {
        p.m = 0.0; p.cx = 0.0; p.cy = 0.0; p.first = kNoFirst;
    } else if (p.cnt == 1) {                       // leaf above the cap: project.cu:400-403
        if constexpr (SHARDED) {
            const int q = c[0].cnt ? 0 : c[1].cnt ? 1 : c[2].cnt ? 2 : 3;
            p.m = c[q].m; p.cx = c[q].cx; p.cy = c[q].cy;
        } else {
            uint32_t b = sidx[p.first];
            double2 x = pos[b];
            p.m = mass[b]; p.cx = x.x; p.cy = x.y;
        }
    } else {                                       // project.cu:480-499
        double tm = 0.0, sx = 0.0, sy = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            tm = __dadd_rn(tm, c[q].m);
            sx = __dadd_rn(sx, __dmul_rn(c[q].m, c[q].cx));
            sy = __dadd_rn(sy, __dmul_rn(c[q].m, c[q].cy));
        }
        if (tm > 0.0) { sx = __ddiv_rn(sx, tm); sy = __ddiv_rn(sy, tm); }
        p.m = tm; p.cx = sx; p.cy = sy;
        if (write_self) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (c[q].cnt == 1u && c[q].first != kNoFirst) self_node[sidx[c[q].first]] = (uint32_t)(child_base + q);
        }
    }
    return p;
}

// G here is G * scale^2 and `scale` the power-of-two coordinate normalisation (StepConsts::scale).
// The traversal record folds the reference's per-node tests into two numbers:
//   thr : the lane accepts iff !(d2 <= thr).  Leaves (and zero-mass nodes, which the reference skips
//         before any test, project.cu:617) get thr = -1, i.e. "always accept, never open";
//   gm  : G M; zero-mass nodes carry gm = 0 and a far-away COM so they add exactly 0.
constexpr float kFarAway = 1.152921504606847e18f;   // 2^60 in scaled units

__device__ __forceinline__ void store_cell(const TreeArrays& t, uint64_t at, const Cell& c, int level, int finest,
                                           double G, double mass_eps, double scale, float thr2_level,
                                           const uint32_t* __restrict__ sidx) {
    t.mass[at] = c.m; t.comx[at] = c.cx; t.comy[at] = c.cy;
    t.count[at] = c.cnt; t.first[at] = c.first;
    const bool nz = c.m > mass_eps, leaf = (c.cnt <= 1u || level == finest);
    t.flags[at] = (nz ? kNodeNonZero : 0u) | (leaf ? kNodeLeaf : 0u) | (c.cnt == 1u ? kNodeSingle : 0u);
    NodeRec r;
    if (nz) {
        const double sx = c.cx * scale, sy = c.cy * scale;
        r.chx = (float)sx; r.chy = (float)sy;
        r.clx = (float)(sx - (double)r.chx); r.cly = (float)(sy - (double)r.chy);
        r.gm = (float)(G * c.m);
        r.thr = leaf ? -1.0f : thr2_level;
    } else {
        r.chx = kFarAway; r.chy = kFarAway; r.clx = 0.f; r.cly = 0.f; r.gm = 0.f; r.thr = -1.0f;
    }
    r.count = c.cnt; r.first = c.first;
    t.rec[at] = r;
    if (level == 0 && c.cnt == 1u && c.first != kNoFirst) t.self_node[sidx[c.first]] = 0u;   // a lone body: the root is its leaf
}

// ---- finest-cell runs: exclusive scan of the per-cell body counts ------------------------------------------------
// The key kernel counts the bodies of every finest cell (bounds_keys.cu); a cell's run in the sorted order starts at
// the number of bodies in all cells with a smaller key.  One pass with decoupled look-back over 4096-cell tiles
// (atomic ticket = tile order, so every tile a block waits on is resident).  Round 1 found the runs from the sorted
// keys (one thread per BODY, binary searches, 20 us at 1M bodies and 157 us at 16M); this is one thread per 16 CELLS,
// 262 144 cells at the default cap whatever N.  Over-full cells are queued here too (one atomic per warp and queue).
#define kScanAgg (1u << 30)
#define kScanIncl (2u << 30)
#define kScanVal ((1u << 30) - 1u)
constexpr int kScanItems = 16;

__global__ void __launch_bounds__(256)
cell_scan_kernel(const uint32_t* __restrict__ cnt_f, uint32_t* __restrict__ first_f, uint64_t ncells,
                 uint32_t exact_leaf_max, uint32_t* __restrict__ heavy_list, uint32_t* __restrict__ heavy_count,
                 uint32_t* __restrict__ huge_list, uint32_t* __restrict__ huge_count, uint32_t* state, uint32_t* ticket) {
    __shared__ uint32_t s_tile, s_prev;
    __shared__ uint32_t s_warp[8];
    pdl_entry();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t c0 = (uint64_t)tile * (256 * kScanItems) + (uint64_t)tid * kScanItems;
    uint32_t cnt[kScanItems];
    uint32_t local = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {        // (the finest level starts at an odd offset of the pyramid: scalar loads)
        cnt[k] = (c0 + k < ncells) ? __ldg(cnt_f + c0 + k) : 0u;
        local += cnt[k];
    }
    uint32_t inc = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { woff += (w < warp) ? s_warp[w] : 0u; total += s_warp[w]; }
    if (warp == 0) {
        // publish this tile's count, then add up the tiles before it (window of 32 states, nearest first)
        volatile uint32_t* st = state;
        if (lane == 0) st[tile] = (tile == 0 ? kScanIncl : kScanAgg) | total;
        uint32_t prev = 0;
        int64_t t = (int64_t)tile - 1;
        while (t >= 0) {
            const int64_t i = t - lane;
            const uint32_t v = i >= 0 ? st[i] : kScanIncl;            // before tile 0: an inclusive prefix of 0
            const uint32_t not_ready = __ballot_sync(0xffffffffu, v == 0u);
            const uint32_t incl = __ballot_sync(0xffffffffu, (v >> 30) == 2u);
            const int first_incl = incl ? __ffs(incl) - 1 : 32;
            if (not_ready & ((first_incl >= 31) ? 0xffffffffu : ((2u << first_incl) - 1u))) continue;   // not published yet: look again
            uint32_t x = (lane <= first_incl) ? (v & kScanVal) : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            prev += x;
            if (incl) break;
            t -= 32;
        }
        if (lane == 0) {
            if (tile > 0) { __threadfence(); st[tile] = kScanIncl | (prev + total); }
            s_prev = prev;
        }
    }
    __syncthreads();
    uint32_t run = s_prev + woff + inc - local;       // sorted position of the first body of this thread's first cell
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const uint64_t c = c0 + k;
        const uint32_t n_c = cnt[k];
        if (c < ncells) first_f[c] = run;
        run += n_c;
        const bool over = n_c > exact_leaf_max, huge = over && n_c > kHugeCellMin, heavy = over && !huge;
        const uint32_t mh = __ballot_sync(0xffffffffu, heavy), mg = __ballot_sync(0xffffffffu, huge);
        const uint32_t lt = (1u << lane) - 1u;
        if (mh) {
            uint32_t base = 0;
            if (lane == __ffs(mh) - 1) base = atomicAdd(heavy_count, (uint32_t)__popc(mh));
            base = __shfl_sync(0xffffffffu, base, __ffs(mh) - 1);
            if (heavy) heavy_list[base + __popc(mh & lt)] = (uint32_t)c;      // one warp each
        }
        if (mg) {
            uint32_t base = 0;
            if (lane == __ffs(mg) - 1) base = atomicAdd(huge_count, (uint32_t)__popc(mg));
            base = __shfl_sync(0xffffffffu, base, __ffs(mg) - 1);
            if (huge) huge_list[base + __popc(mg & lt)] = (uint32_t)c;        // summed by kHugeParts blocks each
        }
    }
}

// ---- parallel (non-sequential) summation for very full finest cells -------------------------------
// heavy_cells_kernel: one block per queued cell (fixed reduction shape => deterministic, identical
// on all ranks); cells above kHugeCellMin bodies (the collapsed regime of the reference's own
// dynamics: almost all bodies in a handful of cells) are re-queued for huge_cells_kernel, where
// kHugeParts blocks sum fixed, contiguous parts of the run and the last block to finish combines
// the parts in part order (atomic ticket) — still a fixed summation order.
__device__ __forceinline__ void block_sum3(double& m, double& sx, double& sy, double (*sm)[256]) {
    sm[0][threadIdx.x] = m; sm[1][threadIdx.x] = sx; sm[2][threadIdx.x] = sy;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            sm[0][threadIdx.x] += sm[0][threadIdx.x + o];
            sm[1][threadIdx.x] += sm[1][threadIdx.x + o];
            sm[2][threadIdx.x] += sm[2][threadIdx.x + o];
        }
        __syncthreads();
    }
    m = sm[0][0]; sx = sm[1][0]; sy = sm[2][0];
    __syncthreads();
}

__device__ __forceinline__ void write_cell_sums(double* m_f, double* cx_f, double* cy_f, uint32_t cell, double tm,
                                                double sx, double sy, bool raw_sums) {
    m_f[cell] = tm;
    if (raw_sums) { cx_f[cell] = sx; cy_f[cell] = sy; }   // sharded: divided after the all-reduce
    else { cx_f[cell] = tm > 0.0 ? sx / tm : 0.0; cy_f[cell] = tm > 0.0 ? sy / tm : 0.0; }
}

constexpr int kHeavyBlocks = 148 * 4;     // blocks [0, kHeavyBlocks) of heavy_huge_kernel sum heavy cells, one cell at a time
constexpr int kHugeSlots = 16;            // the remaining kHugeParts x kHugeSlots blocks sum huge cells, kHugeParts blocks per cell

// One WARP per heavy cell (lane-strided partial sums, then a xor-butterfly: a fixed shape for a given body count, so the
// sums are deterministic and identical on every rank).  A whole block per cell (round 1) left 224 of 256 threads idle on
// the typical heavy cell of ~70-100 bodies: at 16M bodies, where a third of the finest cells is heavy, that kernel alone
// took 695 us.
__device__ __forceinline__ void heavy_cells(uint32_t block, uint32_t nblocks,
                   const uint32_t* __restrict__ heavy_list, const uint32_t* __restrict__ heavy_count,
                   const uint32_t* __restrict__ cnt_f, const uint32_t* __restrict__ first_f,
                   const uint32_t* __restrict__ sidx, const double2* __restrict__ pos,
                   const double* __restrict__ mass, double* __restrict__ m_f, double* __restrict__ cx_f,
                   double* __restrict__ cy_f, bool raw_sums) {
    const uint32_t nheavy = *heavy_count;
    const uint32_t lane = threadIdx.x & 31u, warps_per_block = blockDim.x >> 5;
    for (uint32_t h = block * warps_per_block + (threadIdx.x >> 5); h < nheavy; h += nblocks * warps_per_block) {
        const uint32_t cell = heavy_list[h];
        const uint32_t c = cnt_f[cell], f = first_f[cell];
        double m = 0.0, sx = 0.0, sy = 0.0;
        for (uint32_t i0 = lane; i0 < c; i0 += 128) {      // four independent gathers in flight per lane
            uint32_t b[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) b[k] = (i0 + 32 * k < c) ? __ldg(sidx + f + i0 + 32 * k) : 0xffffffffu;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (b[k] != 0xffffffffu) {
                    const double mb = __ldg(mass + b[k]);
                    const double2 x = __ldg(pos + b[k]);
                    m += mb; sx += mb * x.x; sy += mb * x.y;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m += __shfl_xor_sync(0xffffffffu, m, o);
            sx += __shfl_xor_sync(0xffffffffu, sx, o);
            sy += __shfl_xor_sync(0xffffffffu, sy, o);
        }
        if (lane == 0) write_cell_sums(m_f, cx_f, cy_f, cell, m, sx, sy, raw_sums);
    }
}

// block (p, slot) sums part p of huge cell slot, slot + kHugeSlots, ...
__device__ __forceinline__ void huge_cells(uint32_t p, uint32_t slot, double (*sm)[256], bool& last,
                  const uint32_t* __restrict__ huge_list, const uint32_t* __restrict__ huge_count,
                  const uint32_t* __restrict__ cnt_f, const uint32_t* __restrict__ first_f,
                  const uint32_t* __restrict__ sidx, const double2* __restrict__ pos,
                  const double* __restrict__ mass, double* __restrict__ m_f, double* __restrict__ cx_f,
                  double* __restrict__ cy_f, bool raw_sums, double* __restrict__ partial,
                  uint32_t* __restrict__ tickets) {
    const uint32_t nhuge = *huge_count;
    for (uint32_t h = slot; h < nhuge; h += kHugeSlots) {
        const uint32_t cell = huge_list[h];
        const uint32_t c = cnt_f[cell], f = first_f[cell];
        const uint32_t per = (c + kHugeParts - 1) / kHugeParts;
        const uint32_t lo = p * per, hi = min(c, lo + per);
        double m = 0.0, sx = 0.0, sy = 0.0;
        for (uint32_t i = lo + threadIdx.x; i < hi; i += 256) {
            const uint32_t b = sidx[f + i];
            const double mb = mass[b];
            const double2 x = pos[b];
            m += mb; sx += mb * x.x; sy += mb * x.y;
        }
        block_sum3(m, sx, sy, sm);
        if (threadIdx.x == 0) {
            double* out = partial + ((size_t)h * kHugeParts + p) * 3;
            out[0] = m; out[1] = sx; out[2] = sy;
            __threadfence();
            last = (atomicAdd(&tickets[h], 1u) == kHugeParts - 1);
        }
        __syncthreads();
        if (last && threadIdx.x == 0) {
            __threadfence();
            double tm = 0.0, tx = 0.0, ty = 0.0;
            for (int q = 0; q < kHugeParts; ++q) {   // fixed order
                const volatile double* in = partial + ((size_t)h * kHugeParts + q) * 3;
                tm += in[0]; tx += in[1]; ty += in[2];
            }
            write_cell_sums(m_f, cx_f, cy_f, cell, tm, tx, ty, raw_sums);
        }
        __syncthreads();
    }
}

// One launch for both kinds of over-full finest cells (cell_scan_kernel queues them separately).
__global__ void __launch_bounds__(256)
heavy_huge_kernel(const uint32_t* __restrict__ heavy_list, const uint32_t* __restrict__ heavy_count,
                  const uint32_t* __restrict__ huge_list, const uint32_t* __restrict__ huge_count,
                  const uint32_t* __restrict__ cnt_f, const uint32_t* __restrict__ first_f,
                  const uint32_t* __restrict__ sidx, const double2* __restrict__ pos,
                  const double* __restrict__ mass, double* __restrict__ m_f, double* __restrict__ cx_f,
                  double* __restrict__ cy_f, bool raw_sums, double* __restrict__ partial,
                  uint32_t* __restrict__ tickets) {
    __shared__ double sm[3][256];
    __shared__ bool last;
    pdl_entry();
    if (blockIdx.x < kHeavyBlocks) {
        heavy_cells(blockIdx.x, kHeavyBlocks, heavy_list, heavy_count, cnt_f, first_f, sidx, pos, mass, m_f, cx_f, cy_f,
                    raw_sums);
    } else {
        const uint32_t b = blockIdx.x - kHeavyBlocks;
        huge_cells(b % kHugeParts, b / kHugeParts, sm, last, huge_list, huge_count, cnt_f, first_f, sidx, pos, mass, m_f,
                   cx_f, cy_f, raw_sums, partial, tickets);
    }
}

// ---- sharded build: this rank's partial sums per finest cell (count, m, m x, m y) ------------------
// sums = [4][ncells] doubles, all-reduced over the ranks before the level pass.  Cells queued as
// heavy were already summed (raw) by heavy_cells_kernel.
// PUSH (peer-memory exchange): the sums of every non-empty cell also go into this rank's slot of every rank's inbox
// (its own included); the last block to finish raises the flag of the step at every peer.
template <bool PUSH>
__global__ void __launch_bounds__(256)
cell_partial_kernel(const uint32_t* __restrict__ cnt_f, const uint32_t* __restrict__ first_f, uint64_t ncells,
                    const uint32_t* __restrict__ sidx, const double2* __restrict__ pos,
                    const double* __restrict__ mass, uint32_t exact_leaf_max, double* __restrict__ sums,
                    const __grid_constant__ PeerComm pc) {
    const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncells) {
        const uint32_t cnt = cnt_f[c];
        double m = 0.0, sx = 0.0, sy = 0.0;
        if (cnt > exact_leaf_max) {                    // m, mx, my written by heavy_huge_kernel
            if constexpr (PUSH) { m = sums[ncells + c]; sx = sums[2 * ncells + c]; sy = sums[3 * ncells + c]; }
        } else {
            const uint32_t f = cnt ? first_f[c] : 0u;
            for (uint32_t i = 0; i < cnt; ++i) {
                const uint32_t b = __ldg(sidx + f + i);
                const double mb = __ldg(mass + b);
                const double2 x = __ldg(pos + b);
                m = __dadd_rn(m, mb);
                sx = __dadd_rn(sx, __dmul_rn(mb, x.x));
                sy = __dadd_rn(sy, __dmul_rn(mb, x.y));
            }
            if constexpr (!PUSH) { sums[ncells + c] = m; sums[2 * ncells + c] = sx; sums[3 * ncells + c] = sy; }
        }
        if constexpr (!PUSH) sums[c] = (double)cnt;
        if constexpr (PUSH) {
            if (cnt) {
                for (int r = 0; r < pc.n_ranks; ++r) {
                    double* in = reinterpret_cast<double*>(pc.peer_base[r] + pc.off_inbox) + (uint64_t)pc.rank * 4 * ncells + c;
                    in[0] = (double)cnt; in[ncells] = m; in[2 * ncells] = sx; in[3 * ncells] = sy;
                }
            }
        }
    }
    if constexpr (PUSH) {
        __shared__ bool last;
        uint32_t* ticket = reinterpret_cast<uint32_t*>(pc.peer_base[pc.rank] + pc.off_err) + 2;
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
        __syncthreads();
        if (last) {
            if (threadIdx.x == 0) *ticket = 0;
            const uint32_t seq = *(reinterpret_cast<const uint32_t*>(pc.peer_base[pc.rank] + pc.off_err) + 1);
            __threadfence_system();
            if ((int)threadIdx.x < pc.n_ranks)
                st_release_sys(reinterpret_cast<uint32_t*>(pc.peer_base[threadIdx.x] + pc.off_in_flag) + pc.rank, seq);
        }
    }
}

// ---- top levels (F-5 .. 0): run by the LAST block of tree_bottom_kernel to finish (atomic ticket), level by level
// through global memory — 341 cells at the default cap; a separate single-block launch cost 8 us.
template <bool SHARDED>
__device__ __forceinline__ void tree_top_levels(const TreeArrays& t, const Dims& d, int top_level,
                                                const uint32_t* __restrict__ sidx, const double2* __restrict__ pos,
                                                const double* __restrict__ mass, double G, double mass_eps, double scale,
                                                unsigned long long* __restrict__ counters,
                                                const StepConsts* __restrict__ consts, uint32_t* s_int) {
    const int F = d.finest;
    if (threadIdx.x == 0) *s_int = 0;
    __syncthreads();
    uint32_t n_internal = 0;
    for (int level = top_level; level >= 0; --level) {
        uint64_t ncells = 1ull << (2 * level);
        uint64_t off = d.level_off[level], offc = d.level_off[level + 1];
        for (uint64_t c = threadIdx.x; c < ncells; c += blockDim.x) {
            Cell ch[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint64_t a = offc + 4 * c + q;
                ch[q].m = __ldcg(t.mass + a); ch[q].cx = __ldcg(t.comx + a); ch[q].cy = __ldcg(t.comy + a);
                ch[q].cnt = __ldcg(t.count + a); ch[q].first = __ldcg(t.first + a);
            }
            Cell up = combine4<SHARDED>(ch, sidx, pos, mass, t.self_node, offc + 4 * c, true);
            store_cell(t, off + c, up, level, F, G, mass_eps, scale, consts->thr2[level], sidx);
            n_internal += (up.cnt >= 2u);
        }
        __threadfence();
        __syncthreads();   // level `level` complete and visible to the block
    }
    if (n_internal) atomicAdd(s_int, n_internal);
    __syncthreads();
    if (threadIdx.x == 0) {
        // counters[5] holds the internal cells counted by all blocks' bottom levels
        unsigned long long internal = *reinterpret_cast<volatile unsigned long long*>(counters + 5) + *s_int;
        counters[4] = 1ull + 4ull * internal;   // quadtree.size() of the reference
    }
}

// ---- bottom kernel: finest cells + up to five levels above -----------------------------------------
// Block b owns finest cells [256 b, 256 b + 256) = one subtree rooted four levels up; thread t
// owns ONE finest cell (maximum parallelism for the latency-bound leaf pass).  Levels F-1 and F-2
// are combined inside a warp with shuffles (4, then 16 lanes per parent), F-3 and F-4 through shared
// memory.  Children are always combined in the order 0..3 (ComputeMass, project.cu:483-490).
constexpr int kBottomThreads = 256;    // 256 finest cells per block: levels F .. F-4

__device__ __forceinline__ Cell shfl_cell(const Cell& c, int src_lane) {
    Cell r;
    r.m = __shfl_sync(0xffffffffu, c.m, src_lane);
    r.cx = __shfl_sync(0xffffffffu, c.cx, src_lane);
    r.cy = __shfl_sync(0xffffffffu, c.cy, src_lane);
    r.cnt = __shfl_sync(0xffffffffu, c.cnt, src_lane);
    r.first = __shfl_sync(0xffffffffu, c.first, src_lane);
    return r;
}

template <bool SHARDED>
__global__ void __launch_bounds__(kBottomThreads)
tree_bottom_kernel(TreeArrays t, Dims d, const uint32_t* __restrict__ sidx, const double2* __restrict__ pos,
                   const double* __restrict__ mass, double G0, double mass_eps, uint32_t exact_leaf_max,
                   unsigned long long* __restrict__ counters, const StepConsts* __restrict__ consts,
                   const double* __restrict__ sums, uint32_t* __restrict__ done_ticket, const __grid_constant__ PeerComm pc) {
    pdl_entry();
    const int F = d.finest;
    if constexpr (SHARDED) {
        if (pc.n_ranks > 1) {   // peer-memory exchange: every rank's contributions of this step must have landed in the inbox
            uint8_t* own = pc.peer_base[pc.rank];
            if ((int)threadIdx.x < pc.n_ranks)
                wait_flag(reinterpret_cast<const uint32_t*>(own + pc.off_in_flag) + threadIdx.x,
                          *(reinterpret_cast<const uint32_t*>(own + pc.off_err) + 1), reinterpret_cast<uint32_t*>(own + pc.off_err),
                          pc.timeout_ns);
            __syncthreads();
        }
    }
    const double scale = consts->scale;
    const double G = G0 * scale * scale;
    const int tid = threadIdx.x, lane = tid & 31;
    const uint64_t offF = d.level_off[F];
    __shared__ Cell s_a[kBottomThreads / 16];     // level F-2 results (16 per block)
    __shared__ Cell s_b[kBottomThreads / 64];     // level F-3 results (4 per block)
    __shared__ Cell s_c[1];
    __shared__ uint32_t s_internal[kBottomThreads / 32];
    __shared__ uint32_t s_last;
    uint32_t n_internal = 0;

    // ---- level F: one cell per thread
    uint64_t code = (uint64_t)blockIdx.x * kBottomThreads + tid;
    uint64_t ncells = d.ncells_finest;
    Cell cur; cur.m = 0.0; cur.cx = 0.0; cur.cy = 0.0; cur.cnt = 0; cur.first = kNoFirst;
    if (SHARDED && code < ncells) {
        // global sums of all ranks; `first` stays this rank's own run (self_node bookkeeping)
        const uint32_t lcnt = t.count[offF + code];
        if (lcnt) cur.first = t.first[offF + code];
        double tc, m, sx, sy;
        if (pc.n_ranks > 1) {   // add the ranks' contributions in rank order (identical on every rank); zero what was consumed
            tc = 0.0; m = 0.0; sx = 0.0; sy = 0.0;
            double* in = reinterpret_cast<double*>(pc.peer_base[pc.rank] + pc.off_inbox) + code;
            for (int r = 0; r < pc.n_ranks; ++r, in += 4 * ncells) {
                const double cr = __ldcg(in);
                if (cr != 0.0) {
                    tc += cr; m += __ldcg(in + ncells); sx += __ldcg(in + 2 * ncells); sy += __ldcg(in + 3 * ncells);
                    in[0] = 0.0; in[ncells] = 0.0; in[2 * ncells] = 0.0; in[3 * ncells] = 0.0;
                }
            }
        } else {                // NCCL fallback: `sums` holds the all-reduced values
            tc = sums[code]; m = sums[ncells + code]; sx = sums[2 * ncells + code]; sy = sums[3 * ncells + code];
        }
        cur.cnt = (uint32_t)llrint(tc);
        cur.m = m;
        cur.cx = m > 0.0 ? __ddiv_rn(sx, m) : 0.0;
        cur.cy = m > 0.0 ? __ddiv_rn(sy, m) : 0.0;
        store_cell(t, offF + code, cur, F, F, G, mass_eps, scale, consts->thr2[F], sidx);
    } else if (code < ncells) {
        cur.cnt = t.count[offF + code];
        if (cur.cnt) {
            cur.first = t.first[offF + code];
            if (cur.cnt <= exact_leaf_max) {
                // project.cu:367-373: running weighted average in ascending body index
                // The arithmetic chain is sequential by definition; the gathers feeding it are not: bodies are
                // fetched four at a time (index loads, then mass / position loads, all independent) so that the
                // dependent L2 round trips overlap instead of adding up per body.
                double em = 0.0, ex = 0.0, ey = 0.0;
                for (uint32_t i0 = 0; i0 < cur.cnt; i0 += 4) {
                    uint32_t b[4];
                    double mb[4];
                    double2 x[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) b[k] = (i0 + k < cur.cnt) ? __ldg(sidx + cur.first + i0 + k) : 0u;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (i0 + k < cur.cnt) { mb[k] = __ldg(mass + b[k]); x[k] = __ldg(pos + b[k]); }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (i0 + k < cur.cnt) {
                            const double tot = __dadd_rn(em, mb[k]);
                            ex = __ddiv_rn(__dadd_rn(__dmul_rn(em, ex), __dmul_rn(mb[k], x[k].x)), tot);
                            ey = __ddiv_rn(__dadd_rn(__dmul_rn(em, ey), __dmul_rn(mb[k], x[k].y)), tot);
                            em = tot;              // node[TOTAL_MASS] += mass
                        }
                    }
                }
                cur.m = em; cur.cx = ex; cur.cy = ey;
            } else {                               // summed by heavy_cells_kernel
                cur.m = t.mass[offF + code]; cur.cx = t.comx[offF + code]; cur.cy = t.comy[offF + code];
            }
        }
        store_cell(t, offF + code, cur, F, F, G, mass_eps, scale, consts->thr2[F], sidx);
    }
    if (F == 0) {
        if (blockIdx.x == 0 && tid == 0) counters[4] = 1ull;   // the root alone
        return;
    }
    int level = F;
    // ---- levels F-1, F-2: shuffles inside the warp (groups of 4, then 16 lanes)
#pragma unroll
    for (int stride = 1; stride <= 4; stride <<= 2) {
        if (level == 0) break;
        Cell ch[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) ch[q] = shfl_cell(cur, (lane & ~(4 * stride - 1)) + q * stride);
        const uint64_t child_base = d.level_off[level] + 4 * (code >> 2);
        --level; code >>= 2; ncells >>= 2;
        const bool owner = (lane & (4 * stride - 1)) == 0 && code < ncells;
        Cell up = combine4<SHARDED>(ch, sidx, pos, mass, t.self_node, child_base, owner);
        if (owner) {
            store_cell(t, d.level_off[level] + code, up, level, F, G, mass_eps, scale, consts->thr2[level], sidx);
            n_internal += (up.cnt >= 2u);
        }
        cur = up;
    }
    // ---- levels F-3, F-4: shared memory
    if (level > 0) {   // block-uniform
        if ((lane & 15) == 0) s_a[tid >> 4] = cur;               // 16 cells of level F-2
        __syncthreads();
        const Cell* src = s_a;
        Cell* dst = s_b;
        int nthreads = kBottomThreads / 64;                      // 4 parents at level F-3, then 1 at F-4
#pragma unroll
        for (int stage = 0; stage < 2; ++stage) {
            if (level == 0) break;                               // block-uniform
            const uint64_t pcode = ((uint64_t)blockIdx.x * nthreads) + tid;   // parent code at level-1
            const uint64_t npar = 1ull << (2 * (level - 1));
            if (tid < nthreads) {
                Cell ch[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) ch[q] = src[tid * 4 + q];
                const bool ok = pcode < npar;
                Cell up = combine4<SHARDED>(ch, sidx, pos, mass, t.self_node, d.level_off[level] + 4 * pcode, ok);
                if (ok) {
                    store_cell(t, d.level_off[level - 1] + pcode, up, level - 1, F, G, mass_eps, scale,
                               consts->thr2[level - 1], sidx);
                    n_internal += (up.cnt >= 2u);
                }
                dst[tid] = up;
            }
            __syncthreads();
            --level;
            src = dst; dst = s_c; nthreads >>= 2;
        }
    }
    // ---- count internal cells (reference node count = 1 + 4 * internal)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_internal += __shfl_xor_sync(0xffffffffu, n_internal, o);
    if (lane == 0) s_internal[tid >> 5] = n_internal;
    __syncthreads();
    if (tid == 0) {
        uint32_t tot = 0;
        for (int w = 0; w < kBottomThreads / 32; ++w) tot += s_internal[w];
        if (tot) atomicAdd(&counters[5], (unsigned long long)tot);
        __threadfence();                                   // this block's cells and count are visible before the ticket
        s_last = atomicAdd(done_ticket, 1u) == gridDim.x - 1u;
    }
    __syncthreads();
    if (s_last) {                                          // block-uniform: the last block finishes the pyramid
        __threadfence();
        if (tid == 0) *done_ticket = 0u;
        int top_level = F - 5;                             // this kernel covered F .. F-4
        if (top_level < 0) top_level = -1;                 // the bottom levels already reached the root: only the node count remains
        tree_top_levels<SHARDED>(t, d, top_level, sidx, pos, mass, G, mass_eps, scale, counters, consts, &s_internal[0]);
    }
}

}  // namespace

// Phase 1: finest-cell runs (+ heavy cells, + this rank's partial sums when sharded).
void launch_tree_runs(const uint32_t* skeys, const uint32_t* sidx, const double2* pos, const double* mass,
                      int64_t n, const bh_params& p, const Dims& d, TreeArrays& t, Scratch& s, double* sums,
                      cudaStream_t st, const PeerComm* pc) {
    const int F = d.finest;
    uint32_t* cnt_f = t.count + d.level_off[F];
    uint32_t* first_f = t.first + d.level_off[F];
    const uint64_t nc = d.ncells_finest;
    uint32_t exact_max = (uint32_t)(p.exact_leaf_max < 0 ? 0 : p.exact_leaf_max);
    (void)skeys;
    launch_chain(cell_scan_kernel, dim3((unsigned)((nc + 256 * kScanItems - 1) / (256 * kScanItems))), dim3(256), st, true,
                 (const uint32_t*)cnt_f, first_f, nc, exact_max, s.heavy_list, s.heavy_count, s.huge_list, s.huge_count,
                 s.scan_state, s.scan_ticket);
    ++g_launches;
    const unsigned hh_blocks = kHeavyBlocks + kHugeParts * kHugeSlots;
    if (sums) {
        heavy_huge_kernel<<<hh_blocks, 256, 0, st>>>(s.heavy_list, s.heavy_count, s.huge_list, s.huge_count, cnt_f, first_f,
                                                     sidx, pos, mass, sums + nc, sums + 2 * nc, sums + 3 * nc, true,
                                                     s.huge_partial, s.huge_tickets);
        ++g_launches;
        PeerComm none{};
        none.n_ranks = 1;
        if (pc) cell_partial_kernel<true><<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(cnt_f, first_f, nc, sidx, pos, mass,
                                                                                       exact_max, sums, *pc);
        else cell_partial_kernel<false><<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(cnt_f, first_f, nc, sidx, pos, mass,
                                                                                     exact_max, sums, none);
        ++g_launches;
    } else {
        launch_chain(heavy_huge_kernel, dim3(hh_blocks), dim3(256), st, true, (const uint32_t*)s.heavy_list,
                     (const uint32_t*)s.heavy_count, (const uint32_t*)s.huge_list, (const uint32_t*)s.huge_count,
                     (const uint32_t*)cnt_f, (const uint32_t*)first_f, sidx, pos, mass, t.mass + d.level_off[F],
                     t.comx + d.level_off[F], t.comy + d.level_off[F], false, s.huge_partial, s.huge_tickets);
        ++g_launches;
    }
}

// Phase 2: all levels bottom-up.  `sums` (sharded build) = all-reduced per-cell sums.
void launch_tree_levels(const uint32_t* sidx, const double2* pos, const double* mass, const bh_params& p,
                        const Dims& d, TreeArrays& t, Scratch& s, const StepConsts* consts, const double* sums,
                        cudaStream_t st, const PeerComm* pc) {
    PeerComm none{};
    none.n_ranks = 1;
    const int F = d.finest;
    uint32_t exact_max = (uint32_t)(p.exact_leaf_max < 0 ? 0 : p.exact_leaf_max);
    unsigned blocks = (unsigned)((d.ncells_finest + kBottomThreads - 1) / kBottomThreads);
    // bbox_ticket is free again here (the bounds kernel resets it) — reused as the "blocks done" ticket
    if (sums) tree_bottom_kernel<true><<<blocks, kBottomThreads, 0, st>>>(t, d, sidx, pos, mass, p.G, p.mass_eps, exact_max,
                                                                        s.counters, consts, sums, s.bbox_ticket, pc ? *pc : none);
    else launch_chain(tree_bottom_kernel<false>, dim3(blocks), dim3(kBottomThreads), st, true, t, d, sidx, pos, mass, p.G,
                      p.mass_eps, exact_max, s.counters, consts, (const double*)nullptr, s.bbox_ticket, none);
    ++g_launches;
}

void launch_tree(const uint32_t* skeys, const uint32_t* sidx, const double2* pos, const double* mass,
                 int64_t n, const bh_params& p, const Dims& d, TreeArrays& t, Scratch& s,
                 const StepConsts* consts, cudaStream_t st) {
    launch_tree_runs(skeys, sidx, pos, mass, n, p, d, t, s, nullptr, st);
    launch_tree_levels(sidx, pos, mass, p, d, t, s, consts, nullptr, st);
}

}  // namespace bh
