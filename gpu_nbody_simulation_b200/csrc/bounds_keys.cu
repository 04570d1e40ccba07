// Kernel 1: FP64 bounding box + per-level constants, and per-body cell keys.
//
// Replaces ComputeRootBounds (project.cu:536-573) and the root-to-cap descent that QuadInsert /
// DetermineChild perform per body (project.cu:348-356, :358-453).  Both are bit-exact by
// construction: min/max are order independent, the padding uses the reference's operation
// order with contraction disabled (__dmul_rn / __dadd_rn), and the key is produced by the same
// FP64 bisection `mid = (min + max) / 2` with the same four-way comparison the reference uses.
#include "bh_internal.h"
#include "peer_comm.cuh"

namespace bh {

namespace {

constexpr int kBoundsThreads = 256;

// std::min(a, b) == (b < a) ? b : a ; std::max(a, b) == (a < b) ? b : a   (project.cu:547-550)
__device__ __forceinline__ double ref_min(double cur, double x) { return (x < cur) ? x : cur; }
__device__ __forceinline__ double ref_max(double cur, double x) { return (cur < x) ? x : cur; }

// Padding (project.cu:553-570), per-level cell sizes and acceptance thresholds, FP32 scale.
__device__ void finalize_bounds(StepConsts* __restrict__ consts, double xmin, double xmax, double ymin, double ymax,
                                double pad_frac, double pad_fallback, double theta, double dist_eps, int finest) {
    // project.cu:553-570, same operation order, no FMA contraction
    double dx = __dsub_rn(xmax, xmin), dy = __dsub_rn(ymax, ymin);
    double max_dim = (dx < dy) ? dy : dx;                 // std::max(dx, dy)
    double pad = __dmul_rn(pad_frac, max_dim);
    if (max_dim == 0.0) pad = pad_fallback;
    xmin = __dsub_rn(xmin, pad); xmax = __dadd_rn(xmax, pad);
    ymin = __dsub_rn(ymin, pad); ymax = __dadd_rn(ymax, pad);
    consts->xmin = xmin; consts->xmax = xmax; consts->ymin = ymin; consts->ymax = ymax;
    // power-of-two normalisation of the FP32 traversal coordinates: box extent -> [2^20, 2^21)
    double ext = fmax(__dsub_rn(xmax, xmin), __dsub_rn(ymax, ymin));
    int e2 = (ext > 0.0 && ext < INFINITY) ? ilogb(ext) : 20;
    int se = 20 - e2;
    se = se > 120 ? 120 : (se < -120 ? -120 : se);
    const double scale = scalbn(1.0, se);
    consts->scale = scale;
    consts->feps = (float)(dist_eps * scale);
    // cell extent per level, following the low-side bisection chain (project.cu:417-428)
    double xl = xmin, xh = xmax, yl = ymin, yh = ymax;
    for (int l = 0; l < kMaxLevels; ++l) {
        double w = __dsub_rn(xh, xl), h = __dsub_rn(yh, yl);
        double size = (w > h) ? w : h;                    // project.cu:637-639
        consts->size[l] = size;
        // size / (d + eps) < theta  <=>  d > size/theta - eps
        double thr = (theta > 0.0) ? (size / theta - dist_eps) : INFINITY;
        consts->thr[l] = thr;
        float t2;
        if (!(theta > 0.0)) t2 = INFINITY;
        else if (thr < 0.0) t2 = -1.0f;
        else t2 = (float)((thr * scale) * (thr * scale));
        consts->thr2[l] = t2;
        if (l < finest) {
            xh = __dmul_rn(__dadd_rn(xl, xh), 0.5);
            yh = __dmul_rn(__dadd_rn(yl, yh), 0.5);
        }
    }
}

// Cell boundaries of the finest level along x and y: bnd[i] = lower edge of column / row i (i = 0 .. 2^F),
// produced by the SAME chain of FP64 bisections the reference descends (project.cu:417-428: the child
// takes [min, mid] or [mid, max] with mid = (min + max) / 2), following the bits of i from the root down.
// The finest cell of a coordinate x is then the unique i with bnd[i] <= x < bnd[i + 1] (keys_kernel).
// Called by every thread of the block that finalised the box (consts already written + __syncthreads).
__device__ void fill_cell_bounds(const StepConsts* consts, int finest, double* __restrict__ bnd) {
    if (!bnd) return;
    const uint32_t nc = 1u << finest;
    for (uint32_t t = threadIdx.x; t < 2u * (nc + 1u); t += blockDim.x) {
        const bool is_y = t > nc;
        const uint32_t i = is_y ? t - (nc + 1u) : t;
        double lo = is_y ? consts->ymin : consts->xmin, hi = is_y ? consts->ymax : consts->xmax;
        if (i == nc) {
            lo = hi;
        } else {
            for (int l = finest - 1; l >= 0; --l) {
                const double mid = __dmul_rn(__dadd_rn(lo, hi), 0.5);
                if ((i >> l) & 1u) lo = mid; else hi = mid;
            }
        }
        bnd[t] = lo;
    }
}

__global__ void __launch_bounds__(kBoundsThreads)
bounds_kernel(const double2* __restrict__ pos, int64_t n, double* __restrict__ partial,
              uint32_t* __restrict__ ticket, StepConsts* __restrict__ consts, double pad_frac,
              double pad_fallback, double theta, double dist_eps, int finest, double* __restrict__ raw_out,
              double* __restrict__ cell_bnd, const __grid_constant__ PeerComm pc) {
    pdl_entry();
    double xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double2 p = pos[i];
        xmin = ref_min(xmin, p.x); xmax = ref_max(xmax, p.x);
        ymin = ref_min(ymin, p.y); ymax = ref_max(ymax, p.y);
    }
    __shared__ double s[4][kBoundsThreads / 32];
    __shared__ bool last;
    auto block_reduce = [&]() {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            xmin = ref_min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
            xmax = ref_max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
            ymin = ref_min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
            ymax = ref_max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
        }
        int w = threadIdx.x >> 5, l = threadIdx.x & 31;
        if (l == 0) { s[0][w] = xmin; s[1][w] = xmax; s[2][w] = ymin; s[3][w] = ymax; }
        __syncthreads();
        if (w == 0) {
            constexpr int NW = kBoundsThreads / 32;
            xmin = l < NW ? s[0][l] : INFINITY; xmax = l < NW ? s[1][l] : -INFINITY;
            ymin = l < NW ? s[2][l] : INFINITY; ymax = l < NW ? s[3][l] : -INFINITY;
#pragma unroll
            for (int o = NW / 2; o > 0; o >>= 1) {
                xmin = ref_min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
                xmax = ref_max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
                ymin = ref_min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
                ymax = ref_max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
            }
        }
        __syncthreads();
    };
    block_reduce();
    if (threadIdx.x == 0) {
        double* o = partial + 4 * (size_t)blockIdx.x;
        o[0] = xmin; o[1] = xmax; o[2] = ymin; o[3] = ymax;
        __threadfence();
        last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    xmin = INFINITY; xmax = -INFINITY; ymin = INFINITY; ymax = -INFINITY;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
        const volatile double* o = partial + 4 * (size_t)b;
        xmin = ref_min(xmin, o[0]); xmax = ref_max(xmax, o[1]);
        ymin = ref_min(ymin, o[2]); ymax = ref_max(ymax, o[3]);
    }
    block_reduce();
    if (pc.n_ranks > 1) {
        // fused exchange over NVLink peer memory: this block publishes the rank's box to every peer, waits
        // for theirs and finalises; the step's sequence number lives in the rank's own comm buffer
        __shared__ uint32_t s_seq;
        if (threadIdx.x == 0) {
            uint32_t* seq_dev = reinterpret_cast<uint32_t*>(pc.peer_base[pc.rank] + pc.off_err) + 1;
            s_seq = *seq_dev + 1u;
            *seq_dev = s_seq;
        }
        __syncthreads();
        peer_bbox_exchange(pc, s_seq, xmin, xmax, ymin, ymax);
    }
    if (threadIdx.x == 0) {
        *ticket = 0;
        if (raw_out) {   // NCCL fallback: min / max are all-reduced over the ranks first (max sent negated)
            raw_out[0] = xmin; raw_out[1] = -xmax; raw_out[2] = ymin; raw_out[3] = -ymax;
        } else {
            finalize_bounds(consts, xmin, xmax, ymin, ymax, pad_frac, pad_fallback, theta, dist_eps, finest);
        }
    }
    if (!raw_out) {
        __syncthreads();   // consts written by thread 0 are visible to the block
        fill_cell_bounds(consts, finest, cell_bnd);
    }
}

__global__ void bounds_finalize_kernel(const double* __restrict__ raw, StepConsts* __restrict__ consts, double pad_frac,
                                       double pad_fallback, double theta, double dist_eps, int finest,
                                       double* __restrict__ cell_bnd) {
    if (threadIdx.x == 0)
        finalize_bounds(consts, raw[0], -raw[1], raw[2], -raw[3], pad_frac, pad_fallback, theta, dist_eps, finest);
    __syncthreads();
    fill_cell_bounds(consts, finest, cell_bnd);
}

// Spread the low 16 bits of v to the even bit positions.
__device__ __forceinline__ uint32_t spread_bits(uint32_t v) {
    v &= 0xffffu;
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

// Column (or row) of coordinate x: the unique i with bnd[i] <= x < bnd[i + 1].  The multiply only
// proposes a candidate; the comparisons against the bisection-chain boundaries decide, so the result
// is the reference's descent bit for bit (x >= mid goes up at every level).  Typically 0 corrections.
__device__ __forceinline__ uint32_t locate_cell(double x, double x0, double inv_w, const double* __restrict__ bnd,
                                                uint32_t nc) {
    double g = (x - x0) * inv_w;
    int i = (g >= 0.0) ? ((g < (double)nc) ? (int)g : (int)nc - 1) : 0;   // NaN -> 0
    while (i > 0 && x < __ldg(bnd + i)) --i;
    while (i < (int)nc - 1 && x >= __ldg(bnd + i + 1)) ++i;
    return (uint32_t)i;
}

// One thread per body.  Also accumulates the radix digit histograms of all passes.
// TABLE = false (default): the descent itself.  TABLE = true (BH_KEYS_TABLE=1): finest cell by lookup in
// the boundary table written by the bounds kernel (2 multiplies + ~4 compares per body instead of 2 x F
// dependent FP64 bisections; measured no faster on B200).  Both are bit-exact restatements of
// project.cu:348-356.
template <bool TABLE>
__global__ void __launch_bounds__(256)
keys_kernel(const double2* __restrict__ pos, int64_t n, int finest, const StepConsts* __restrict__ consts,
            uint32_t* __restrict__ keys, uint32_t* __restrict__ cell_count, uint32_t* __restrict__ digit_hist,
            int passes, int bits_per_pass, uint32_t idx_base, const double* __restrict__ cell_bnd) {
    __shared__ uint32_t hist[kMaxSortPasses * kMaxBins];
    pdl_entry();
    for (int i = threadIdx.x; i < passes * kMaxBins; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const double bx0 = consts->xmin, bx1 = consts->xmax, by0 = consts->ymin, by1 = consts->ymax;
    const uint32_t dmask = (1u << bits_per_pass) - 1u;
    const uint32_t nc = 1u << finest;
    const double inv_wx = (double)nc / (bx1 - bx0), inv_wy = (double)nc / (by1 - by0);
    const double* __restrict__ bnd_x = cell_bnd;
    const double* __restrict__ bnd_y = cell_bnd + (nc + 1u);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double2 p = pos[i];
        uint32_t key = 0;
        // project.cu:352-355 tests (x<mx && y<my) -> 0, (x>=mx && y<my) -> 1, (x<mx && y>=my) -> 2, else 3.
        // For ordered values that is (x >= mx) | (y >= my) << 1; a NaN coordinate fails all three
        // tests and goes to child 3 at every level.
        const bool unordered = (p.x != p.x) || (p.y != p.y);
        if constexpr (TABLE) {
            const uint32_t ix = unordered ? nc - 1u : locate_cell(p.x, bx0, inv_wx, bnd_x, nc);
            const uint32_t iy = unordered ? nc - 1u : locate_cell(p.y, by0, inv_wy, bnd_y, nc);
            key = spread_bits(ix) | (spread_bits(iy) << 1);       // level 0 decision = most significant pair
        } else {
            double xl = bx0, xh = bx1, yl = by0, yh = by1;
            for (int l = 0; l < finest; ++l) {
                const double mx = __dmul_rn(__dadd_rn(xl, xh), 0.5);   // (min + max) / 2, exact halving
                const double my = __dmul_rn(__dadd_rn(yl, yh), 0.5);
                const bool bx = unordered || (p.x >= mx), by = unordered || (p.y >= my);
                xl = bx ? mx : xl; xh = bx ? xh : mx;               // child bounds project.cu:421-429
                yl = by ? my : yl; yh = by ? yh : my;
                key = (key << 2) | (uint32_t)bx | ((uint32_t)by << 1);
            }
        }
        keys[i] = key;               // (the first sort pass generates the body indices idx_base + i itself)
        {   // bodies per finest cell (zeroed every step): the cell runs of the sorted order follow from these counts by
            // a scan (tree_build.cu: cell_scan_kernel), without a pass over the sorted keys.  One atomic per DISTINCT key
            // of the warp — for the cell count and for the digit histograms alike: bodies in resident order share their
            // cell with their neighbours, so a warp has ~5 distinct keys (and 32-way conflicts on the high digit otherwise).
            const uint32_t act = __activemask();
            const uint32_t peers = __match_any_sync(act, key);
            if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) {
                const uint32_t cnt = (uint32_t)__popc(peers);
                atomicAdd(cell_count + key, cnt);
                for (int ps = 0; ps < passes; ++ps)
                    atomicAdd(&hist[ps * kMaxBins + ((key >> (ps * bits_per_pass)) & dmask)], cnt);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * kMaxBins; i += blockDim.x) {
        uint32_t v = hist[i];
        if (v) atomicAdd(&digit_hist[i], v);
    }
}

}  // namespace

void launch_bounds(const double2* pos, int64_t n, const bh_params& p, const Dims& d, Scratch& s,
                   StepConsts* consts, int grid, cudaStream_t st, double* raw_out, const PeerComm* pc) {
    PeerComm none{};
    none.n_ranks = 1;
    bounds_kernel<<<grid, kBoundsThreads, 0, st>>>(pos, n, s.bbox_partial, s.bbox_ticket, consts, p.pad_frac,
                                                   p.pad_fallback, p.theta, p.dist_eps, d.finest, raw_out,
                                                   s.cell_bnd, pc ? *pc : none);
    ++g_launches;
}

void launch_bounds_finalize(const double* raw, const bh_params& p, const Dims& d, Scratch& s, StepConsts* consts,
                            cudaStream_t st) {
    bounds_finalize_kernel<<<1, kBoundsThreads, 0, st>>>(raw, consts, p.pad_frac, p.pad_fallback, p.theta, p.dist_eps,
                                                         d.finest, s.cell_bnd);
    ++g_launches;
}

void launch_keys(const double2* pos, int64_t n, const Dims& d, const SortPlan& sp, const StepConsts* consts,
                 uint32_t* keys, uint32_t* cell_count, uint32_t* digit_hist, cudaStream_t st, uint32_t idx_base,
                 const double* cell_bnd) {
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    if (cell_bnd)
        launch_chain(keys_kernel<true>, dim3((unsigned)blocks), dim3(256), st, true, pos, n, d.finest, consts, keys, cell_count,
                     digit_hist, sp.passes, sp.bits_per_pass, idx_base, cell_bnd);
    else
        launch_chain(keys_kernel<false>, dim3((unsigned)blocks), dim3(256), st, true, pos, n, d.finest, consts, keys, cell_count,
                     digit_hist, sp.passes, sp.bits_per_pass, idx_base, (const double*)nullptr);
    ++g_launches;
}

}  // namespace bh
