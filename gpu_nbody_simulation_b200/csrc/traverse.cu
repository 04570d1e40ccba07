// Kernel 4 (+5): warp-coherent theta-criterion traversal with the integrator fused as epilogue.
//
// Replaces computeForcesGpu (project.cu:679-793; CPU twin computeForces :593-675) and
// updateAccVelPos (project.cu:819-836).
//
// One warp owns 32 consecutive bodies of the Morton-sorted order (lane = body) and walks the
// dense pyramid depth first with ONE warp-shared stack in shared memory.  A stack entry is
// (parent cell, mask of the lanes that opened it).  Popping a parent evaluates its four children
// (one 128-byte line of NodeRec, warp-uniform 128-bit loads) for all participating lanes; every
// lane applies the reference's test on its own:
//
//     skip   if mass <= mass_eps                                        project.cu:731
//     accept if leaf || size / (sqrt(d2) + eps) < theta                 project.cu:757
//     self   if leaf && sole occupant == this body -> no force          project.cu:760
//     open   otherwise -> the lane's bit goes into the child's mask     project.cu:776-785
//
// __ballot_sync over "open" is the child's mask; a child nobody opens is never pushed.  Because a
// lane takes part in a parent's children only if its bit is in the parent's mask, each body sees
// exactly the node set the reference's per-body DFS visits (per-lane acceptance semantics, SURVEY
// H2) while the warp fetches every node once.  Only the floating-point summation order differs.
//
// Arithmetic.  FP32 mode (default): the displacement COM - x is formed from double-float pairs
// (hi + lo), exact to ~2^-48 of the coordinate, everything after it is FP32 (forces are
// dominated by self-inclusive cap-leaf interactions at distances ~1e-8 of coordinates ~0.1).
// Coordinates are pre-multiplied by a power of two (StepConsts::scale) so that d2 (d + eps) stays
// inside the FP32 exponent range; the factor cancels in the force.  size/(d+eps) < theta is
// evaluated as d2 > (size/theta - eps)^2 with the per-level constant precomputed in FP64.
// Limit of this mode: separations below ~2^-48 of the coordinate (exactly coincident bodies whose
// COM differs from them by one FP64 ulp) cannot be resolved; use the FP64 mode for those.  FP64 mode (BH_FLAG_FP64_TRAVERSAL): the reference's expressions verbatim.
#include "bh_internal.h"

namespace bh {

namespace {

constexpr int kTravThreads = 256;
constexpr int kTravWarps = kTravThreads / 32;
constexpr int kStackCap = 3 * kMaxDepthDense + 8;

struct TravArgs {
    const uint32_t* skeys;      // sorted cell keys
    const uint32_t* sidx;       // body index per sorted position
    const uint32_t* own_list;   // optional: sorted positions owned by this rank (multi-GPU)
    double2* pos;
    double2* vel;
    double2* acc;
    double2* force;
    const double* mass;
    const NodeRec* rec;
    const double* t_mass;
    const double* t_comx;
    const double* t_comy;
    const StepConsts* consts;
    unsigned long long* counters;
    int64_t n_slots;            // bodies this launch evaluates
    double G, dt, theta, dist_eps;
    int finest;
    uint32_t level_off[kMaxLevels];
};

__device__ __forceinline__ float approx_sqrt(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float approx_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool FP64, bool INTEGRATE, bool COUNT>
__global__ void __launch_bounds__(kTravThreads)
traverse_kernel(const __grid_constant__ TravArgs a) {
    __shared__ uint2 s_stack[kTravWarps][kStackCap];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t slot = ((int64_t)blockIdx.x * kTravWarps + warp) * 32 + lane;
    const bool live = slot < a.n_slots;
    const int F = a.finest;

    uint32_t body = 0, key = 0;
    double px = 0.0, py = 0.0, mi = 0.0;
    if (live) {
        uint32_t sp = a.own_list ? a.own_list[slot] : (uint32_t)slot;
        body = a.sidx[sp];
        key = a.skeys[sp];
        double2 p = a.pos[body];
        px = p.x; py = p.y;
        mi = a.mass[body];
    }
    // double-float split of the (power-of-two scaled) body position (FP32 mode)
    const double scale = a.consts->scale;
    const double spx = px * scale, spy = py * scale;
    const float xh = (float)spx, yh = (float)spy;
    const float xl = (float)(spx - (double)xh), yl = (float)(spy - (double)yh);
    const float feps = a.consts->feps;

    float ax = 0.f, ay = 0.f;        // FP32 mode: sum of G M d / (d2 (d+eps)), times m_i at the end
    double sx = 0.0, sy = 0.0;       // FP64 mode: the reference's `sum`
    uint32_t c_int = 0, c_vis = 0, c_open = 0, c_steps = 0;

    // Evaluate one node for this lane.  Returns true if the lane opens it.
    auto eval = [&](uint32_t node_index, uint32_t level, uint32_t code, bool active, float thr2, double size) -> bool {
        const NodeRec r = a.rec[node_index];
        const bool nz = r.flags & kNodeNonZero, leaf = r.flags & kNodeLeaf;
        const bool self = leaf && (r.flags & kNodeSingle) && ((key >> (2 * (F - (int)level))) == code);
        bool accept;
        if constexpr (FP64) {
            const double M = a.t_mass[node_index];
            const double dx = a.t_comx[node_index] - px, dy = a.t_comy[node_index] - py;
            const double d2 = dx * dx + dy * dy;
            const double d = sqrt(d2) + a.dist_eps;                       // project.cu:748
            accept = leaf || (size / d < a.theta);                        // project.cu:757
            if (active && nz && accept && !self) {
                const double fm = (a.G * mi * M) / d2;                    // project.cu:765
                sx += fm * (dx / d);                                      // project.cu:768-772
                sy += fm * (dy / d);
            }
        } else {
            const float dx = (r.chx - xh) + (r.clx - xl);
            const float dy = (r.chy - yh) + (r.cly - yl);
            const float d2 = fmaf(dx, dx, dy * dy);
            accept = leaf || (d2 > thr2);
            // G M / (d2 (d + eps)), project.cu:765-769; d2 == 0 -> inf, times dx == 0 -> NaN like the reference
            const float w = d2 * (approx_sqrt(d2) + feps);
            float f = r.gm * approx_rcp(w);
            f = (active && nz && accept && !self) ? f : 0.f;
            ax = fmaf(f, dx, ax);
            ay = fmaf(f, dy, ay);
        }
        if constexpr (COUNT) {
            c_vis += active;
            c_int += (active && nz && accept && !self);
            c_open += (active && nz && !accept);
        }
        return active && nz && !accept;
    };

    int top = 0;
    {   // the root (project.cu:711-715 pushes node 0)
        bool open = eval(0u, 0u, 0u, live, a.consts->thr2[0], FP64 ? a.consts->size[0] : 0.0);
        uint32_t m = __ballot_sync(0xffffffffu, open);
        if (m) {
            if (lane == 0) s_stack[warp][0] = make_uint2(0u, m);
            top = 1;
        }
        __syncwarp();
    }
    while (top > 0) {
        const uint2 e = s_stack[warp][--top];
        __syncwarp();
        const uint32_t plevel = e.x >> 28, pcode = e.x & 0x0fffffffu;
        const bool active = (e.y >> lane) & 1u;
        const uint32_t level = plevel + 1u;
        const uint32_t base = a.level_off[level] + 4u * pcode;
        const float thr2 = a.consts->thr2[level];
        const double size = FP64 ? a.consts->size[level] : 0.0;
        if constexpr (COUNT) c_steps += (lane == 0);
#pragma unroll
        for (uint32_t q = 0; q < 4; ++q) {
            const uint32_t code = 4u * pcode + q;
            bool open = eval(base + q, level, code, active, thr2, size);
            uint32_t m = __ballot_sync(0xffffffffu, open);
            if (m) {
                if (lane == 0) s_stack[warp][top] = make_uint2((level << 28) | code, m);
                ++top;
            }
        }
        __syncwarp();
    }

    // ---- epilogue: force, and optionally a = F/m, v += a dt, x += v dt (project.cu:827-834) ----
    if (live) {
        double fx, fy;
        if constexpr (FP64) { fx = sx; fy = sy; }
        else { fx = mi * (double)ax; fy = mi * (double)ay; }
        a.force[body] = make_double2(fx, fy);
        if constexpr (INTEGRATE) {
            const double accx = __ddiv_rn(fx, mi), accy = __ddiv_rn(fy, mi);
            double2 v = a.vel[body];
            v.x = __dadd_rn(v.x, __dmul_rn(accx, a.dt));
            v.y = __dadd_rn(v.y, __dmul_rn(accy, a.dt));
            const double nx = __dadd_rn(px, __dmul_rn(v.x, a.dt));
            const double ny = __dadd_rn(py, __dmul_rn(v.y, a.dt));
            a.acc[body] = make_double2(accx, accy);
            a.vel[body] = v;
            a.pos[body] = make_double2(nx, ny);
        }
    }
    if constexpr (COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            c_int += __shfl_xor_sync(0xffffffffu, c_int, o);
            c_vis += __shfl_xor_sync(0xffffffffu, c_vis, o);
            c_open += __shfl_xor_sync(0xffffffffu, c_open, o);
        }
        if (lane == 0) {
            atomicAdd(&a.counters[0], (unsigned long long)c_int);
            atomicAdd(&a.counters[1], (unsigned long long)c_vis);
            atomicAdd(&a.counters[2], (unsigned long long)c_open);
            atomicAdd(&a.counters[3], (unsigned long long)c_steps);
        }
    }
}

// standalone integrator (used by the phase-split API: bh_integrate)
__global__ void __launch_bounds__(256)
integrate_kernel(double2* __restrict__ pos, double2* __restrict__ vel, double2* __restrict__ acc,
                 const double2* __restrict__ force, const double* __restrict__ mass, int64_t lo, int64_t hi,
                 double dt) {
    int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const double m = mass[i];
    const double2 f = force[i];
    double2 v = vel[i], x = pos[i];
    const double accx = __ddiv_rn(f.x, m), accy = __ddiv_rn(f.y, m);     // project.cu:827-828
    v.x = __dadd_rn(v.x, __dmul_rn(accx, dt));                            // project.cu:830-831
    v.y = __dadd_rn(v.y, __dmul_rn(accy, dt));
    x.x = __dadd_rn(x.x, __dmul_rn(v.x, dt));                             // project.cu:833-834
    x.y = __dadd_rn(x.y, __dmul_rn(v.y, dt));
    acc[i] = make_double2(accx, accy);
    vel[i] = v;
    pos[i] = x;
}

// Multi-GPU: sorted positions whose body index lies in [lo, hi), order preserved inside a block.
__global__ void __launch_bounds__(256)
own_list_kernel(const uint32_t* __restrict__ sidx, int64_t n, uint32_t lo, uint32_t hi,
                uint32_t* __restrict__ own_list, uint32_t* __restrict__ own_count) {
    __shared__ uint32_t s_w[8];
    __shared__ uint32_t s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool mine = false;
    if (j < n) { uint32_t b = sidx[j]; mine = b >= lo && b < hi; }
    uint32_t m = __ballot_sync(0xffffffffu, mine);
    if (lane == 0) s_w[warp] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
        for (int w = 0; w < 8; ++w) { uint32_t c = s_w[w]; s_w[w] = tot; tot += c; }
        s_base = tot ? atomicAdd(own_count, tot) : 0u;
    }
    __syncthreads();
    if (mine) own_list[s_base + s_w[warp] + __popc(m & ((1u << lane) - 1u))] = (uint32_t)j;
}

}  // namespace

void launch_traverse(const uint32_t* skeys, const uint32_t* sidx, double2* pos, double2* vel, double2* acc,
                     double2* force, const double* mass, int64_t n, int64_t own_lo, int64_t own_hi,
                     const uint32_t* own_list, const uint32_t* own_count_dev, int64_t own_n,
                     const bh_params& p, const Dims& d, const TreeArrays& t, const StepConsts* consts,
                     unsigned long long* counters, bool integrate, cudaStream_t st) {
    (void)own_lo; (void)own_hi; (void)own_count_dev; (void)n;
    TravArgs a;
    a.skeys = skeys; a.sidx = sidx; a.own_list = own_list;
    a.pos = pos; a.vel = vel; a.acc = acc; a.force = force; a.mass = mass;
    a.rec = t.rec; a.t_mass = t.mass; a.t_comx = t.comx; a.t_comy = t.comy;
    a.consts = consts; a.counters = counters;
    a.n_slots = own_n;
    a.G = p.G; a.dt = p.dt; a.theta = p.theta; a.dist_eps = p.dist_eps;
    a.finest = d.finest;
    for (int l = 0; l < kMaxLevels; ++l) a.level_off[l] = (uint32_t)d.level_off[l < d.max_depth ? l : d.max_depth];
    if (own_n <= 0) return;
    unsigned blocks = (unsigned)((own_n + kTravThreads - 1) / kTravThreads);
    const bool fp64 = p.flags & BH_FLAG_FP64_TRAVERSAL, count = p.flags & BH_FLAG_COUNTERS;
#define BH_TRAV(F, I, C) traverse_kernel<F, I, C><<<blocks, kTravThreads, 0, st>>>(a)
    if (fp64) {
        if (integrate) { if (count) BH_TRAV(true, true, true); else BH_TRAV(true, true, false); }
        else { if (count) BH_TRAV(true, false, true); else BH_TRAV(true, false, false); }
    } else {
        if (integrate) { if (count) BH_TRAV(false, true, true); else BH_TRAV(false, true, false); }
        else { if (count) BH_TRAV(false, false, true); else BH_TRAV(false, false, false); }
    }
#undef BH_TRAV
    ++g_launches;
}

void launch_integrate(double2* pos, double2* vel, double2* acc, const double2* force, const double* mass,
                      int64_t lo, int64_t hi, double dt, cudaStream_t st) {
    if (hi <= lo) return;
    integrate_kernel<<<(unsigned)((hi - lo + 255) / 256), 256, 0, st>>>(pos, vel, acc, force, mass, lo, hi, dt);
    ++g_launches;
}

void launch_own_list(const uint32_t* sidx, int64_t n, int64_t lo, int64_t hi, uint32_t* own_list,
                     uint32_t* own_count, cudaStream_t st) {
    own_list_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sidx, n, (uint32_t)lo, (uint32_t)hi, own_list,
                                                                 own_count);
    ++g_launches;
}

}  // namespace bh
