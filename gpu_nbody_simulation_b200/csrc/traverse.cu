// Kernel 4 (+5): warp-coherent theta-criterion traversal with the integrator fused as epilogue.
//
// Replaces computeForcesGpu (project.cu:679-793; CPU twin computeForces :593-675) and
// updateAccVelPos (project.cu:819-836).
//
// Kernels in this file: traverse_f32_list_kernel (production from 500k bodies on: group-classified walk, see its own
// header further down), traverse_f32_kernel (generic, 1 or 2 bodies per lane: small inputs, counters, exact leaves),
// traverse_f32_pair_kernel (round 1's production kernel: A/B variant and packed exact leaves), traverse_f64_kernel
// (verification mode).  What they share — written for the generic / pair kernels first:
//
// One warp owns 32*BPL consecutive bodies of the Morton-sorted order (each lane BPL of them) and
// walks the dense pyramid depth first with ONE warp-shared stack in shared memory.  A stack entry
// is (cell p, mask of the bodies that opened it).  Popping p evaluates its four children
// 4p+1 .. 4p+4 (one 128-byte line of NodeRec, warp-uniform vector loads) for all participating
// bodies; every body applies the reference's test on its own:
//
//     skip   if mass <= mass_eps                                        project.cu:731
//     accept if leaf || size / (sqrt(d2) + eps) < theta                 project.cu:757
//     self   if leaf && sole occupant == this body -> no force          project.cu:760
//     open   otherwise -> the body's bit goes into the child's mask     project.cu:776-785
//
// __ballot_sync over "open" is the child's mask; a child nobody opens is never pushed.  Because a
// body takes part in a cell's children only if its bit is in the cell's mask, each body sees
// exactly the node set the reference's per-body DFS visits (per-lane acceptance semantics, SURVEY
// H2) while the warp fetches every node once.  Only the floating-point summation order differs.
//
// FP32 mode (default).  The tree build folds the node tests into the record: `thr` is the level's
// (size/theta - eps)^2 for internal nodes and -1 for leaves and skipped nodes, so "accept" is the
// single compare !(d2 <= thr); skipped nodes carry gm = 0 and a far-away COM; the self test is one
// integer compare against the body's own leaf index (self_node, written by the build).  The
// displacement COM - x is formed from double-float pairs (hi + lo), exact to ~2^-48 of the
// coordinate (forces are dominated by self-inclusive cap-leaf interactions at distances
// ~1e-8 of coordinates ~0.1), with packed FP32x2 instructions (FADD2 / FFMA2, sm_100); everything
// after it is FP32.  Coordinates are pre-multiplied by a power of two (StepConsts::scale) so that
// d2 (d + eps) stays inside the FP32 exponent range; the factor cancels in the force.
// Limit of this mode: separations below ~2^-48 of the coordinate (exactly coincident bodies whose
// COM differs from them by one FP64 ulp) cannot be resolved; use the FP64 mode for those.
//
// FP64 mode (BH_FLAG_FP64_TRAVERSAL): the reference's expressions verbatim on the FP64 tree arrays.
#include <algorithm>

#include "bh_internal.h"

namespace bh {

namespace {

#ifndef BH_TRAV_THREADS
#define BH_TRAV_THREADS 128   // A/B on B200 at N=1M (pair kernel): 256 thr/48 regs 463 us, 256/67 412 us, 128/67 404 us
#endif
constexpr int kTravThreads = BH_TRAV_THREADS;
constexpr int kTravWarps = kTravThreads / 32;
constexpr int kStackCap = 3 * kMaxDepthDense + 8;

struct TravArgs {
    const uint32_t* skeys;      // finest-cell key per sorted position
    const uint32_t* sidx;       // body index per sorted position
    const uint32_t* own_list;   // optional: sorted positions owned by this rank (multi-GPU)
    const uint32_t* self_node;  // per body: its own single-occupant leaf (pyramid index) or 0xffffffff
    const double2* pos_in;      // positions / velocities the step starts from (== pos / vel, or the
    const double2* vel_in;      // snapshot when the step restarts from it: no restore copy needed)
    double2* pos;               // outputs of the fused integrator
    double2* vel;
    double2* acc;
    double2* force;
    const double* mass;
    const NodeRec* rec;
    const uint32_t* flags;
    const double* t_mass;
    const double* t_comx;
    const double* t_comy;
    const StepConsts* consts;
    unsigned long long* counters;
    int64_t n_slots;            // bodies this launch evaluates
    double G, dt, theta, dist_eps;
    // BH_FLAG_EXACT_LEAVES only (appended: the layout seen by the other kernels is unchanged)
    const uint32_t* t_count;    // bodies per pyramid cell
    const uint32_t* t_first;    // sorted position of a cell's first body
    uint32_t finest_off;        // pyramid index of the first cap-level cell
    // list kernel: [0] next 64-body tile (warps fetch their work one tile at a time), [1] warps that found the queue
    // empty — the last one resets both words, so the pair is zero again when the kernel ends
    uint32_t* tile_queue;
};

__device__ __forceinline__ float approx_sqrt(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float approx_rsqrt(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float approx_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// epilogue shared by all variants: force, and optionally a = F/m, v += a dt, x += v dt
template <bool INTEGRATE>
__device__ __forceinline__ void finish_body(const TravArgs& a, uint32_t body, double px, double py, double mi,
                                            double fx, double fy) {
    a.force[body] = make_double2(fx, fy);
    if constexpr (INTEGRATE) {
        const double accx = __ddiv_rn(fx, mi), accy = __ddiv_rn(fy, mi);      // project.cu:827-828
        double2 v = a.vel_in[body];
        v.x = __dadd_rn(v.x, __dmul_rn(accx, a.dt));                          // project.cu:830-831
        v.y = __dadd_rn(v.y, __dmul_rn(accy, a.dt));
        const double nx = __dadd_rn(px, __dmul_rn(v.x, a.dt));               // project.cu:833-834
        const double ny = __dadd_rn(py, __dmul_rn(v.y, a.dt));
        a.acc[body] = make_double2(accx, accy);
        a.vel[body] = v;
        a.pos[body] = make_double2(nx, ny);
    }
}

// ---- FP64 rescue of bodies the FP32 frame cannot resolve ---------------------------------------------------------
// The double-float displacement resolves separations down to 2^-48 of the frame (the scaled root box, or the warp's
// box in the list kernel).  The one systematic case below that is a body whose OWN multi-body cap-level cell has its
// centre of mass (almost) on the body: coincident bodies, whose COM differs from them by an FP64 ulp — and the
// reference applies that cell to the body itself, so the pair decides the body's whole force.  Such a
// body is detected in the prologue (one FP64 subtraction against its own cell), taken out of the warp's FP32 walk and
// evaluated by the reference's own per-body FP64 walk (project.cu:593-675) in the epilogue.  Costs nothing in the hot
// loops; never triggers on the benchmark workloads.
constexpr double kTinyRel2 = 5.6e-17;   // (2^-27)^2: below 2^-27 of the frame the double-float form loses > 1e-6 of d

__device__ __forceinline__ bool own_cell_unresolved(const TravArgs& a, uint32_t sorted_pos, double px, double py,
                                                    double scale, double frame2) {
    const uint32_t ci = a.finest_off + __ldg(a.skeys + sorted_pos);
    if (__ldg(a.t_count + ci) < 2u) return false;
    const double dx = (__ldg(a.t_comx + ci) - px) * scale, dy = (__ldg(a.t_comy + ci) - py) * scale;
    return dx * dx + dy * dy < kTinyRel2 * frame2;          // false for NaN
}

__device__ __noinline__ void fp64_body_walk(const TravArgs& a, uint32_t selfn, double px, double py, double& sx, double& sy) {
    uint32_t stack[3 * kMaxDepthDense + 8];
    int top = 0;
    stack[top++] = 0u;
    double ax = 0.0, ay = 0.0;
    while (top > 0) {
        const uint32_t idx = stack[--top];
        const uint32_t fl = a.flags[idx];
        if (!(fl & kNodeNonZero)) continue;                                   // project.cu:617
        const int level = (31 - __clz(3u * idx + 1u)) >> 1;
        const double dx = a.t_comx[idx] - px, dy = a.t_comy[idx] - py;
        const double d2 = dx * dx + dy * dy;
        const double d = sqrt(d2) + a.dist_eps;                               // project.cu:634
        if ((fl & kNodeLeaf) || (a.consts->size[level] / d < a.theta)) {      // project.cu:643
            if (selfn != idx) {                                               // project.cu:646
                const double fm = (a.G * a.t_mass[idx]) / d2;                 // (times m_i in finish_body)
                ax += fm * (dx / d);
                ay += fm * (dy / d);
            }
        } else {
            for (uint32_t q = 0; q < 4; ++q) stack[top++] = 4u * idx + 1u + q;
        }
    }
    sx = ax; sy = ay;
}

// ------------------------------------------------------------------------------------------------
// FP32 traversal, BPL bodies per lane
// ------------------------------------------------------------------------------------------------
#ifndef BH_PAIR_MIN_BLOCKS
#define BH_PAIR_MIN_BLOCKS 3   // per 256 threads: leaves the register allocator free (67 regs); 5 (48 regs) is 13 % slower
#endif
constexpr int kPairMinBlocks = BH_PAIR_MIN_BLOCKS;
constexpr float kFarLane = -1.152921504606847e18f;   // -2^60: where bodies outside a cell's mask "stand"

template <int BPL> struct StackEntry;
template <> struct StackEntry<1> {
    static constexpr uint32_t kBytes = 8;
    static __device__ __forceinline__ void store(uint32_t addr, uint32_t node, const uint32_t (&m)[1]) {
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(node), "r"(m[0]) : "memory");
    }
    static __device__ __forceinline__ void store_if(uint32_t doit, uint32_t addr, uint32_t node, const uint32_t (&m)[1]) {
        asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %0, 0;\n @p st.shared.v2.u32 [%1], {%2, %3};\n}" ::"r"(doit),
                     "r"(addr), "r"(node), "r"(m[0]) : "memory");
    }
    static __device__ __forceinline__ void load(uint32_t addr, uint32_t& node, uint32_t (&m)[1]) {
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(node), "=r"(m[0]) : "r"(addr) : "memory");
    }
};
template <> struct StackEntry<2> {
    static constexpr uint32_t kBytes = 16;
    static __device__ __forceinline__ void store(uint32_t addr, uint32_t node, const uint32_t (&m)[2]) {
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %3};" ::"r"(addr), "r"(node), "r"(m[0]), "r"(m[1]) : "memory");
    }
    static __device__ __forceinline__ void store_if(uint32_t doit, uint32_t addr, uint32_t node, const uint32_t (&m)[2]) {
        asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %0, 0;\n @p st.shared.v4.u32 [%1], {%2, %3, %4, %4};\n}" ::"r"(doit),
                     "r"(addr), "r"(node), "r"(m[0]), "r"(m[1]) : "memory");
    }
    static __device__ __forceinline__ void load(uint32_t addr, uint32_t& node, uint32_t (&m)[2]) {
        uint32_t pad;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(node), "=r"(m[0]), "=r"(m[1]), "=r"(pad) : "r"(addr) : "memory");
    }
};

#ifndef BH_GENERIC_MIN_BLOCKS
#define BH_GENERIC_MIN_BLOCKS 4   // 48 registers; 8 (32 regs) is 25 % slower at N = 40k
#endif
// EXACT (BH_FLAG_EXACT_LEAVES, an extension — SURVEY 8f row f1, specified by
// oracle/bh_oracle.c:bho_compute_forces_exact_leaves): a multi-body leaf at the depth cap is not applied
// as one monopole at its centre of mass (which contains the body itself when it lives there)
// but as the sum over the leaf's bodies j != i of the same pair expression.  The leaf test depends on the
// node only, so the member loop is warp-uniform; members are fetched through the sorted list.
template <int BPL, bool INTEGRATE, bool COUNT, bool EXACT = false>
__global__ void __launch_bounds__(kTravThreads, (BPL == 1 ? BH_GENERIC_MIN_BLOCKS : 4) * (256 / kTravThreads))
traverse_f32_kernel(const __grid_constant__ TravArgs a) {
    using SE = StackEntry<BPL>;
    __shared__ __align__(16) uint8_t s_stack[kTravWarps][kStackCap * SE::kBytes];
    pdl_entry();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warp_slot0 = ((int64_t)blockIdx.x * kTravWarps + warp) * (32 * BPL);
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(&s_stack[warp][0]);

    uint32_t body[BPL], selfn[BPL];
    bool rescue[BPL];            // evaluated by fp64_body_walk in the epilogue instead of the warp's FP32 walk
    float2 nh[BPL], nl[BPL];     // minus the scaled body position, hi and lo floats
    float2 acc2[BPL];            // sum of G M d / (d2 (d + eps)); times m_i at the end
    const float feps = a.consts->feps;
    [[maybe_unused]] double scale_d = 0.0, Gs = 0.0;   // EXACT: coordinate scale and G * scale^2
    if constexpr (EXACT) { scale_d = a.consts->scale; Gs = a.G * scale_d * scale_d; }
    {
        const double scale = a.consts->scale;
#pragma unroll
        for (int b = 0; b < BPL; ++b) {
            const int64_t slot = warp_slot0 + b * 32 + lane;
            body[b] = 0xffffffffu; selfn[b] = 0xffffffffu;
            rescue[b] = false;
            double px = 0.0, py = 0.0;
            uint32_t sp = 0;
            if (slot < a.n_slots) {
                sp = a.own_list ? a.own_list[slot] : (uint32_t)slot;
                body[b] = a.sidx[sp];
                selfn[b] = a.self_node[body[b]];
                double2 p = a.pos_in[body[b]];
                px = p.x; py = p.y;
            }
            const double sx = px * scale, sy = py * scale;
            if constexpr (!EXACT) {
                if (slot < a.n_slots) rescue[b] = own_cell_unresolved(a, sp, px, py, scale, sx * sx + sy * sy);
            }
            const float xh = (float)sx, yh = (float)sy;
            nh[b] = make_float2(-xh, -yh);
            nl[b] = make_float2(-(float)(sx - (double)xh), -(float)(sy - (double)yh));
            acc2[b] = make_float2(0.f, 0.f);
        }
    }
    uint32_t c_int = 0, c_vis = 0, c_open = 0, c_steps = 0;

    // Evaluate one node for body slot b of this lane; `mh` is minus the body's scaled position (hi
    // floats) or the far-away stand-in when the body is not in the parent's mask: such a body accepts
    // everything (never opens) and receives exactly 0 (w overflows to inf, 1/inf == 0).
    // Returns the warp's ballot of the bodies (slot b of every lane) that open the node.
    auto eval = [&](const float4 A, const float2 B, uint32_t idx, int b, const float2 mh, bool active) -> uint32_t {
        const float2 dh = __fadd2_rn(make_float2(A.x, A.y), mh);
        const float2 dl = __fadd2_rn(make_float2(A.z, A.w), nl[b]);
        const float2 d = __fadd2_rn(dh, dl);
        const float d2 = fmaf(d.x, d.x, d.y * d.y);
        // G M / (d2 (d + eps)), project.cu:765-769; d2 == 0 -> inf, times dx == 0 -> NaN like the reference
        const float w = d2 * (approx_sqrt(d2) + feps);
        const float gmr = B.x * approx_rcp(w);
        float f;
        uint32_t m;
        if constexpr (COUNT) {
            const bool accept = !(d2 <= B.y);                               // leaf || size/(d+eps) < theta
            const bool use = accept && (selfn[b] != idx);
            f = use ? gmr : 0.f;
            const uint32_t fl = a.flags[idx];
            c_vis += active;
            c_int += (active && use && (fl & kNodeNonZero));
            c_open += !accept;
            m = __ballot_sync(0xffffffffu, !accept);
        } else {
            // accept = !(d2 <= thr); use = accept && not the body's own leaf; f = use ? gmr : 0; ballot(!accept)
            asm volatile("{\n .reg .pred pa, pu;\n setp.gtu.f32 pa, %2, %3;\n setp.ne.and.u32 pu, %4, %5, pa;\n"
                         " selp.f32 %0, %6, 0f00000000, pu;\n vote.sync.ballot.b32 %1, !pa, 0xffffffff;\n}"
                         : "=f"(f), "=r"(m) : "f"(d2), "f"(B.y), "r"(selfn[b]), "r"(idx), "f"(gmr));
        }
        acc2[b] = __ffma2_rn(make_float2(f, f), d, acc2[b]);
        return m;
    };

    // EXACT: is cell `idx` (warp-uniform) a multi-body leaf at the depth cap with mass > mass_eps?  If so apply
    // its bodies one by one to the lanes whose body reached it (act[b]) and return true.
    [[maybe_unused]] auto exact_leaf = [&](uint32_t idx, const bool (&act)[BPL]) -> bool {
        if (idx < a.finest_off) return false;
        const uint32_t cnt = __ldg(a.t_count + idx);
        if (cnt < 2u || !(__ldg(a.flags + idx) & kNodeNonZero)) return false;
        const uint32_t first = __ldg(a.t_first + idx);
        for (uint32_t j = 0; j < cnt; ++j) {
            const uint32_t bj = __ldg(a.sidx + first + j);
            const double2 pj = a.pos_in[bj];
            const double sxj = pj.x * scale_d, syj = pj.y * scale_d;
            const float xh = (float)sxj, yh = (float)syj;
            const float2 jl = make_float2((float)(sxj - (double)xh), (float)(syj - (double)yh));
            const float gmj = (float)(Gs * a.mass[bj]);
#pragma unroll
            for (int b = 0; b < BPL; ++b) {
                const float2 d = __fadd2_rn(__fadd2_rn(make_float2(xh, yh), nh[b]), __fadd2_rn(jl, nl[b]));
                const float d2 = fmaf(d.x, d.x, d.y * d.y);
                const float w = d2 * (approx_sqrt(d2) + feps);
                const bool use = act[b] && (body[b] != bj);
                const float f = use ? gmj * approx_rcp(w) : 0.f;    // coincident j != i: inf * 0 = NaN, as the pair formula
                acc2[b] = __ffma2_rn(make_float2(f, f), d, acc2[b]);
                if constexpr (COUNT) c_int += use;
            }
        }
        if constexpr (COUNT) {
#pragma unroll
            for (int b = 0; b < BPL; ++b) c_vis += act[b];
        }
        return true;
    };

    int top = 0;
    {   // the root (project.cu:711-715 pushes node 0)
        const float4 A = __ldg(reinterpret_cast<const float4*>(a.rec));
        const float2 B = __ldg(reinterpret_cast<const float2*>(a.rec) + 2);
        uint32_t m[BPL];
        uint32_t any = 0;
        bool root_done = false;
        if constexpr (EXACT) {
            bool live[BPL];
#pragma unroll
            for (int b = 0; b < BPL; ++b) live[b] = body[b] != 0xffffffffu;
            root_done = exact_leaf(0u, live);
        }
#pragma unroll
        for (int b = 0; b < BPL; ++b) {
            const bool live = body[b] != 0xffffffffu && !rescue[b];
            const float2 mh = live ? nh[b] : make_float2(kFarLane, kFarLane);
            m[b] = root_done ? 0u : eval(A, B, 0u, b, mh, live);
            any |= m[b];
        }
        if (any) {
            if (lane == 0) SE::store(sbase, 0u, m);
            top = 1;
        }
        __syncwarp();
    }
    while (top > 0) {
        --top;
        uint32_t node, pm[BPL];
        SE::load(sbase + (uint32_t)top * SE::kBytes, node, pm);
        __syncwarp();
        const uint32_t base = 4u * node + 1u;
        const NodeRec* __restrict__ rp = a.rec + base;
        float2 mh[BPL];
        bool active[BPL];
#pragma unroll
        for (int b = 0; b < BPL; ++b) {
            active[b] = (pm[b] >> lane) & 1u;
            mh[b] = active[b] ? nh[b] : make_float2(kFarLane, kFarLane);
        }
        if constexpr (COUNT) c_steps += (lane == 0);
#pragma unroll
        for (uint32_t q = 0; q < 4; ++q) {
            const float4 A = __ldg(reinterpret_cast<const float4*>(rp + q));        // chx chy clx cly
            const float2 B = __ldg(reinterpret_cast<const float2*>(rp + q) + 2);    // gm thr
            uint32_t m[BPL];
            uint32_t any = 0;
            if constexpr (EXACT) {
                if (exact_leaf(base + q, active)) continue;                         // a leaf: nobody opens it
            }
#pragma unroll
            for (int b = 0; b < BPL; ++b) {
                m[b] = eval(A, B, base + q, b, mh[b], active[b]);
                any |= m[b];
            }
            // branch-free push: one straight-line block per step keeps four independent chains in flight
            SE::store_if((any != 0u) & (lane == 0), sbase + (uint32_t)top * SE::kBytes, base + q, m);
            top += (any != 0u);
        }
        __syncwarp();
    }

#pragma unroll
    for (int b = 0; b < BPL; ++b) {
        if (body[b] != 0xffffffffu) {
            const double2 p = a.pos_in[body[b]];
            const double mi = a.mass[body[b]];
            double fx = (double)acc2[b].x, fy = (double)acc2[b].y;
            if (rescue[b]) fp64_body_walk(a, selfn[b], p.x, p.y, fx, fy);
            finish_body<INTEGRATE>(a, body[b], p.x, p.y, mi, mi * fx, mi * fy);
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

// ------------------------------------------------------------------------------------------------
// FP32 traversal, two bodies per lane with the FP32x2 instructions packed ACROSS the two bodies:
// every arithmetic step of the evaluation (displacement, d2, d2 (d + eps), G M / w, accumulation)
// is one packed instruction for both bodies; only the two MUFU pairs and the predicate tails are
// per body.  Same semantics as traverse_f32_kernel<2, ...>; this is the production variant.
// ------------------------------------------------------------------------------------------------
// EXACT_EPS = false (default): 1 / (d + eps) is taken to first order in eps / d, one MUFU.RSQ per body
// (the SFU pipe, 16 lanes/clk/SM on B200, is this kernel's scarcest resource: ncu math_pipe_throttle);
// relative error (eps/d)^2, i.e. < 1e-6 for separations above 1e-12 (eps = 1e-15).
// EXACT_EPS = true (BH_FLAG_EXACT_EPS): MUFU.SQRT + MUFU.RCP, exact for any separation.
// LEAVES (BH_FLAG_EXACT_LEAVES together with reserved[0] == 2): the exact-leaves extension (see
// traverse_f32_kernel) in the pair kernel — the member loop is packed across the lane's two bodies as well.
// Measured and removed in round 2 (warm A/B at N = 1M, 407 us baseline, profiles/r02_traverse_variants_ab.txt): an L1
// prefetch of a pushed child's children (433 us), an SM-local block order (417 us), both (438 us), and a
// software-pipelined pop + fetch of the next cell between the test and the force phase (608 us, 121 registers).
template <bool INTEGRATE, bool EXACT_EPS, bool LEAVES = false>
__global__ void __launch_bounds__(kTravThreads, kPairMinBlocks * (256 / kTravThreads))
traverse_f32_pair_kernel(const __grid_constant__ TravArgs a) {
    using SE = StackEntry<2>;
    __shared__ __align__(16) uint8_t s_stack[kTravWarps][kStackCap * SE::kBytes];
    pdl_entry();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warp_slot0 = ((int64_t)blockIdx.x * kTravWarps + warp) * 64;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(&s_stack[warp][0]);

    uint32_t body[2], selfn[2];
    float2 nxh, nyh, nxl, nyl;   // minus the scaled positions of body 0 (.x) and body 1 (.y), hi / lo floats
    float2 accx = make_float2(0.f, 0.f), accy = make_float2(0.f, 0.f);
    const float feps = a.consts->feps;
    {
        const double scale = a.consts->scale;
        float t[2][4];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int64_t slot = warp_slot0 + b * 32 + lane;
            body[b] = 0xffffffffu; selfn[b] = 0xffffffffu;
            double px = 0.0, py = 0.0;
            if (slot < a.n_slots) {
                uint32_t sp = a.own_list ? a.own_list[slot] : (uint32_t)slot;
                body[b] = a.sidx[sp];
                selfn[b] = a.self_node[body[b]];
                double2 p = a.pos_in[body[b]];
                px = p.x; py = p.y;
            }
            const double sx = px * scale, sy = py * scale;
            const float xh = (float)sx, yh = (float)sy;
            t[b][0] = -xh; t[b][1] = -yh;
            t[b][2] = -(float)(sx - (double)xh); t[b][3] = -(float)(sy - (double)yh);
        }
        nxh = make_float2(t[0][0], t[1][0]); nyh = make_float2(t[0][1], t[1][1]);
        nxl = make_float2(t[0][2], t[1][2]); nyl = make_float2(t[0][3], t[1][3]);
    }
    const float2 eps2 = make_float2(feps, feps), neg_eps2 = make_float2(-feps, -feps);
    (void)eps2; (void)neg_eps2;
    [[maybe_unused]] double scale_d = 0.0, Gs = 0.0;   // LEAVES: coordinate scale and G * scale^2
    if constexpr (LEAVES) { scale_d = a.consts->scale; Gs = a.G * scale_d * scale_d; }

    // LEAVES: if cell `idx` (warp-uniform) is a multi-body leaf at the depth cap with mass > mass_eps, apply its bodies
    // one by one to the lane's bodies that reached it (act0 / act1), self excluded, and return true.
    [[maybe_unused]] auto exact_leaf = [&](uint32_t idx, bool act0, bool act1) -> bool {
        if (idx < a.finest_off) return false;
        const uint32_t cnt = __ldg(a.t_count + idx);
        if (cnt < 2u || !(__ldg(a.flags + idx) & kNodeNonZero)) return false;
        const uint32_t first = __ldg(a.t_first + idx);
        for (uint32_t j = 0; j < cnt; ++j) {
            const uint32_t bj = __ldg(a.sidx + first + j);
            const double2 pj = a.pos_in[bj];
            const double sxj = pj.x * scale_d, syj = pj.y * scale_d;
            const float xh = (float)sxj, yh = (float)syj;
            const float xl = (float)(sxj - (double)xh), yl = (float)(syj - (double)yh);
            const float gmj = (float)(Gs * a.mass[bj]);
            const float2 dx = __fadd2_rn(__fadd2_rn(make_float2(xh, xh), nxh), __fadd2_rn(make_float2(xl, xl), nxl));
            const float2 dy = __fadd2_rn(__fadd2_rn(make_float2(yh, yh), nyh), __fadd2_rn(make_float2(yl, yl), nyl));
            const float2 d2 = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
            float2 g;
            if constexpr (EXACT_EPS) {
                const float2 w = __fmul2_rn(d2, __fadd2_rn(make_float2(approx_sqrt(d2.x), approx_sqrt(d2.y)), eps2));
                g = __fmul2_rn(make_float2(gmj, gmj), make_float2(approx_rcp(w.x), approx_rcp(w.y)));
            } else {
                const float2 inv = make_float2(approx_rsqrt(d2.x), approx_rsqrt(d2.y));
                const float2 t = __fmul2_rn(inv, inv);
                const float2 u = __ffma2_rn(neg_eps2, t, inv);
                g = __fmul2_rn(make_float2(gmj, gmj), __fmul2_rn(t, u));
            }
            // the select comes last: a body's own entry has d2 == 0 (g is NaN there) and must contribute exactly 0
            const float2 f = make_float2((act0 && body[0] != bj) ? g.x : 0.f, (act1 && body[1] != bj) ? g.y : 0.f);
            accx = __ffma2_rn(f, dx, accx);
            accy = __ffma2_rn(f, dy, accy);
        }
        return true;
    };

    // Evaluate one node for both bodies of this lane; returns the two ballots of "opens".
    auto eval = [&](const float4 A, const float2 B, uint32_t idx, const float2 mxh, const float2 myh, uint32_t& m0,
                    uint32_t& m1) {
        const float2 dx = __fadd2_rn(__fadd2_rn(make_float2(A.x, A.x), mxh), __fadd2_rn(make_float2(A.z, A.z), nxl));
        const float2 dy = __fadd2_rn(__fadd2_rn(make_float2(A.y, A.y), myh), __fadd2_rn(make_float2(A.w, A.w), nyl));
        const float2 d2 = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
        // G M / (d2 (d + eps)), project.cu:765-769; d2 == 0 -> inf, times dx == 0 -> NaN like the reference
        float2 g;
        if constexpr (EXACT_EPS) {
            const float2 w = __fmul2_rn(d2, __fadd2_rn(make_float2(approx_sqrt(d2.x), approx_sqrt(d2.y)), eps2));
            g = __fmul2_rn(make_float2(B.x, B.x), make_float2(approx_rcp(w.x), approx_rcp(w.y)));
        } else {
            const float2 inv = make_float2(approx_rsqrt(d2.x), approx_rsqrt(d2.y));
            const float2 t = __fmul2_rn(inv, inv);                       // 1 / d2
            const float2 u = __ffma2_rn(neg_eps2, t, inv);               // 1 / (d + eps) ~= 1/d - eps/d2
            g = __fmul2_rn(make_float2(B.x, B.x), __fmul2_rn(t, u));
        }
        float2 f;
        // per body: accept = !(d2 <= thr); use = accept && not the body's own leaf; f = use ? g : 0; ballot(!accept)
        asm volatile("{\n .reg .pred pa, pu;\n setp.gtu.f32 pa, %2, %3;\n setp.ne.and.u32 pu, %4, %5, pa;\n"
                     " selp.f32 %0, %6, 0f00000000, pu;\n vote.sync.ballot.b32 %1, !pa, 0xffffffff;\n}"
                     : "=f"(f.x), "=r"(m0) : "f"(d2.x), "f"(B.y), "r"(selfn[0]), "r"(idx), "f"(g.x));
        asm volatile("{\n .reg .pred pa, pu;\n setp.gtu.f32 pa, %2, %3;\n setp.ne.and.u32 pu, %4, %5, pa;\n"
                     " selp.f32 %0, %6, 0f00000000, pu;\n vote.sync.ballot.b32 %1, !pa, 0xffffffff;\n}"
                     : "=f"(f.y), "=r"(m1) : "f"(d2.y), "f"(B.y), "r"(selfn[1]), "r"(idx), "f"(g.y));
        accx = __ffma2_rn(f, dx, accx);
        accy = __ffma2_rn(f, dy, accy);
    };

    // Depth-first walk over a warp-shared stack.  (Keeping the first opened child in registers and
    // prefetching its children was tried — profiles/r01_traverse_v7_pair_ncu_summary.txt — and lost:
    // +28 % instructions for the bookkeeping, no latency gain.)
    uint32_t sp = sbase;              // shared-memory address of the next free stack slot
    {   // the root (project.cu:711-715 pushes node 0)
        const float4 A = __ldg(reinterpret_cast<const float4*>(a.rec));
        const float2 B = __ldg(reinterpret_cast<const float2*>(a.rec) + 2);
        const bool l0 = body[0] != 0xffffffffu, l1 = body[1] != 0xffffffffu;
        const float2 mxh = make_float2(l0 ? nxh.x : kFarLane, l1 ? nxh.y : kFarLane);
        const float2 myh = make_float2(l0 ? nyh.x : kFarLane, l1 ? nyh.y : kFarLane);
        uint32_t m[2] = {0u, 0u};
        bool root_done = false;
        if constexpr (LEAVES) root_done = exact_leaf(0u, l0, l1);
        if (!root_done) eval(A, B, 0u, mxh, myh, m[0], m[1]);
        const bool push = (m[0] | m[1]) != 0u;
        SE::store_if(push & (lane == 0), sp, 0u, m);
        sp += push ? SE::kBytes : 0u;
    }
    while (sp != sbase) {
        sp -= SE::kBytes;
        uint32_t node, pm[2];
        __syncwarp();             // lane 0's stores of earlier steps are visible to the whole warp
        SE::load(sp, node, pm);
        __syncwarp();             // nobody overwrites the slot before everybody has read it
        const uint32_t base = 4u * node + 1u;
        const NodeRec* __restrict__ rp = a.rec + base;
        const bool a0 = (pm[0] >> lane) & 1u, a1 = (pm[1] >> lane) & 1u;
        const float2 mxh = make_float2(a0 ? nxh.x : kFarLane, a1 ? nxh.y : kFarLane);
        const float2 myh = make_float2(a0 ? nyh.x : kFarLane, a1 ? nyh.y : kFarLane);
#pragma unroll
        for (uint32_t q = 0; q < 4; ++q) {
            const float4 A = __ldg(reinterpret_cast<const float4*>(rp + q));        // chx chy clx cly
            const float2 B = __ldg(reinterpret_cast<const float2*>(rp + q) + 2);    // gm thr
            uint32_t m[2];
            if constexpr (LEAVES) {
                if (exact_leaf(base + q, a0, a1)) continue;                          // a leaf: nobody opens it
            }
            eval(A, B, base + q, mxh, myh, m[0], m[1]);
            // branch-free push: one straight-line block per step keeps four independent chains in flight
            const bool push = (m[0] | m[1]) != 0u;
            SE::store_if(push & (lane == 0), sp, base + q, m);
            sp += push ? SE::kBytes : 0u;
        }
    }
    const float ax[2] = {accx.x, accx.y}, ay[2] = {accy.x, accy.y};
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        if (body[b] != 0xffffffffu) {
            const double2 p = a.pos_in[body[b]];
            const double mi = a.mass[body[b]];
            finish_body<INTEGRATE>(a, body[b], p.x, p.y, mi, mi * (double)ax[b], mi * (double)ay[b]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// FP32 traversal, production variant from round 2 on ("list kernel"): group-classified walk in a warp-local frame.
//
// The pair kernel above lets all 32 lanes test the SAME node (every lane = two bodies), so a warp spends a full
// 37-instruction evaluation on every node it touches — also on the ~13 % of nodes that every body of the warp opens
// and on the ~70 % that every body of the (sub)group accepts, where the per-body test is a foregone conclusion.
// Here the TEST is done lane-per-NODE first: one round pops up to 8 cells from the warp's stack and hands their 32
// children to the 32 lanes; every lane compares its child with the bounding box of the warp's 64 bodies
// (conservative bounds on the distance of any body of the warp to the node's centre of mass):
//
//     all bodies certainly accept   (dmin^2 > thr (1 + delta), or the node is a leaf)  -> class A: force only
//     all bodies certainly open     (dmax^2 <= thr (1 - delta))                        -> class O: push, no arithmetic
//     zero-mass / empty node        (the reference skips it, project.cu:731)           -> dropped
//     anything else                                                                   -> class M: per-body test
//
// Class-A nodes go to staging lists in shared memory and are then applied to all 64 bodies with the force
// arithmetic only (no test, no predicate, no ballot).  Class-M nodes are evaluated like in the pair kernel (per-body
// test + force + ballot of the opening bodies).  A stack entry is (cell, mask of the bodies that opened it); the
// children of a cell are only ever applied to the bodies of its mask, so EACH BODY STILL SEES EXACTLY THE NODE SET
// OF THE REFERENCE'S PER-BODY DFS: the box test only short-cuts decisions that the per-body test
// `!(d2 <= thr)` would take identically for every body of the mask (the margins cover the FP32 rounding of both).
// A single-body leaf is the own leaf of at most one body of the warp: its bit is cleared from the entry's mask
// (the reference's self test, project.cu:760) by the classifying lane, which knows the body's sorted position from
// the record (`first`).
//
// Warp-local frame.  All coordinates are taken relative to the centre O of the warp's bounding box (FP64
// subtraction by the classifying lane / once per body), so a class-A node that is FAR from the box (dmin^2 >=
// diag^2 / 64, 85 % of them) needs no double-float arithmetic at all: fl(C - O) - fl(x - O) has a relative error of
// at most 2^-24 (2 + 2 diag / d) <= 1.1e-6 per interaction, and the displacement is ONE packed add per coordinate
// instead of three.  NEAR class-A nodes and class-M nodes keep the double-float displacement (2^-48 of the frame),
// which the dominant close interactions need.
// Measured at N = 1M (uniform disk, cap 10), per warp of 64 bodies: 505 nodes touched, of which 20 zero-mass, 61
// class O, 348 class A (about 290 far), 76 class M (tools/classify_nodes.c reproduces these counts on the CPU).
// ------------------------------------------------------------------------------------------------
constexpr int kListStackCap = 128;         // entries; rounds shrink to plain DFS (1 cell) above kListStackSoft
constexpr int kListStackSoft = 88;
constexpr float kListDelta = 1e-4f;        // relative safety band of the box test on d^2 (FP32 rounding is ~1e-6)

struct ListWarpConsts {                    // per warp, written once, read by every round's classification
    float oxh, oxl, oyh, oyl;              // frame origin (scaled units), exact as hi + lo
    float bx0, bx1, by0, by1;              // the warp's bounding box in the local frame
    float slack, far2;                     // absolute slack of the box test; diag^2 / 64
    uint32_t live0, live1;                 // lanes whose body 0 / body 1 exists
};

__device__ __forceinline__ void sts_v4_if(bool doit, uint4* p, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %0, 0;\n @p st.shared.v4.u32 [%1], {%2, %3, %4, %5};\n}" ::"r"((uint32_t)doit),
                 "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

#ifndef BH_LIST_MIN_BLOCKS
#define BH_LIST_MIN_BLOCKS 8    // per 128 threads: <= 72 registers, 28 warps per SM
#endif
// Entries per iteration of the far partial-mask loop (its list holds ~10 entries per round: x4 measured 305 -> 302 us at
// 1M, 824 -> 806 us at 4M) and of the near loop (~2.4 entries per round: x4 would mostly run in the remainder code).
#ifndef BH_FARM_UNROLL
#define BH_FARM_UNROLL 4
#endif
#ifndef BH_NEAR_UNROLL
#define BH_NEAR_UNROLL 2
#endif
constexpr int kNearUnroll = BH_NEAR_UNROLL;
constexpr int kFarMaskedUnroll = BH_FARM_UNROLL;
// Far class-A nodes without the distance offset: G M / (d^2 (d + eps)) differs from G M / d^3 by eps / d relative,
// and a node only counts as FAR when d >= 2^25 eps (folded into the warp's far2 threshold), i.e. below half an FP32
// ulp — one packed operation less per far entry (3 FMUL2 instead of FMUL2 + FFMA2 + 2 FMUL2): 317 -> 305 us at 1M
// (profiles/r02_far_no_eps_ab.txt).  0 = the expression with the offset, as in the near / mixed loops.
#ifndef BH_FAR_NO_EPS
#define BH_FAR_NO_EPS 1
#endif
// LEAVES (BH_FLAG_EXACT_LEAVES on whole-set launches over >= kTwoBodiesPerLaneMin bodies, or reserved[0] = 9; a separate
// instantiation, the production kernel's code is unchanged): a multi-body leaf at the depth cap is a seventh class.  It is not staged as one monopole but queued with
// its mask; after the round's lists have been applied, its bodies are converted into the warp-local frame by one lane
// each (32 members per pass), staged like near class-A nodes and applied to the bodies of the mask, self excluded by
// body index — the same pair expression as the pair kernel's member loop, ~26 instead of ~45 instructions per member.
template <bool INTEGRATE, bool EXACT_EPS, bool LEAVES = false>
__global__ void __launch_bounds__(kTravThreads, BH_LIST_MIN_BLOCKS)
traverse_f32_list_kernel(const __grid_constant__ TravArgs a) {
    __shared__ __align__(16) uint4 s_leaf[LEAVES ? kTravWarps : 1][LEAVES ? 32 : 1];   // (first, count, mask0, mask1)
    __shared__ __align__(16) uint4 s_stack[kTravWarps][kListStackCap];   // (cell, mask0, mask1, -)
    __shared__ __align__(16) float4 s_nodeA[kTravWarps][32];             // staged nodes, local frame: far class A: cx cy gm - ;
                                                                         //   near class A / class M: chx chy clx cly
    __shared__ __align__(16) uint4 s_nodeB[kTravWarps][32];              // gm thr mask0 mask1 (bit patterns)
    __shared__ uint32_t s_cell[kTravWarps][32];                          //              pyramid index (class M only)
    __shared__ __align__(16) ListWarpConsts s_wc[kTravWarps];
    pdl_entry();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4* const stack = s_stack[warp];
    float4* const nodeA = s_nodeA[warp];
    uint4* const nodeB = s_nodeB[warp];
    uint32_t* const cellv = s_cell[warp];
    const float feps = a.consts->feps;
    const uint32_t n_tiles = (uint32_t)((a.n_slots + 63) >> 6);
    // Persistent warps: every warp fetches 64-body tiles from a global queue until it is empty, so a warp slot is
    // never idle while work is left (with one tile per warp and 4-warp blocks, a block's slot stayed occupied by its
    // slowest warp: ncu showed 23 of 28 possible warps resident on average).
  for (;;) {
    uint32_t tile = 0;
    if (lane == 0) tile = atomicAdd(a.tile_queue, 1u);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= n_tiles) {
        if (lane == 0) {
            const uint32_t total = gridDim.x * kTravWarps;
            if (atomicAdd(a.tile_queue + 1, 1u) == total - 1u) { a.tile_queue[0] = 0u; a.tile_queue[1] = 0u; }
        }
        return;
    }
    const int64_t warp_slot0 = (int64_t)tile * 64;

    uint32_t body[2], selfn[2], spos[2];
    float2 nxh, nyh, nxl, nyl;   // minus the local-frame positions of body 0 (.x) and body 1 (.y), hi / lo floats
    float2 accx = make_float2(0.f, 0.f), accy = make_float2(0.f, 0.f);
    {
        const double scale = a.consts->scale;
        double sx[2], sy[2];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int64_t slot = warp_slot0 + b * 32 + lane;
            body[b] = 0xffffffffu; selfn[b] = 0xffffffffu; spos[b] = 0u;
            double px = 0.0, py = 0.0;
            if (slot < a.n_slots) {
                uint32_t sp = a.own_list ? a.own_list[slot] : (uint32_t)slot;
                spos[b] = sp;
                body[b] = a.sidx[sp];
                selfn[b] = a.self_node[body[b]];
                double2 p = a.pos_in[body[b]];
                px = p.x; py = p.y;
            }
            sx[b] = px * scale; sy[b] = py * scale;
        }
        bool l0 = body[0] != 0xffffffffu, l1 = body[1] != 0xffffffffu;
        // (first pass over the frame: the box of ALL bodies; bodies whose own cell the local frame cannot resolve are then
        // taken out of the walk — see fp64_body_walk — which can only shrink the box)
        const uint32_t live0 = __ballot_sync(0xffffffffu, l0), live1 = __ballot_sync(0xffffffffu, l1);
        // bounding box of the warp's bodies (floats rounded outwards by the slack below), frame origin = its centre
        const float inf = __int_as_float(0x7f800000);
        float x0 = inf, x1 = -inf, y0 = inf, y1 = -inf;
        if (l0) { x0 = x1 = (float)sx[0]; y0 = y1 = (float)sy[0]; }
        if (l1) { x0 = fminf(x0, (float)sx[1]); x1 = fmaxf(x1, (float)sx[1]); y0 = fminf(y0, (float)sy[1]); y1 = fmaxf(y1, (float)sy[1]); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, o)); x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, o));
            y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, o)); y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, o));
        }
        const double ox = 0.5 * ((double)x0 + (double)x1), oy = 0.5 * ((double)y0 + (double)y1);
        float t[2][4];
        bool resc[2] = {false, false};
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const double rx = sx[b] - ox, ry = sy[b] - oy;
            if (!LEAVES && body[b] != 0xffffffffu)   // (LEAVES: a body's own multi-body cell is never applied to it as a monopole)
                // (the node records themselves carry the COM as a double-float of the GLOBAL scaled coordinate: that, not
                // the local frame, bounds what the near-field arithmetic resolves)
                resc[b] = own_cell_unresolved(a, spos[b], sx[b] / scale, sy[b] / scale, scale, sx[b] * sx[b] + sy[b] * sy[b]);
            const float xh = (float)rx, yh = (float)ry;
            t[b][0] = -xh; t[b][1] = -yh;
            t[b][2] = -(float)(rx - (double)xh); t[b][3] = -(float)(ry - (double)yh);
        }
        nxh = make_float2(t[0][0], t[1][0]); nyh = make_float2(t[0][1], t[1][1]);
        nxl = make_float2(t[0][2], t[1][2]); nyl = make_float2(t[0][3], t[1][3]);
        if (lane == 0) {
            ListWarpConsts wc;
            wc.oxh = (float)ox; wc.oxl = (float)(ox - (double)wc.oxh);     // midpoint of two floats: hi + lo is exact
            wc.oyh = (float)oy; wc.oyl = (float)(oy - (double)wc.oyh);
            // the float box was formed from rounded coordinates: |true - float| <= 2^-24 |coordinate|
            const float mag = fmaxf(fmaxf(fabsf(x0), fabsf(x1)), fmaxf(fabsf(y0), fabsf(y1)));
            wc.bx0 = (float)((double)x0 - ox); wc.bx1 = (float)((double)x1 - ox);
            wc.by0 = (float)((double)y0 - oy); wc.by1 = (float)((double)y1 - oy);
            wc.slack = 1.2e-7f * mag;
            const float wx = wc.bx1 - wc.bx0, wy = wc.by1 - wc.by0;
            wc.far2 = 0.015625f * fmaf(wx, wx, wy * wy);
#if BH_FAR_NO_EPS
            if constexpr (!EXACT_EPS) { const float dfar = 33554432.f * feps; wc.far2 = fmaxf(wc.far2, dfar * dfar); }
#endif
            s_wc[warp] = wc;
        }
        // bodies the frame cannot resolve leave the walk: bit 31 of spos marks them for the epilogue
        const uint32_t walk0 = __ballot_sync(0xffffffffu, l0 && !resc[0]), walk1 = __ballot_sync(0xffffffffu, l1 && !resc[1]);
        if (lane == 0) { s_wc[warp].live0 = walk0; s_wc[warp].live1 = walk1; }
        spos[0] = resc[0] ? 1u : 0u; spos[1] = resc[1] ? 1u : 0u;
        __syncwarp();
    }
    const float2 eps2 = make_float2(feps, feps), neg_eps2 = make_float2(-feps, -feps);
    (void)eps2; (void)neg_eps2;
    const uint32_t lanebit = 1u << lane;

    // G M / (d2 (d + eps)) for both bodies of the lane (project.cu:765-769); d2 == 0 -> NaN like the reference
    auto gfactor = [&](const float2 d2, const float gm) -> float2 {
        if constexpr (EXACT_EPS) {
            const float2 w = __fmul2_rn(d2, __fadd2_rn(make_float2(approx_sqrt(d2.x), approx_sqrt(d2.y)), eps2));
            return __fmul2_rn(make_float2(gm, gm), make_float2(approx_rcp(w.x), approx_rcp(w.y)));
        } else {
            const float2 inv = make_float2(approx_rsqrt(d2.x), approx_rsqrt(d2.y));
            const float2 t = __fmul2_rn(inv, inv);                       // 1 / d2
            const float2 u = __ffma2_rn(neg_eps2, t, inv);               // 1 / (d + eps) ~= 1/d - eps/d2
            return __fmul2_rn(make_float2(gm, gm), __fmul2_rn(t, u));
        }
    };
    // far nodes (d >= 2^25 eps by the far2 threshold): G M / d^3, the offset is below the FP32 rounding
    auto gfactor_far = [&](const float2 d2, const float gm) -> float2 {
#if BH_FAR_NO_EPS
        if constexpr (!EXACT_EPS) {
            const float2 inv = make_float2(approx_rsqrt(d2.x), approx_rsqrt(d2.y));
            return __fmul2_rn(__fmul2_rn(make_float2(gm, gm), inv), __fmul2_rn(inv, inv));
        } else
#endif
            return gfactor(d2, gm);
    };
    // far class A: single-float displacement in the local frame
    auto apply_far = [&](const float4 F) {
        const float2 dx = __fadd2_rn(make_float2(F.x, F.x), nxh), dy = __fadd2_rn(make_float2(F.y, F.y), nyh);
        const float2 d2 = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
        const float2 g = gfactor_far(d2, F.z);
        accx = __ffma2_rn(g, dx, accx);
        accy = __ffma2_rn(g, dy, accy);
    };
    auto apply_far_masked = [&](const float4 F, const uint2 m) {
        const float2 dx = __fadd2_rn(make_float2(F.x, F.x), nxh), dy = __fadd2_rn(make_float2(F.y, F.y), nyh);
        const float2 d2 = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
        const float2 g = gfactor_far(d2, F.z);
        const float2 f = make_float2((m.x & lanebit) ? g.x : 0.f, (m.y & lanebit) ? g.y : 0.f);
        accx = __ffma2_rn(f, dx, accx);
        accy = __ffma2_rn(f, dy, accy);
    };
    // near class A: double-float displacement; the select comes last — a body's own leaf has d2 == 0
    auto apply_near = [&](const float4 A, const float gm, const uint32_t m0, const uint32_t m1) {
        const float2 dx = __fadd2_rn(__fadd2_rn(make_float2(A.x, A.x), nxh), __fadd2_rn(make_float2(A.z, A.z), nxl));
        const float2 dy = __fadd2_rn(__fadd2_rn(make_float2(A.y, A.y), nyh), __fadd2_rn(make_float2(A.w, A.w), nyl));
        const float2 d2 = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
        const float2 g = gfactor(d2, gm);
        const float2 f = make_float2((m0 & lanebit) ? g.x : 0.f, (m1 & lanebit) ? g.y : 0.f);
        accx = __ffma2_rn(f, dx, accx);
        accy = __ffma2_rn(f, dy, accy);
    };
    // class M: the per-body test of the pair kernel; returns the ballots of the bodies that open the node
    auto eval_mixed = [&](const float4 A, const float gm, const float thr, const uint32_t idx, const uint32_t pm0,
                          const uint32_t pm1, uint32_t& m0, uint32_t& m1) {
        const bool a0 = pm0 & lanebit, a1 = pm1 & lanebit;
        const float2 mxh = make_float2(a0 ? nxh.x : kFarLane, a1 ? nxh.y : kFarLane);
        const float2 myh = make_float2(a0 ? nyh.x : kFarLane, a1 ? nyh.y : kFarLane);
        const float2 dx = __fadd2_rn(__fadd2_rn(make_float2(A.x, A.x), mxh), __fadd2_rn(make_float2(A.z, A.z), nxl));
        const float2 dy = __fadd2_rn(__fadd2_rn(make_float2(A.y, A.y), myh), __fadd2_rn(make_float2(A.w, A.w), nyl));
        const float2 d2 = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
        const float2 g = gfactor(d2, gm);
        float2 f;
        // per body: accept = !(d2 <= thr); use = accept && not the body's own leaf; f = use ? g : 0; ballot(!accept)
        asm volatile("{\n .reg .pred pa, pu;\n setp.gtu.f32 pa, %2, %3;\n setp.ne.and.u32 pu, %4, %5, pa;\n"
                     " selp.f32 %0, %6, 0f00000000, pu;\n vote.sync.ballot.b32 %1, !pa, 0xffffffff;\n}"
                     : "=f"(f.x), "=r"(m0) : "f"(d2.x), "f"(thr), "r"(selfn[0]), "r"(idx), "f"(g.x));
        asm volatile("{\n .reg .pred pa, pu;\n setp.gtu.f32 pa, %2, %3;\n setp.ne.and.u32 pu, %4, %5, pa;\n"
                     " selp.f32 %0, %6, 0f00000000, pu;\n vote.sync.ballot.b32 %1, !pa, 0xffffffff;\n}"
                     : "=f"(f.y), "=r"(m1) : "f"(d2.y), "f"(thr), "r"(selfn[1]), "r"(idx), "f"(g.y));
        accx = __ffma2_rn(f, dx, accx);
        accy = __ffma2_rn(f, dy, accy);
    };
    // a record's centre of mass (ch + cl) in the local frame, as a double-float pair: TwoSum of the hi parts (exact),
    // the lo parts and the rounding error folded in, then renormalised — FP32 adds only (the FP64 pipe and its
    // conversions share the SFU issue port with the MUFUs of the force loops)
    auto to_local = [&](const float4 R, const ListWarpConsts& wc) -> float4 {
        float4 out;
        {
            const float a_ = R.x, b_ = -wc.oxh;
            const float s_ = a_ + b_, bb = s_ - a_;
            const float e_ = (a_ - (s_ - bb)) + (b_ - bb);
            const float t_ = (R.z - wc.oxl) + e_;
            out.x = s_ + t_;
            out.z = t_ - (out.x - s_);
        }
        {
            const float a_ = R.y, b_ = -wc.oyh;
            const float s_ = a_ + b_, bb = s_ - a_;
            const float e_ = (a_ - (s_ - bb)) + (b_ - bb);
            const float t_ = (R.w - wc.oyl) + e_;
            out.y = s_ + t_;
            out.w = t_ - (out.y - s_);
        }
        return out;
    };

    // LEAVES: the bodies first .. first + cnt of the sorted order (one cap-level cell) act one by one on the bodies of
    // the mask, self excluded (oracle/bh_oracle.c: bho_compute_forces_exact_leaves).  Uses the staging arrays: call it
    // only when nobody reads them any more.
    [[maybe_unused]] auto apply_leaf_members = [&](const uint32_t first, const uint32_t cnt, const uint32_t pm0, const uint32_t pm1) {
        const double scale_d = a.consts->scale, Gs = a.G * scale_d * scale_d;
        const double ox = (double)s_wc[warp].oxh + (double)s_wc[warp].oxl, oy = (double)s_wc[warp].oyh + (double)s_wc[warp].oyl;
        const bool a0 = pm0 & lanebit, a1 = pm1 & lanebit;
        for (uint32_t j0 = 0; j0 < cnt; j0 += 32u) {
            __syncwarp();                                   // the previous pass / the round's loops are done with the arrays
            const uint32_t j = j0 + (uint32_t)lane;
            if (j < cnt) {
                const uint32_t bj = __ldg(a.sidx + first + j);
                const double2 pj = a.pos_in[bj];
                const double rx = pj.x * scale_d - ox, ry = pj.y * scale_d - oy;
                const float xh = (float)rx, yh = (float)ry;
                nodeA[lane] = make_float4(xh, yh, (float)(rx - (double)xh), (float)(ry - (double)yh));
                nodeB[lane] = make_uint4(__float_as_uint((float)(Gs * a.mass[bj])), bj, 0u, 0u);
            }
            __syncwarp();
            const int m = (int)(cnt - j0 < 32u ? cnt - j0 : 32u);
#pragma unroll 2
            for (int i = 0; i < m; ++i) {
                const float4 A = nodeA[i];
                const uint4 MB = nodeB[i];
                const float2 dx = __fadd2_rn(__fadd2_rn(make_float2(A.x, A.x), nxh), __fadd2_rn(make_float2(A.z, A.z), nxl));
                const float2 dy = __fadd2_rn(__fadd2_rn(make_float2(A.y, A.y), nyh), __fadd2_rn(make_float2(A.w, A.w), nyl));
                const float2 d2 = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
                const float2 g = gfactor(d2, __uint_as_float(MB.x));
                // the select comes last: a body's own entry has d2 == 0 (g is NaN there) and must contribute exactly 0
                const float2 f = make_float2((a0 && body[0] != MB.y) ? g.x : 0.f, (a1 && body[1] != MB.y) ? g.y : 0.f);
                accx = __ffma2_rn(f, dx, accx);
                accy = __ffma2_rn(f, dy, accy);
            }
        }
        __syncwarp();
    };

    int top = 0;
    {   // the root (project.cu:711-715 pushes node 0): per-body test
        const float4 R = __ldg(reinterpret_cast<const float4*>(a.rec));
        const float2 B = __ldg(reinterpret_cast<const float2*>(a.rec) + 2);
        const ListWarpConsts& wc = s_wc[warp];
        uint32_t m0 = 0u, m1 = 0u;
        bool root_done = false;
        if constexpr (LEAVES) {
            const uint2 CF = __ldg(reinterpret_cast<const uint2*>(a.rec) + 3);   // count first
            if (B.y < 0.f && B.x != 0.f && CF.x >= 2u) {       // the root itself is the multi-body leaf (depth cap 1)
                apply_leaf_members(CF.y, CF.x, wc.live0, wc.live1);
                root_done = true;
            }
        }
        if (!root_done) eval_mixed(to_local(R, wc), B.x, B.y, 0u, wc.live0, wc.live1, m0, m1);
        if ((m0 | m1) != 0u) {
            if (lane == 0) stack[0] = make_uint4(0u, m0, m1, 0u);
            top = 1;
        }
        __syncwarp();
    }
    const uint32_t lanemask_lt = lanebit - 1u;
    while (top > 0) {
        // ---- pop up to 8 cells; lane l takes child (l & 3) of popped cell (l >> 2) ----
        int k = (kListStackSoft - top) / 3;
        k = k < 1 ? 1 : (k > 8 ? 8 : k);
        k = k > top ? top : k;
        const int e = lane >> 2;
        const bool valid = e < k;
        uint4 ent = make_uint4(0u, 0u, 0u, 0u);
        if (valid) ent = stack[top - 1 - e];
        top -= k;
        const uint32_t child = 4u * ent.x + 1u + (uint32_t)(lane & 3);
        float4 R = make_float4(0.f, 0.f, 0.f, 0.f);
        uint4 Bq = make_uint4(0u, 0u, 0u, 0u);
        if (valid) {
            R = __ldg(reinterpret_cast<const float4*>(a.rec + child));
            Bq = __ldg(reinterpret_cast<const uint4*>(a.rec + child) + 1);       // gm thr count first
        }
        const float gm = __uint_as_float(Bq.x), thr = __uint_as_float(Bq.y);
        uint32_t pm0 = ent.y, pm1 = ent.z;
        // ---- classify: 0 dropped, 1 far A full mask, 2 far A partial mask, 3 near A, 4 O, 5 M ----
        int cls = 0;
        float4 L = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid && gm != 0.f) {
            const ListWarpConsts wc = s_wc[warp];
            L = to_local(R, wc);
            const float ex = wc.slack + 1.2e-7f * fabsf(L.x), ey = wc.slack + 1.2e-7f * fabsf(L.y);
            const float nx = fmaxf(fmaxf(wc.bx0 - L.x, L.x - wc.bx1) - ex, 0.f), ny = fmaxf(fmaxf(wc.by0 - L.y, L.y - wc.by1) - ey, 0.f);
            const float dmin2 = fmaf(nx, nx, ny * ny);
            if (thr < 0.f) {                                  // a leaf: everybody in the mask accepts it
                cls = 1;
                if (Bq.z == 1u) {                             // single body: it may be one of ours -> the self test
                    if (a.own_list) cls = 5;                  // slots are indirect: let the per-body test decide
                    else {
                        const int64_t slot = (int64_t)Bq.w - warp_slot0;
                        if (slot >= 0 && slot < 64) { if (slot < 32) pm0 &= ~(1u << slot); else pm1 &= ~(1u << (slot - 32)); }
                    }
                } else if constexpr (LEAVES) {
                    cls = 6;                                  // several bodies (only possible at the depth cap): members
                }
            } else {
                const float fx = fmaxf(L.x - wc.bx0, wc.bx1 - L.x) + ex, fy = fmaxf(L.y - wc.by0, wc.by1 - L.y) + ey;
                const float dmax2 = fmaf(fx, fx, fy * fy);
                cls = (dmin2 > thr * (1.f + kListDelta)) ? 1 : (dmax2 <= thr * (1.f - kListDelta)) ? 4 : 5;
            }
            if (cls == 1) cls = !(dmin2 >= wc.far2) ? 3 : (pm0 != wc.live0 || pm1 != wc.live1) ? 2 : 1;
            if ((pm0 | pm1) == 0u) cls = 0;                   // (a lone body's own leaf)
        }
        const uint32_t b1 = __ballot_sync(0xffffffffu, cls == 1), b2 = __ballot_sync(0xffffffffu, cls == 2);
        const uint32_t b3 = __ballot_sync(0xffffffffu, cls == 3), b4 = __ballot_sync(0xffffffffu, cls == 4);
        const uint32_t b5 = __ballot_sync(0xffffffffu, cls == 5);
        const int n1 = __popc(b1), n2 = __popc(b2), n3 = __popc(b3), n5 = __popc(b5);
        sts_v4_if(cls == 4, stack + top + __popc(b4 & lanemask_lt), child, pm0, pm1, 0u);
        top += __popc(b4);
        [[maybe_unused]] int n6 = 0;
        if constexpr (LEAVES) {
            const uint32_t b6 = __ballot_sync(0xffffffffu, cls == 6);
            n6 = __popc(b6);
            sts_v4_if(cls == 6, s_leaf[warp] + __popc(b6 & lanemask_lt), Bq.w, Bq.z, pm0, pm1);
        }
        {   // one staging array for the four applied classes, in class order: far full | far masked | near | M
            const uint32_t mine = cls == 1 ? b1 : cls == 2 ? b2 : cls == 3 ? b3 : b5;
            const int base = cls == 1 ? 0 : cls == 2 ? n1 : cls == 3 ? n1 + n2 : n1 + n2 + n3;
            const int at = base + __popc(mine & lanemask_lt);
            const bool far = cls == 1 || cls == 2, staged = cls != 0 && cls != 4 && (!LEAVES || cls != 6);
            // far entries: (cx, cy, gm, -) so that one 128-bit load feeds the force loop; near / M: the double-float pair
            sts_v4_if(staged, reinterpret_cast<uint4*>(nodeA + at), __float_as_uint(L.x), __float_as_uint(L.y),
                      far ? Bq.x : __float_as_uint(L.z), __float_as_uint(L.w));
            sts_v4_if(staged, nodeB + at, Bq.x, Bq.y, pm0, pm1);
            if (cls == 5) cellv[at] = child;
        }
        __syncwarp();
        // ---- apply the staged nodes to the warp's 64 bodies ----
        int i = 0;
#pragma unroll 4
        for (; i < n1; ++i) apply_far(nodeA[i]);
#pragma unroll kFarMaskedUnroll
        for (; i < n1 + n2; ++i) {
            const uint4 SB = nodeB[i];
            apply_far_masked(nodeA[i], make_uint2(SB.z, SB.w));
        }
#pragma unroll kNearUnroll
        for (; i < n1 + n2 + n3; ++i) {
            const uint4 SB = nodeB[i];
            apply_near(nodeA[i], __uint_as_float(SB.x), SB.z, SB.w);
        }
        // class M: the per-body test, the only class that pushes here; two nodes per iteration (independent chains)
        const int endM = n1 + n2 + n3 + n5;
        for (; i + 1 < endM; i += 2) {
            const float4 SA0 = nodeA[i], SA1 = nodeA[i + 1];
            const uint4 SB0 = nodeB[i], SB1 = nodeB[i + 1];
            const uint32_t c0 = cellv[i], c1 = cellv[i + 1];
            uint32_t p0, p1, q0, q1;
            eval_mixed(SA0, __uint_as_float(SB0.x), __uint_as_float(SB0.y), c0, SB0.z, SB0.w, p0, p1);
            eval_mixed(SA1, __uint_as_float(SB1.x), __uint_as_float(SB1.y), c1, SB1.z, SB1.w, q0, q1);
            const bool pushp = (p0 | p1) != 0u, pushq = (q0 | q1) != 0u;
            sts_v4_if(pushp && lane == 0, stack + top, c0, p0, p1, 0u);
            top += pushp;
            sts_v4_if(pushq && lane == 0, stack + top, c1, q0, q1, 0u);
            top += pushq;
        }
        if (i < endM) {
            const float4 SA = nodeA[i];
            const uint4 SB = nodeB[i];
            const uint32_t c = cellv[i];
            uint32_t m0, m1;
            eval_mixed(SA, __uint_as_float(SB.x), __uint_as_float(SB.y), c, SB.z, SB.w, m0, m1);
            const bool push = (m0 | m1) != 0u;
            sts_v4_if(push && lane == 0, stack + top, c, m0, m1, 0u);
            top += push;
        }
        __syncwarp();
        if constexpr (LEAVES) {
            for (int e6 = 0; e6 < n6; ++e6) {
                const uint4 Lf = s_leaf[warp][e6];
                apply_leaf_members(Lf.x, Lf.y, Lf.z, Lf.w);
            }
        }
    }
    const float ax[2] = {accx.x, accx.y}, ay[2] = {accy.x, accy.y};
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        if (body[b] != 0xffffffffu) {
            const double2 p = a.pos_in[body[b]];
            const double mi = a.mass[body[b]];
            double fx = (double)ax[b], fy = (double)ay[b];
            if (spos[b]) fp64_body_walk(a, selfn[b], p.x, p.y, fx, fy);       // rescued body: the reference's FP64 walk
            finish_body<INTEGRATE>(a, body[b], p.x, p.y, mi, mi * fx, mi * fy);
        }
    }
    __syncwarp();       // the warp's shared arrays are reused by its next tile
  }
}

// ------------------------------------------------------------------------------------------------
// FP64 verification traversal: the reference's expressions, one body per lane
// ------------------------------------------------------------------------------------------------
template <bool INTEGRATE, bool COUNT, bool EXACT = false>
__global__ void __launch_bounds__(kTravThreads)
traverse_f64_kernel(const __grid_constant__ TravArgs a) {
    __shared__ uint2 s_stack[kTravWarps][kStackCap];
    pdl_entry();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t slot = ((int64_t)blockIdx.x * kTravWarps + warp) * 32 + lane;
    const bool live = slot < a.n_slots;
    uint32_t body = 0, selfn = 0xffffffffu;
    double px = 0.0, py = 0.0, mi = 0.0;
    if (live) {
        uint32_t sp = a.own_list ? a.own_list[slot] : (uint32_t)slot;
        body = a.sidx[sp];
        selfn = a.self_node[body];
        double2 p = a.pos_in[body];
        px = p.x; py = p.y;
        mi = a.mass[body];
    }
    double sx = 0.0, sy = 0.0;       // the reference's `sum`
    uint32_t c_int = 0, c_vis = 0, c_open = 0, c_steps = 0;

    auto eval = [&](uint32_t idx, bool active) -> bool {
        const uint32_t fl = a.flags[idx];
        const bool nz = fl & kNodeNonZero, leaf = fl & kNodeLeaf;
        if constexpr (EXACT) {   // multi-body leaf at the depth cap (warp-uniform test): its bodies one by one
            if (nz && idx >= a.finest_off && a.t_count[idx] >= 2u) {
                const uint32_t cnt = a.t_count[idx], first = a.t_first[idx];
                for (uint32_t j = 0; j < cnt; ++j) {
                    const uint32_t bj = a.sidx[first + j];
                    const double2 pj = a.pos_in[bj];
                    const double ex = pj.x - px, ey = pj.y - py;
                    const double e2 = ex * ex + ey * ey;
                    const double e = sqrt(e2) + a.dist_eps;
                    const bool usej = active && (bj != body);
                    if (usej) {
                        const double fm = (a.G * mi * a.mass[bj]) / e2;
                        sx += fm * (ex / e);
                        sy += fm * (ey / e);
                    }
                    if constexpr (COUNT) c_int += usej;
                }
                if constexpr (COUNT) c_vis += active;
                return false;
            }
        }
        const int level = (31 - __clz(3u * idx + 1u)) >> 1;                  // 3 off[l] + 1 == 4^l
        const double M = a.t_mass[idx];
        const double dx = a.t_comx[idx] - px, dy = a.t_comy[idx] - py;
        const double d2 = dx * dx + dy * dy;
        const double d = sqrt(d2) + a.dist_eps;                               // project.cu:748
        const bool accept = leaf || (a.consts->size[level] / d < a.theta);    // project.cu:757
        const bool use = active && nz && accept && (selfn != idx);            // project.cu:731, :760
        if (use) {
            const double fm = (a.G * mi * M) / d2;                            // project.cu:765
            sx += fm * (dx / d);                                              // project.cu:768-772
            sy += fm * (dy / d);
        }
        if constexpr (COUNT) {
            c_vis += active;
            c_int += use;
            c_open += (active && nz && !accept);
        }
        return active && nz && !accept;
    };

    int top = 0;
    {
        bool open = eval(0u, live);
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
        const uint32_t base = 4u * e.x + 1u;
        const bool active = (e.y >> lane) & 1u;
        if constexpr (COUNT) c_steps += (lane == 0);
#pragma unroll
        for (uint32_t q = 0; q < 4; ++q) {
            bool open = eval(base + q, active);
            uint32_t m = __ballot_sync(0xffffffffu, open);
            if (m) {
                if (lane == 0) s_stack[warp][top] = make_uint2(base + q, m);
                ++top;
            }
        }
        __syncwarp();
    }
    if (live) finish_body<INTEGRATE>(a, body, px, py, mi, sx, sy);
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

// ---- chunk lists for the pipelined host step (bh_step_host) ----------------------------------------
// The bodies are split into n_chunks contiguous ranges of ORIGINAL index (so that every range's results
// are one contiguous device-to-host copy); list k holds the sorted positions of the bodies of range k, in
// sorted (Morton) order — a stable partition of 0 .. n-1, built by count / scan / scatter so that a
// traversal warp still gets spatially neighbouring bodies.  List k starts at lists[lo[k]] (range k has
// exactly lo[k+1] - lo[k] bodies).
__device__ __forceinline__ int chunk_of(uint32_t body, const ChunkBounds& cb) {
    int c = 0;
#pragma unroll
    for (int k = 1; k < kMaxHostChunks; ++k) c += (k < cb.n_chunks && body >= cb.lo[k]);
    return c;
}

__global__ void __launch_bounds__(256)
chunk_count_kernel(const uint32_t* __restrict__ sidx, int64_t n, const __grid_constant__ ChunkBounds cb,
                   uint32_t* __restrict__ counts, int nblocks) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = (j < n) ? chunk_of(sidx[j], cb) : -1;
    for (int k = 0; k < cb.n_chunks; ++k) {
        const int cnt = __syncthreads_count(c == k);
        if (threadIdx.x == 0) counts[(size_t)k * nblocks + blockIdx.x] = (uint32_t)cnt;
    }
}

// exclusive scan of every row of counts[n_chunks][nblocks]; one block per row
__global__ void __launch_bounds__(1024)
chunk_scan_kernel(uint32_t* __restrict__ counts, int nblocks) {
    uint32_t* row = counts + (size_t)blockIdx.x * nblocks;
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + tid;
        const uint32_t v = (i < nblocks) ? row[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const uint32_t w = s_warp[lane];
            uint32_t wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += u;
            }
            s_warp[lane] = wi - w;
        }
        __syncthreads();
        const uint32_t excl = s_carry + s_warp[warp] + inc - v;
        if (i < nblocks) row[i] = excl;
        __syncthreads();
        if (tid == 1023) s_carry = excl + v;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
chunk_scatter_kernel(const uint32_t* __restrict__ sidx, int64_t n, const __grid_constant__ ChunkBounds cb,
                     const uint32_t* __restrict__ offsets, int nblocks, uint32_t* __restrict__ lists) {
    __shared__ uint32_t s_w[kMaxHostChunks][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = (j < n) ? chunk_of(sidx[j], cb) : -1;
    uint32_t mine = 0;
    for (int k = 0; k < cb.n_chunks; ++k) {
        const uint32_t m = __ballot_sync(0xffffffffu, c == k);
        if (lane == 0) s_w[k][warp] = __popc(m);
        if (c == k) mine = m;
    }
    __syncthreads();
    if (c >= 0) {
        uint32_t pre = 0;
        for (int w = 0; w < warp; ++w) pre += s_w[c][w];
        lists[cb.lo[c] + offsets[(size_t)c * nblocks + blockIdx.x] + pre + __popc(mine & ((1u << lane) - 1u))] = (uint32_t)j;
    }
}

}  // namespace

void launch_traverse(const uint32_t* skeys, const uint32_t* sidx, const double2* pos_in, const double2* vel_in,
                     double2* pos, double2* vel, double2* acc,
                     double2* force, const double* mass, int64_t n, int64_t own_lo, int64_t own_hi,
                     const uint32_t* own_list, const uint32_t* own_count_dev, int64_t own_n,
                     const bh_params& p, const Dims& d, const TreeArrays& t, const StepConsts* consts,
                     unsigned long long* counters, bool integrate, cudaStream_t st) {
    (void)own_lo; (void)own_hi; (void)own_count_dev; (void)n;
    TravArgs a;
    a.skeys = skeys; a.sidx = sidx; a.own_list = own_list; a.self_node = t.self_node;
    a.pos_in = pos_in; a.vel_in = vel_in;
    a.pos = pos; a.vel = vel; a.acc = acc; a.force = force; a.mass = mass;
    a.rec = t.rec; a.flags = t.flags; a.t_mass = t.mass; a.t_comx = t.comx; a.t_comy = t.comy;
    a.consts = consts; a.counters = counters;
    a.n_slots = own_n;
    a.G = p.G; a.dt = p.dt; a.theta = p.theta; a.dist_eps = p.dist_eps;
    a.t_count = t.count; a.t_first = t.first; a.finest_off = (uint32_t)d.level_off[d.finest];
    a.tile_queue = t.tile_queue;
    if (own_n <= 0) return;
    const bool fp64 = p.flags & BH_FLAG_FP64_TRAVERSAL, count = p.flags & BH_FLAG_COUNTERS;
    // two bodies per lane halve the node traffic and the control overhead per body, but need >= ~400k
    // bodies to keep every SM's warp slots full (ncu: profiles/r01_traverse_v4_*)
    const bool exact_leaves = p.flags & BH_FLAG_EXACT_LEAVES;   // extension: generic 1-body-per-lane / FP64 kernels only
    const bool leaves_pair = exact_leaves && p.reserved[0] == 2 && !(p.flags & (BH_FLAG_COUNTERS | BH_FLAG_FP64_TRAVERSAL));
    // exact leaves in the list kernel: on request (9) and by default from kTwoBodiesPerLaneMin bodies on (launches over an
    // own-list — multi-rank exact leaves, host step — arrive here with reserved[0] = 1 / 2 and keep the generic / pair kernel)
    const bool leaves_list = exact_leaves && !(p.flags & (BH_FLAG_COUNTERS | BH_FLAG_FP64_TRAVERSAL)) &&
                             (p.reserved[0] == 9 || (p.reserved[0] == 0 && own_n >= kTwoBodiesPerLaneMin));
    const int bpl = leaves_pair ? 2 : exact_leaves ? 1 : (p.reserved[0] == 1) ? 1 : (p.reserved[0] == 2 || p.reserved[0] == 3 || p.reserved[0] == 8) ? 2 : (own_n >= kTwoBodiesPerLaneMin ? 2 : 1);
    // previous operation on the stream = tree_top_kernel (or a peer-exchange kernel; g_pdl is off there)
#define BH_GO(K) launch_chain(K, dim3(blocks), dim3(kTravThreads), st, true, a)
    if (fp64) {
        unsigned blocks = (unsigned)((own_n + kTravThreads - 1) / kTravThreads);
        if (exact_leaves) {
            if (integrate) { if (count) BH_GO((traverse_f64_kernel<true, true, true>)); else BH_GO((traverse_f64_kernel<true, false, true>)); }
            else { if (count) BH_GO((traverse_f64_kernel<false, true, true>)); else BH_GO((traverse_f64_kernel<false, false, true>)); }
        } else
        if (integrate) { if (count) BH_GO((traverse_f64_kernel<true, true>)); else BH_GO((traverse_f64_kernel<true, false>)); }
        else { if (count) BH_GO((traverse_f64_kernel<false, true>)); else BH_GO((traverse_f64_kernel<false, false>)); }
    } else {
        const int64_t per_block = (int64_t)kTravThreads * bpl;
        unsigned blocks = (unsigned)((own_n + per_block - 1) / per_block);
#define BH_TRAV(B, I, C) BH_GO((traverse_f32_kernel<B, I, C>))
        if (leaves_list) {      // exact leaves in the list kernel (persistent warps over 64-body tiles)
            const bool exact = p.flags & BH_FLAG_EXACT_EPS;
            static int resident_l = 0;
            if (!resident_l) {
                int dev = 0, sms = 148, per_sm = BH_LIST_MIN_BLOCKS;
                cudaGetDevice(&dev);
                cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, traverse_f32_list_kernel<true, false, true>, kTravThreads, 0);
                resident_l = sms * (per_sm > 0 ? per_sm : BH_LIST_MIN_BLOCKS);
            }
            const unsigned tiles = (unsigned)((own_n + 63) / 64);
            blocks = std::min<unsigned>((tiles + kTravWarps - 1) / kTravWarps, (unsigned)resident_l);
            if (integrate) { if (exact) BH_GO((traverse_f32_list_kernel<true, true, true>)); else BH_GO((traverse_f32_list_kernel<true, false, true>)); }
            else { if (exact) BH_GO((traverse_f32_list_kernel<false, true, true>)); else BH_GO((traverse_f32_list_kernel<false, false, true>)); }
        } else if (leaves_pair) {
            const bool exact = p.flags & BH_FLAG_EXACT_EPS;
            if (integrate) { if (exact) BH_GO((traverse_f32_pair_kernel<true, true, true>)); else BH_GO((traverse_f32_pair_kernel<true, false, true>)); }
            else { if (exact) BH_GO((traverse_f32_pair_kernel<false, true, true>)); else BH_GO((traverse_f32_pair_kernel<false, false, true>)); }
        } else if (exact_leaves) {
            if (integrate) { if (count) BH_GO((traverse_f32_kernel<1, true, true, true>)); else BH_GO((traverse_f32_kernel<1, true, false, true>)); }
            else { if (count) BH_GO((traverse_f32_kernel<1, false, true, true>)); else BH_GO((traverse_f32_kernel<1, false, false, true>)); }
        } else if (bpl == 2 && !count && (p.reserved[0] == 0 || p.reserved[0] == 8)) {
            const bool exact = p.flags & BH_FLAG_EXACT_EPS;
            // persistent warps: as many blocks as can be resident (or fewer when there are not that many tiles)
            static int resident = 0;
            if (!resident) {
                int dev = 0, sms = 148, per_sm = BH_LIST_MIN_BLOCKS;
                cudaGetDevice(&dev);
                cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, traverse_f32_list_kernel<true, false>, kTravThreads, 0);
                resident = sms * (per_sm > 0 ? per_sm : BH_LIST_MIN_BLOCKS);
            }
            const unsigned tiles = (unsigned)((own_n + 63) / 64);
            blocks = std::min<unsigned>((tiles + kTravWarps - 1) / kTravWarps, (unsigned)resident);
            if (integrate) { if (exact) BH_GO((traverse_f32_list_kernel<true, true>)); else BH_GO((traverse_f32_list_kernel<true, false>)); }
            else { if (exact) BH_GO((traverse_f32_list_kernel<false, true>)); else BH_GO((traverse_f32_list_kernel<false, false>)); }
        } else if (bpl == 2 && !count && p.reserved[0] != 3) {
            const bool exact = p.flags & BH_FLAG_EXACT_EPS;
            if (integrate) { if (exact) BH_GO((traverse_f32_pair_kernel<true, true>)); else BH_GO((traverse_f32_pair_kernel<true, false>)); }
            else { if (exact) BH_GO((traverse_f32_pair_kernel<false, true>)); else BH_GO((traverse_f32_pair_kernel<false, false>)); }
        } else if (bpl == 2) {
            if (integrate) { if (count) BH_TRAV(2, true, true); else BH_TRAV(2, true, false); }
            else { if (count) BH_TRAV(2, false, true); else BH_TRAV(2, false, false); }
        } else {
            if (integrate) { if (count) BH_TRAV(1, true, true); else BH_TRAV(1, true, false); }
            else { if (count) BH_TRAV(1, false, true); else BH_TRAV(1, false, false); }
        }
#undef BH_TRAV
    }
#undef BH_GO
    ++g_launches;
}

void launch_integrate(double2* pos, double2* vel, double2* acc, const double2* force, const double* mass,
                      int64_t lo, int64_t hi, double dt, cudaStream_t st) {
    if (hi <= lo) return;
    integrate_kernel<<<(unsigned)((hi - lo + 255) / 256), 256, 0, st>>>(pos, vel, acc, force, mass, lo, hi, dt);
    ++g_launches;
}

void launch_chunk_lists(const uint32_t* sidx, int64_t n, const ChunkBounds& cb, uint32_t* counts, uint32_t* lists,
                        cudaStream_t st) {
    const int nblocks = (int)((n + 255) / 256);
    chunk_count_kernel<<<nblocks, 256, 0, st>>>(sidx, n, cb, counts, nblocks);
    chunk_scan_kernel<<<cb.n_chunks, 1024, 0, st>>>(counts, nblocks);
    chunk_scatter_kernel<<<nblocks, 256, 0, st>>>(sidx, n, cb, counts, nblocks, lists);
    g_launches += 3;
}

}  // namespace bh
