// Kernel 2: onesweep LSD radix sort of (cell key, body index) pairs.
//
// No reference counterpart: the reference inserts bodies one at a time on the host
// (project.cu:586-588).  Sorting by cell key makes (a) every finest cell's bodies one contiguous,
// index-ordered run (stable sort => the reference's insertion order inside a cell is kept, which
// is what makes the cap-level running average of project.cu:367-373 reproducible bit for bit) and
// (b) consecutive bodies spatially coherent for the warp-level traversal.
//
// One kernel per digit pass ("onesweep"): every tile of 4096 keys ranks its keys locally, publishes
// its per-digit counts and obtains the counts of all earlier tiles by decoupled look-back over a
// per-pass status array, so each key is read once and written once per pass.  The per-pass global
// digit histograms are produced by the key-generation kernel (bounds_keys.cu), so there is no
// separate histogram read of the keys.  Tile ids come from an atomic ticket, which guarantees that
// every tile a block waits on is already resident (forward progress without co-scheduling
// assumptions).  Stable: ranks follow (warp, item, lane) order == global index order.
#include "bh_internal.h"

namespace bh {

namespace {

constexpr uint32_t kFlagAgg = 1u << 30;   // tile published its own digit count
constexpr uint32_t kFlagIncl = 2u << 30;  // tile published the inclusive prefix over tiles 0..t
constexpr uint32_t kValMask = (1u << 30) - 1u;

// FIRST: the values of the first pass are the body indices in input order (val[g] = val_base + g): they are
// generated instead of being read, and the key kernel does not have to write them.
#ifndef BH_SORT_MIN_BLOCKS
#define BH_SORT_MIN_BLOCKS 3     // <= 80 registers: 24 warps per SM instead of 16 (the pass is latency-bound)
#endif
template <int NBINS_LOG2, bool FIRST, int ITEMS>
__global__ void __launch_bounds__(kSortThreads, BH_SORT_MIN_BLOCKS)
onesweep_pass(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
              uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n, int shift,
              int bits, const uint32_t* __restrict__ digit_hist, uint32_t* tile_state, uint32_t* ticket,
              uint32_t val_base) {
    constexpr int NBINS = 1 << NBINS_LOG2;
    constexpr int NWARPS = kSortThreads / 32;
    constexpr int PER_T = NBINS / kSortThreads;  // digits owned per thread (1 or 2)
    __shared__ uint32_t s_warp[NWARPS][NBINS];   // per-warp digit counts -> per-warp offsets
    __shared__ uint32_t s_base[NBINS];           // global position of this tile's first key of digit d
    __shared__ uint32_t s_scan[NWARPS];
    __shared__ uint32_t s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_entry();
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < NWARPS * NBINS; i += kSortThreads) (&s_warp[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t dmask = (1u << bits) - 1u;
    constexpr int kSortItems = ITEMS;     // (shadows the namespace constant: this instantiation's tile shape)
    const int64_t base = (int64_t)tile * (kSortThreads * ITEMS) + (int64_t)warp * (32 * ITEMS) + lane;

    uint32_t key[kSortItems];
    uint32_t rank[kSortItems];
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        int64_t g = base + (int64_t)k * 32;
        key[k] = (g < n) ? keys_in[g] : 0xffffffffu;
    }
    // warp-level stable ranking.  First all peer masks (independent MATCH instructions, pipelined),
    // then the running per-warp digit counts, item by item (the only sequential part).
    uint32_t peers[kSortItems];
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const bool valid = base + (int64_t)k * 32 < n;
        const uint32_t d = (key[k] >> shift) & dmask;
        peers[k] = __match_any_sync(0xffffffffu, valid ? d : (0x10000u | (uint32_t)lane));  // invalid lanes match nobody
    }
    const uint32_t lanemask_lt = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const bool valid = base + (int64_t)k * 32 < n;
        const uint32_t d = (key[k] >> shift) & dmask;
        const uint32_t lt = peers[k] & lanemask_lt;
        const uint32_t running = valid ? s_warp[warp][d] : 0u;
        rank[k] = running + __popc(lt);
        __syncwarp();
        if (valid && lt == 0) s_warp[warp][d] = running + __popc(peers[k]);
        __syncwarp();
    }
    __syncthreads();

    // per digit: exclusive scan over warps, tile total, look-back over earlier tiles
    uint32_t tile_count[PER_T], excl_prev[PER_T];
#pragma unroll
    for (int j = 0; j < PER_T; ++j) {
        int d = tid + j * kSortThreads;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < NWARPS; ++w) {
            uint32_t c = s_warp[w][d];
            s_warp[w][d] = run;
            run += c;
        }
        tile_count[j] = run;
        volatile uint32_t* st = tile_state + (size_t)tile * NBINS + d;
        if (tile == 0) *st = kFlagIncl | run;
        else *st = kFlagAgg | run;
    }
    // Decoupled look-back, kWindow predecessor states in flight per digit (the loads are
    // independent, so the serial chain is one L2 round trip per kWindow tiles, not per tile).
    {
        constexpr int kWindow = 8;
        int64_t t[PER_T];
        bool done[PER_T];
#pragma unroll
        for (int j = 0; j < PER_T; ++j) { excl_prev[j] = 0; t[j] = (int64_t)tile - 1; done[j] = (tile == 0); }
        bool all_done = (tile == 0);
        while (!all_done) {
            all_done = true;
#pragma unroll
            for (int j = 0; j < PER_T; ++j) {
                if (done[j]) continue;
                const int d = tid + j * kSortThreads;
                uint32_t v[kWindow];
#pragma unroll
                for (int w = 0; w < kWindow; ++w) {
                    const int64_t tt = t[j] - w;
                    volatile uint32_t* st = tile_state + (size_t)(tt < 0 ? 0 : tt) * NBINS + d;
                    v[w] = *st;
                    if (tt < 0) v[w] = 2u << 30;   // tile 0 always publishes an inclusive prefix
                }
#pragma unroll
                for (int w = 0; w < kWindow; ++w) {
                    if (done[j]) break;
                    if (v[w] == 0) break;                // not published yet: retry from here
                    excl_prev[j] += v[w] & kValMask;
                    --t[j];
                    if ((v[w] >> 30) == 2u) done[j] = true;
                }
                all_done = all_done && done[j];
            }
        }
        if (tile > 0) {
#pragma unroll
            for (int j = 0; j < PER_T; ++j) {
                volatile uint32_t* mine = tile_state + (size_t)tile * NBINS + (tid + j * kSortThreads);
                *mine = kFlagIncl | (excl_prev[j] + tile_count[j]);
            }
        }
    }
    // exclusive scan of the global digit histogram (block-wide, NBINS values)
    {
        uint32_t h[PER_T], local = 0;
#pragma unroll
        for (int j = 0; j < PER_T; ++j) {
            // thread owns digits tid*PER_T + j here (contiguous) for the scan
            h[j] = digit_hist[tid * PER_T + j];
            local += h[j];
        }
        uint32_t inc = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) s_scan[warp] = inc;
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < NWARPS; ++w) woff += (w < warp) ? s_scan[w] : 0u;
        uint32_t ex = woff + inc - local;
#pragma unroll
        for (int j = 0; j < PER_T; ++j) {
            s_base[tid * PER_T + j] = ex;   // global exclusive start of digit
            ex += h[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < PER_T; ++j) {
        int d = tid + j * kSortThreads;
        s_base[d] += excl_prev[j];
    }
    __syncthreads();

    // scatter keys and values
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        int64_t g = base + (int64_t)k * 32;
        if (g < n) {
            uint32_t d = (key[k] >> shift) & dmask;
            uint32_t dst = s_base[d] + s_warp[warp][d] + rank[k];
            keys_out[dst] = key[k];
            if constexpr (FIRST) vals_out[dst] = val_base + (uint32_t)g;
            else vals_out[dst] = vals_in[g];
        }
    }
}

}  // namespace

void launch_sort(uint32_t* keys[2], uint32_t* vals[2], int64_t n, const SortPlan& sp, Scratch& s,
                 int* result_buf, cudaStream_t st, uint32_t val_base) {
    int cur = 0;
    const int nbins = 1 << sp.nbins_log2;
    for (int pass = 0; pass < sp.passes; ++pass) {
        uint32_t* state = s.tile_state + (size_t)pass * sp.ntiles * nbins;
        const uint32_t* hist = s.digit_hist + (size_t)pass * kMaxBins;
        int shift = pass * sp.bits_per_pass;
#define BH_PASS(NB, F, IT)                                                                                               \
        launch_chain(onesweep_pass<NB, F, IT>, dim3(sp.ntiles), dim3(kSortThreads), st, true, (const uint32_t*)keys[cur], \
                     (const uint32_t*)vals[cur], keys[cur ^ 1], vals[cur ^ 1], n, shift, sp.bits_per_pass, hist, state,  \
                     s.tickets + pass, val_base)
#define BH_PASS_IT(NB, F) do { if (sp.items == kSortItemsSmall) BH_PASS(NB, F, kSortItemsSmall); else BH_PASS(NB, F, kSortItems); } while (0)
        if (sp.nbins_log2 == 9) { if (pass == 0) BH_PASS_IT(9, true); else BH_PASS_IT(9, false); }
        else { if (pass == 0) BH_PASS_IT(8, true); else BH_PASS_IT(8, false); }
#undef BH_PASS_IT
#undef BH_PASS
        ++g_launches;
        cur ^= 1;
    }
    *result_buf = cur;
}

}  // namespace bh
