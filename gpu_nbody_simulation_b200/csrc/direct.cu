// Direct all-pairs O(N^2) force kernel (BASELINE.json config 5) and the FP32 FMA peak probe.
//
// Formula of the reference's direct-sum prototype, main_approach_1.cpp:53-75:
//     F_i = sum_{j != i} G m_i m_j (r_j - r_i) / (d^2 * d)          (no softening)
// FP32 arithmetic on coordinates re-centred on the bounding-box centre (FP64 subtraction before
// the conversion), classic shared-memory tiling: every thread owns one body i, a block streams
// tiles of 256 bodies j through shared memory.  Pairs with d == 0 (i == j, or exactly coincident
// bodies, for which the reference produces NaN) contribute nothing.
#include "bh_internal.h"

namespace bh {

namespace {

constexpr int kDirectThreads = 256;

__global__ void __launch_bounds__(256)
pack_kernel(const double2* __restrict__ pos, const double* __restrict__ mass, int64_t n, double cx, double cy,
            float4* __restrict__ packed) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double2 p = pos[i];
    packed[i] = make_float4((float)(p.x - cx), (float)(p.y - cy), (float)mass[i], 0.f);
}

__global__ void __launch_bounds__(kDirectThreads)
direct_kernel(const float4* __restrict__ packed, int64_t n, double G, const double* __restrict__ mass,
              double2* __restrict__ force) {
    __shared__ float4 tile[kDirectThreads];
    const int64_t i = (int64_t)blockIdx.x * kDirectThreads + threadIdx.x;
    float4 me = (i < n) ? packed[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float ax = 0.f, ay = 0.f;
    for (int64_t j0 = 0; j0 < n; j0 += kDirectThreads) {
        int64_t j = j0 + threadIdx.x;
        tile[threadIdx.x] = (j < n) ? packed[j] : make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
#pragma unroll 16
        for (int k = 0; k < kDirectThreads; ++k) {
            float4 o = tile[k];
            float dx = o.x - me.x, dy = o.y - me.y;
            float d2 = fmaf(dx, dx, dy * dy);
            float inv = rsqrtf(d2);
            float w = o.z * inv * inv * inv;          // m_j / (d^2 d)
            w = (d2 > 0.f) ? w : 0.f;                 // i == j (and padded entries: mass 0)
            ax = fmaf(w, dx, ax);
            ay = fmaf(w, dy, ay);
        }
        __syncthreads();
    }
    if (i < n) {
        double gm = G * mass[i];
        force[i] = make_double2(gm * (double)ax, gm * (double)ay);
    }
}

// 8 independent FMA chains per thread, 4096 x 8 FMAs each.
__global__ void __launch_bounds__(256)
fma_peak_kernel(float* out, float a, float b, int iters) {
    float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
          x7 = x0 + 7.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 12345.678f) out[0] = s;   // never true; keeps the chains alive
}

}  // namespace

void launch_direct(const double2* pos, const double* mass, int64_t n, double G, float4* packed,
                   double2* force, cudaStream_t st) {
    // centre: mean of the first / last body is good enough to halve the magnitude; use 0 when unknown.
    pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pos, mass, n, 0.0, 0.0, packed);
    ++g_launches;
    direct_kernel<<<(unsigned)((n + kDirectThreads - 1) / kDirectThreads), kDirectThreads, 0, st>>>(packed, n, G, mass,
                                                                                                 force);
    ++g_launches;
}

int measure_fp32_peak(int device, double* tflops, double* mhz) {
    int prev = 0;
    BH_CUDA_OK(cudaGetDevice(&prev));
    if (device >= 0) BH_CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    int dev = 0;
    BH_CUDA_OK(cudaGetDevice(&dev));
    BH_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
    float* out = nullptr;
    BH_CUDA_OK(cudaMalloc(&out, 4));
    cudaEvent_t e0, e1;
    BH_CUDA_OK(cudaEventCreate(&e0));
    BH_CUDA_OK(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8, iters = 2048;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        BH_CUDA_OK(cudaEventRecord(e0));
        fma_peak_kernel<<<blocks, 256>>>(out, 1.0000001f, 1e-9f, iters);
        BH_CUDA_OK(cudaEventRecord(e1));
        BH_CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0.f;
        BH_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        double fl = 2.0 * 8.0 * 16.0 * iters * 256.0 * blocks;
        double tf = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    BH_CUDA_OK(cudaGetLastError());
    *tflops = best;
    if (mhz) *mhz = best * 1e12 / (2.0 * 128.0 * prop.multiProcessorCount) / 1e6;  // clock implied at 128 FMA/clk/SM
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    if (device >= 0) cudaSetDevice(prev);
    return BH_OK;
}

}  // namespace bh
