// Micro-benchmark: issue / pipe throughput of scalar FP32 ops vs packed FP32x2 ops on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe tools/pipe_probe.cu && ./pipe_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, float a, float b, int iters) {
    float2 x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    float2 av = make_float2(a, a + 1e-7f), bv = make_float2(b, b + 1e-9f);
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = threadIdx.x * 0.5f + i;
    float ar = a * 1.0000001f, br = b + 3e-9f;   // register operands (not constant bank)
    asm volatile("" : "+f"(ar), "+f"(br));
    asm volatile("" : "+f"(av.x), "+f"(av.y), "+f"(bv.x), "+f"(bv.y));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) s[i] = fmaf(s[i], ar, br);                        // 8 scalar FFMA (3 regs)
                if (MODE == 1) x[i] = __ffma2_rn(x[i], av, bv);                  // 8 packed FFMA2
                if (MODE == 2) { if (i < 4) x[i] = __ffma2_rn(x[i], av, bv); else s[i] = fmaf(s[i], ar, br); }  // 4 + 4
                if (MODE == 3) s[i] = s[i] + ar;                                 // 8 scalar FADD
                if (MODE == 4) x[i] = __fadd2_rn(x[i], av);                      // 8 packed FADD2
                if (MODE == 5) { if (i < 4) x[i] = __fadd2_rn(x[i], av); else s[i] = s[i] + ar; }
                if (MODE == 6) { if (i < 2) x[i] = __ffma2_rn(x[i], av, bv); else s[i] = fmaf(s[i], ar, br); }  // 2 + 6
            }
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += x[i].x + x[i].y + s[i];
    if (r == 1234.5f) out[0] = r;
}

template <int MODE>
void run(const char* name, double lane_ops_per_inner) {
    float* out; cudaMalloc(&out, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int dev; cudaGetDevice(&dev); cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    const int blocks = p.multiProcessorCount * 8, iters = 4096;
    probe<MODE><<<blocks, 256>>>(out, 1.0000001f, 1e-9f, 16);
    float best = 1e9f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        probe<MODE><<<blocks, 256>>>(out, 1.0000001f, 1e-9f, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double warp_instr = (double)blocks * 8 * iters * 4 * 8;            // warp-instructions issued
    double per_smsp_cycle = warp_instr / (p.multiProcessorCount * 4.0) / (best * 1e-3 * 1.9e9);
    double flops = (double)blocks * 256 * iters * 4 * lane_ops_per_inner;
    printf("%-34s %8.3f ms  %6.2f warp-instr/clk/SMSP (at 1.9 GHz)  %7.1f Glane-op/s\n", name, best, per_smsp_cycle,
           flops / (best * 1e-3) / 1e9);
    cudaFree(out);
}

int main() {
    run<0>("8x FFMA  (scalar, 3-reg)", 8);
    run<1>("8x FFMA2 (packed)", 16);
    run<2>("4x FFMA2 + 4x FFMA", 12);
    run<6>("2x FFMA2 + 6x FFMA", 10);
    run<3>("8x FADD  (scalar)", 8);
    run<4>("8x FADD2 (packed)", 16);
    run<5>("4x FADD2 + 4x FADD", 12);
    return 0;
}
