// Seeded initial conditions, on the device and on the host.
//
// Replaces the reference's time-seeded random initialisation (initializeGpu / initializeCpu,
// project.cu:298-341: cuRAND states seeded with time(0), project.cu:219-228; value ranges
// project.cu:30-35) with a counter-based generator: body i's values are a pure function of
// (seed, i), so any rank can generate exactly its own slice straight into HBM, the host can produce
// the same bodies without a GPU, and a run is reproducible.  Distributions: the reference's uniform
// square, and BASELINE.json's uniform disk (config 2/4) and projected Plummer sphere (config 3).
//
// Generator: Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11), counter = (body lo, body hi, stream,
// draw), key = (seed lo, seed hi); checked against the Random123 known-answer vectors in
// tests/test_generate.py.  One call yields two doubles in [0, 1) with 53 random bits each.
// The per-body code is one __host__ __device__ function; host and device differ only in the last bits
// of libm / CUDA transcendentals (sqrt, sin, cos, pow), additions and multiplications are not contracted.
#include <math.h>

#include "bh_internal.h"

namespace bh {

namespace {

constexpr double kLowerM = 1e-1, kHigherM = 5e-1;     // project.cu:30-31
constexpr double kLowerP = -1e-1, kHigherP = 1e-1;    // project.cu:32-33
constexpr double kLowerV = -1e-4, kHigherV = 1e-4;    // project.cu:34-35
constexpr double kDiskRadius = 0.1, kPlummerA = 0.02, kPlummerRmax = 0.1;
constexpr double kTwoPi = 6.283185307179586476925286766559;

__host__ __device__ inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        if (r) { k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[1] = (uint32_t)p1; c[3] = (uint32_t)p0; c[0] = n0; c[2] = n2;
    }
}

// two uniforms in [0, 1) for (body, stream, draw)
__host__ __device__ inline void uniform2(uint64_t seed, uint64_t body, uint32_t stream, uint32_t draw, double& u0,
                                         double& u1) {
    uint32_t c[4] = {(uint32_t)body, (uint32_t)(body >> 32), stream, draw};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t a = ((uint64_t)c[1] << 32) | c[0], b = ((uint64_t)c[3] << 32) | c[2];
    u0 = (double)(a >> 11) * (1.0 / 9007199254740992.0);
    u1 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
}

__host__ __device__ inline double affine(double lo, double u, double span) {   // lo + u * span, not contracted
#ifdef __CUDA_ARCH__
    return __dadd_rn(lo, __dmul_rn(u, span));
#else
    return lo + u * span;   // x86-64 host code is built without FMA instructions
#endif
}
__host__ __device__ inline double mul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}

// streams: 0 position (draw = rejection attempt), 1 velocity, 2 mass
__host__ __device__ inline void generate_body(int kind, uint64_t seed, uint64_t i, double* pos2, double* vel2,
                                              double* mass1) {
    double u0, u1;
    if (kind == BH_GEN_UNIFORM_SQUARE) {                 // project.cu:93-95 ranges
        uniform2(seed, i, 0u, 0u, u0, u1);
        pos2[0] = affine(kLowerP, u0, kHigherP - kLowerP);
        pos2[1] = affine(kLowerP, u1, kHigherP - kLowerP);
    } else if (kind == BH_GEN_UNIFORM_DISK) {            // r = R sqrt(u0), phi = 2 pi u1
        uniform2(seed, i, 0u, 0u, u0, u1);
        const double r = mul(kDiskRadius, sqrt(u0)), phi = mul(kTwoPi, u1);
        pos2[0] = mul(r, cos(phi));
        pos2[1] = mul(r, sin(phi));
    } else {                                             // Plummer sphere, z dropped, 3-D radius <= rmax
        double x = 0.0, y = 0.0;
        for (uint32_t attempt = 0; attempt < 4096u; ++attempt) {
            double cz_u, ph_u;
            uniform2(seed, i, 0u, 2u * attempt, u0, cz_u);
            uniform2(seed, i, 0u, 2u * attempt + 1u, ph_u, u1);
            if (!(u0 > 0.0)) continue;
            const double r = kPlummerA / sqrt(pow(u0, -2.0 / 3.0) - 1.0);   // r = a / sqrt(u^(-2/3) - 1)
            if (!(r <= kPlummerRmax)) continue;
            const double cz = affine(-1.0, cz_u, 2.0), s = sqrt(1.0 - mul(cz, cz)), ph = mul(kTwoPi, ph_u);
            x = mul(mul(r, s), cos(ph));
            y = mul(mul(r, s), sin(ph));
            break;
        }
        pos2[0] = x; pos2[1] = y;
    }
    uniform2(seed, i, 1u, 0u, u0, u1);
    vel2[0] = affine(kLowerV, u0, kHigherV - kLowerV);
    vel2[1] = affine(kLowerV, u1, kHigherV - kLowerV);
    uniform2(seed, i, 2u, 0u, u0, u1);
    // project.cu:86-89 / :99-101: log-uniform mass, 10 ^ (log10(lo) + u (log10(hi) - log10(lo)))
    *mass1 = pow(10.0, affine(log10(kLowerM), u0, log10(kHigherM) - log10(kLowerM)));
}

__global__ void __launch_bounds__(256)
generate_kernel(int kind, uint64_t seed, int64_t lo, int64_t hi, double2* __restrict__ pos, double2* __restrict__ vel,
                double* __restrict__ mass) {
    const int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    double p[2], v[2], m;
    generate_body(kind, seed, (uint64_t)i, p, v, &m);
    pos[i] = make_double2(p[0], p[1]);
    vel[i] = make_double2(v[0], v[1]);
    mass[i] = m;
}

}  // namespace

int generate_host(int kind, uint64_t seed, int64_t i0, int64_t i1, double* pos, double* vel, double* mass) {
    if (kind < BH_GEN_UNIFORM_SQUARE || kind > BH_GEN_PLUMMER_2D) { set_error("unknown generator kind %d", kind); return BH_ERR_INVALID; }
    for (int64_t i = i0; i < i1; ++i) generate_body(kind, seed, (uint64_t)i, pos + 2 * (i - i0), vel + 2 * (i - i0), mass + (i - i0));
    return BH_OK;
}

void philox_host(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {counter[0], counter[1], counter[2], counter[3]};
    philox4x32_10(c, key[0], key[1]);
    for (int i = 0; i < 4; ++i) out[i] = c[i];
}

void launch_generate(int kind, uint64_t seed, int64_t lo, int64_t hi, double2* pos, double2* vel, double* mass,
                     cudaStream_t st) {
    if (hi <= lo) return;
    generate_kernel<<<(unsigned)((hi - lo + 255) / 256), 256, 0, st>>>(kind, seed, lo, hi, pos, vel, mass);
    ++g_launches;
}

}  // namespace bh
