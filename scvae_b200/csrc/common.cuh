// Shared device/host helpers for libscvae_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/scvae_b200.h"

namespace scvae {

// ---- error plumbing ---------------------------------------------------------------------
void set_error(const char *fmt, ...);
void count_launch();

#define SCVAE_CHECK_ARG(cond, ...)                 \
    do {                                           \
        if (!(cond)) {                             \
            ::scvae::set_error(__VA_ARGS__);       \
            return 1;                              \
        }                                          \
    } while (0)

#define SCVAE_CHECK_LAUNCH(name)                                                   \
    do {                                                                           \
        ::scvae::count_launch();                                                   \
        cudaError_t e_ = cudaGetLastError();                                       \
        if (e_ != cudaSuccess) {                                                   \
            ::scvae::set_error("%s: launch failed: %s", name, cudaGetErrorString(e_)); \
            return 2;                                                              \
        }                                                                          \
    } while (0)

__host__ __device__ static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- constants --------------------------------------------------------------------------
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kHalfLog2Pi = 0.9189385332046727f;
// log(float32 tiny): sigmoid heads are clipped to [tiny, 1] (VAE:2481-2485), i.e. their
// logits to >= log(tiny) (sigmoid(a) < tiny  <=>  a < log(tiny) up to rounding).
constexpr float kLogitFloor = -87.33654475f;
constexpr float kBnEps = 1e-3f;   // tf.contrib.layers.batch_norm default (MU:62-70)
constexpr float kBnDecay = 0.999f;

// ---- streaming 128-bit global access ----------------------------------------------------
__device__ __forceinline__ float4 ldg_stream4(const float *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream4(float *p, const float4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// ---- reductions -------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Deterministic block sum (fixed tree); result valid in every thread. `red` >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float *red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : 0.f;
    r = warp_sum(r);
    return r;
}

// ---- scalar math ------------------------------------------------------------------------
// log1p for e in [0, 1]: exact-ish series below 2^-6, log(1+e) above.
__device__ __forceinline__ float log1p_unit(float e) {
    const float series = e * (1.f - e * (0.5f - e * (0.33333334f - 0.25f * e)));
    const float direct = __logf(1.f + e);
    return e < 0.015625f ? series : direct;
}
// softplus(a) = log(1 + exp(a)); also returns e = exp(-|a|) for the sigmoid.
__device__ __forceinline__ float softplus_e(float a, float &e) {
    e = __expf(-fabsf(a));
    return fmaxf(a, 0.f) + log1p_unit(e);
}
// sigmoid(a) from e = exp(-|a|)
__device__ __forceinline__ float sigmoid_from_e(float a, float e) {
    const float inv = __frcp_rn(1.f + e);
    return a >= 0.f ? inv : e * inv;
}

// Stirling series, z >= 8 (truncation error < 3e-8).
__device__ __forceinline__ float lgamma_big(float z) {
    const float iz = __frcp_rn(z);
    const float iz2 = iz * iz;
    return (z - 0.5f) * logf(z) - z + kHalfLog2Pi + iz * (0.083333336f - iz2 * 0.0027777778f);
}
__device__ __forceinline__ float digamma_big(float z) {
    const float iz = __frcp_rn(z);
    const float iz2 = iz * iz;
    return logf(z) - 0.5f * iz -
           iz2 * (0.083333336f - iz2 * (0.008333334f - iz2 * 0.003968254f));
}
// prod_{i<8}(z+i) and its logarithmic derivative sum_{i<8} 1/(z+i), z in (0, 8).
__device__ __forceinline__ void rising8(float z, float &prod, float &dlog) {
    float p = z, dp = 1.f;
#pragma unroll
    for (int i = 1; i < 8; ++i) {
        const float f = z + (float)i;
        dp = dp * f + p;
        p = p * f;
    }
    prod = p;
    dlog = __fdividef(dp, p);
}
__device__ __forceinline__ void lgamma_digamma_pos(float z, float &lg, float &dg) {
    if (z >= 8.f) {
        lg = lgamma_big(z);
        dg = digamma_big(z);
    } else {
        float p, dl;
        rising8(z, p, dl);
        lg = lgamma_big(z + 8.f) - logf(p);
        dg = digamma_big(z + 8.f) - dl;
    }
}
// D = lgamma(r + x) - lgamma(r), P = digamma(r + x) - digamma(r), for r > 0, x > 0.
// Cancellation-free for the common cases (small integer counts; large r).
__device__ __forceinline__ void lgamma_diff(float r, float x, float &D, float &P) {
    if (x <= 8.f && x == floorf(x)) {
        // x integer in [1, 8]: Gamma(r+x)/Gamma(r) = prod_{i<x} (r+i)
        float p = r, dp = 1.f;
        const int n = (int)x;
        for (int i = 1; i < n; ++i) {
            const float f = r + (float)i;
            dp = dp * f + p;
            p = p * f;
        }
        D = logf(p);
        P = __fdividef(dp, p);
    } else if (r >= 8.f) {
        const float s = r + x;
        const float q = __fdividef(x, r);
        // log1p(q) for any q > 0
        const float l1p = q < 0.015625f ? q * (1.f - q * (0.5f - q * (0.33333334f - 0.25f * q)))
                                        : __logf(1.f + q);
        const float ir = __frcp_rn(r), is = __frcp_rn(s);
        const float ir2 = ir * ir, is2 = is * is;
        D = x * logf(s) + ((r - 0.5f) * l1p - x) +
            (is * (0.083333336f - is2 * 0.0027777778f) - ir * (0.083333336f - ir2 * 0.0027777778f));
        P = l1p - 0.5f * (is - ir) -
            (is2 * (0.083333336f - is2 * (0.008333334f - is2 * 0.003968254f)) -
             ir2 * (0.083333336f - ir2 * (0.008333334f - ir2 * 0.003968254f)));
    } else {
        float lg1, dg1, lg0, dg0;
        lgamma_digamma_pos(r + x, lg1, dg1);
        lgamma_digamma_pos(r, lg0, dg0);
        D = lg1 - lg0;
        P = dg1 - dg0;
    }
}

}  // namespace scvae
