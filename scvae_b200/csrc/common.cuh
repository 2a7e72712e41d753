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

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// A kernel launched with launch_pdl(site, ...) may be SCHEDULED while its predecessor in the stream
// (or in the captured graph) still runs: its CTAs become resident as the predecessor's CTAs retire
// and park in pdl_wait(), which returns once the whole predecessor grid has completed and flushed.
// Every kernel calls pdl_wait() before anything else, so the data dependencies are those of plain
// stream order; what overlaps is the launch latency and the ramp-up (~3 us per kernel boundary of
// the training step).  pdl_trigger() in the predecessor allows that early scheduling.  Both are
// no-ops in launches without the attribute.  SCVAE_PDL = bit mask of the launch sites (0 = off).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
enum PdlSite { kPdlGemm = 1, kPdlReduce = 2, kPdlMidFwd = 4, kPdlHeads = 8, kPdlFinish = 16, kPdlMidBwd = 32,
               kPdlAdam = 64 };
int pdl_mask();
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(int site, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                     cudaStream_t s, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl_mask() & site) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

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
    // not volatile: read-only data, so the compiler may hoist these above earlier stores (MLP)
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
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
// TF Adam's step size lr * sqrt(1 - beta2^t) / (1 - beta1^t) in double: 1 - beta^t = -expm1(t ln beta)
// (exact for small t; cheaper than two double pow in a prologue that every optimiser CTA pays).
// One definition for the single-GPU and the data-parallel optimiser kernels: their updates must agree
// bit for bit.
__device__ __forceinline__ float adam_step_size(double lr_eff, float beta1, float beta2, double t) {
    return (float)(lr_eff * sqrt(-expm1(t * log((double)beta2))) / (-expm1(t * log((double)beta1))));
}

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
// single-MUFU special functions (no denormal fix-up code around them)
__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_exp(float x) { return fast_ex2(x * kLog2e); }
// log1p(e) for e in [0, 1]: lg2(u) ln2 with u = 1 + e, plus the first-order compensation of
// the rounding of u (e - (u - 1) is exact), which keeps small e accurate without a series.
__device__ __forceinline__ float log1p_from_u(float e, float u) {
    return fmaf(fast_lg2(u), kLn2, e - (u - 1.f));
}
// softplus(a) = log(1 + exp(a)); also returns e = exp(-|a|) and u = 1 + e for the sigmoid.
__device__ __forceinline__ float softplus_eu(float a, float &e, float &u) {
    e = fast_ex2(-fabsf(a) * kLog2e);
    u = 1.f + e;
    return fmaxf(a, 0.f) + log1p_from_u(e, u);
}
__device__ __forceinline__ float softplus_e(float a, float &e) {
    float u;
    return softplus_eu(a, e, u);
}
// sigmoid(a) from e = exp(-|a|), u = 1 + e
__device__ __forceinline__ float sigmoid_from_eu(float a, float e, float u) {
    const float inv = fast_rcp(u);
    return a >= 0.f ? inv : e * inv;
}
__device__ __forceinline__ float sigmoid_from_e(float a, float e) { return sigmoid_from_eu(a, e, 1.f + e); }

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
// Slow path (large or non-integer x): kept out of line, it is rare.
static __device__ __noinline__ void lgamma_diff_slow(float r, float x, float &D, float &P) {
    if (r >= 8.f) {  // cancellation-free Stirling difference
        const float s = r + x;
        const float q = __fdividef(x, r);
        const float l1p = q < 0.015625f ? q * (1.f - q * (0.5f - q * (0.33333334f - 0.25f * q)))
                                        : logf(1.f + q);
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
// Fast path: x integer in [1, 8] (almost every non-zero single-cell count):
// Gamma(r+x)/Gamma(r) = prod_{i<x} (r+i); its log-derivative is dp/p.
__device__ __forceinline__ void lgamma_diff(float r, float x, float &D, float &P) {
    const int n = (int)x;
    if (x == (float)n && n <= 8) {
        float p = r, dp = 1.f;
#pragma unroll 1
        for (int i = 1; i < n; ++i) {   // typical counts are 2-3: one or two trips
            const float f = r + (float)i;
            dp = fmaf(dp, f, p);
            p = p * f;
        }
        D = fast_lg2(p) * kLn2;
        P = dp * fast_rcp(p);
    } else {
        lgamma_diff_slow(r, x, D, P);
    }
}

}  // namespace scvae
