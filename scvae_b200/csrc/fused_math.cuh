// Per-element math of the fused heads epilogue (heads_fused.cu): the count likelihoods of
// DU:206-305 / ZI:194-199 and their gradients written for the fewest issue slots.
//
// The reference clips every head (DU:31-50, VAE:2481-2485): logits to >= log(tiny), log_r and
// log_lambda to [-10, 10].  A 16-gene chunk whose pre-activations all lie strictly inside those
// ranges (|logit| <= 80, |log| <= 10: every chunk of a healthy model) takes the FAST path below,
// where the clips and their gradient masks are identities and drop out; any other chunk takes
// the exact, masked path of likelihood_math.cuh.  Both compute the same function.
//
// FAST path, per (cell, gene), NB:   t = e^a_p, u = 1 + t, r = e^a_r
//     log p   = x a_p - (x + r) log u + [x == 1] a_r   (+ lgamma fix-up for x >= 2)
//     d/da_p  = x - (x + r) t / u
//     d/da_r  = -r log u + [x == 1]                     (+ digamma fix-up for x >= 2)
// 4 MUFU (2 ex2, lg2, rcp) and ~17 FP32 instructions; the sums over genes are kept as
// accA + (-ln 2) accB so that log u never leaves the log2 domain.
#pragma once

#include <cuda_fp16.h>

#include "likelihood_math.cuh"

namespace scvae {

// |logit| <= 40: e^40 (1 + e^40) still fits fp32, which lets the zero-inflated forms below share
// one reciprocal between pi and the zero-count responsibility
constexpr float kFastLogitMax = 40.f;
constexpr float kFastLogMax = 10.f;     // the reference's clip of log_r / log_lambda

// two 16-bit targets of one 32-bit word -> fp32 (u16 via the 2^23 mantissa trick: no I2F)
template <bool T_HALF>
__device__ __forceinline__ void fused_cvt2(uint32_t w, float &lo, float &hi) {
    if (T_HALF) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w));
        lo = f.x;
        hi = f.y;
    } else {
        lo = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7410)) - 8388608.f;
        hi = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7432)) - 8388608.f;
    }
}
// bits 14 / 30 set where the low / high target of the word is >= 2 (integer counts)
template <bool T_HALF>
__device__ __forceinline__ uint32_t fused_flags2(uint32_t w) {
    if (T_HALF) return w & 0x40004000u;           // fp16 >= 2.0  <=>  exponent bit 14
    const uint32_t t = w & 0xFFFEFFFEu;
    const uint32_t s = ((t & 0x7FFF7FFFu) + 0x7FFF7FFFu) | t;   // bit 15 / 31: half != 0
    return (s >> 1) & 0x40004000u;
}

// D = lgamma(r + x) - lgamma(r) and Pd = digamma(r + x) - digamma(r) for r > 0, x >= 2 (any real x),
// branch-free: Stirling series at s = r + x and z0 = r + 2 (both >= 2, truncation < 5e-6) with the
// difference taken in cancellation-free form ((z0 - 1/2) log1p(d / z0) - d, d = x - 2), plus
// log(r (r + 1)) resp. (2 r + 1) / (r (r + 1)) for the shift by 2.  x == 2 is exact.  6 MUFU.
__device__ __forceinline__ void lgamma_diff_ge2(float r, float x, float &D, float &Pd) {
    const float d = x - 2.f, z0 = r + 2.f, s = z0 + d, qr = r * (r + 1.f);
    const float i0 = fast_rcp(z0), is = fast_rcp(s), iq = fast_rcp(qr);
    const float q = d * i0;
    const float l_ser = q * (1.f - q * (0.5f - q * (0.33333334f - 0.25f * q)));
    const float l_log = kLn2 * fast_lg2(1.f + q);
    const float L = q < 0.015625f ? l_ser : l_log;          // log(s / z0)
    const float ls = kLn2 * fast_lg2(s), lq = kLn2 * fast_lg2(qr);
    const float is2 = is * is, i02 = i0 * i0;
    const float cs = is * (0.083333336f - is2 * (0.0027777778f - is2 * 0.00079365080f));
    const float c0 = i0 * (0.083333336f - i02 * (0.0027777778f - i02 * 0.00079365080f));
    D = (fmaf(z0 - 0.5f, L, -d) + d * ls) + (cs - c0) + lq;
    const float ts = fmaf(0.5f, is, is2 * (0.083333336f - is2 * (0.008333334f - is2 * 0.003968254f)));
    const float t0 = fmaf(0.5f, i0, i02 * (0.083333336f - i02 * (0.008333334f - i02 * 0.003968254f)));
    Pd = (L - (ts - t0)) + fmaf(2.f, r, 1.f) * iq;
}

// The same pair for integer 2 <= n <= 6 as a predicated rising product (Gamma(r+n)/Gamma(r) =
// prod_{i<n} (r+i); its logarithmic derivative is dp/p): 2 MUFU instead of 6.
constexpr int kProdMax = 6;
__device__ __forceinline__ void lgamma_diff_prod(float r, float x, float &D, float &Pd) {
    float p = r, dp = 1.f;
#pragma unroll
    for (int i = 1; i < kProdMax; ++i) {
        const float f = r + (float)i;
        const bool on = x > (float)i;
        const float dpn = fmaf(dp, f, p), pn = p * f;
        dp = on ? dpn : dp;
        p = on ? pn : p;
    }
    D = kLn2 * fast_lg2(p);
    Pd = dp * fast_rcp(p);
}

// 1 / u for u in [1, 2^120) on the FMA pipe (MUFU is the busiest pipe of the fused epilogue):
// exponent-flip seed (12 % error) and two Newton steps -> 2e-4 relative, for fp16 gradients only.
__device__ __forceinline__ float rcp_newton2(float u) {
    float y = __uint_as_float(0x7EF311C7u - __float_as_uint(u));
    y = fmaf(y, fmaf(-u, y, 1.f), y);
    y = fmaf(y, fmaf(-u, y, 1.f), y);
    return y;
}

// lgamma(1 + x) for x >= 2: Stirling series at z = 1 + x >= 3 (truncation < 3e-7), branch-free.
__device__ __forceinline__ float lgamma1p_ge2(float x) {
    const float z = x + 1.f;
    const float iz = fast_rcp(z), iz2 = iz * iz;
    return fmaf(z - 0.5f, kLn2 * fast_lg2(z), -z) + kHalfLog2Pi +
           iz * (0.083333336f - iz2 * (0.0027777778f - iz2 * 0.00079365080f));
}

// 8 (cell, gene) terms of one cell, no clip active.  accA += natural-log terms,
// accB += log2 terms (caller applies -ln 2).  g[h][j] = d log p / d a_h * gsv.
// r[j] = total_count (NB kinds) for the x >= 2 fix-ups.
template <int KIND>
__device__ __forceinline__ void fused_fast8(const float (&x)[8], const float (&a)[3][8], float gsv,
                                            float &accA, float &accB, float (&r)[8], float (&g)[3][8]) {
    using T = Lik<KIND>;
    constexpr int iD = T::ZI ? 1 : 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float xj = x[j];
        if (!T::ZI) {
            if (T::NB) {
                const float ap = a[0][j], ar = a[1][j];
                const float t = fast_ex2(ap * kLog2e);
                const float rj = fast_ex2(ar * kLog2e);
                const float u = 1.f + t;
                const float lg = fast_lg2(u);
                const float p = t * fast_rcp(u);
                const float xr = xj + rj;
                const float one = xj == 1.f ? 1.f : 0.f;
                accA = fmaf(xj, ap, accA);
                accA = fmaf(one, ar, accA);
                accB = fmaf(xr, lg, accB);
                g[0][j] = fmaf(-xr, p, xj) * gsv;
                g[1][j] = fmaf(rj * lg, -kLn2, one) * gsv;
                r[j] = rj;
            } else {
                const float al = a[0][j];
                const float lam = fast_ex2(al * kLog2e);
                accA = fmaf(xj, al, accA) - lam;
                g[0][j] = (xj - lam) * gsv;
                r[j] = 0.f;
            }
        } else {
            // zero inflation, pi = sigmoid(a_pi):
            //   x > 0: log p = lp_d - softplus(a_pi);  x == 0: softplus(l0 - a_pi) + a_pi - softplus(a_pi)
            // With tp = e^{a_pi}, up = 1 + tp, e0 = e^{l0} = p_d(0), s = tp + e0:
            //   pi = tp / up,  x == 0: log p = log(s / up), responsibility of the count part
            //   wz = e0 / s;   x > 0: log p = lp_d + log(1 / up).
            // One reciprocal (of s up) and one log2 serve both cases: 4 MUFU for the inflation.
            const float api = a[0][j];
            const float tp = fast_ex2(api * kLog2e);
            const float up = 1.f + tp;
            float l0, lpd, gd0, gd1 = 0.f;
            if (T::NB) {
                const float ap = a[iD][j], ar = a[iD + 1][j];
                const float t = fast_ex2(ap * kLog2e);
                const float rj = fast_ex2(ar * kLog2e);
                const float u = 1.f + t;
                const float sp = fast_lg2(u) * kLn2;
                const float p = t * fast_rcp(u);
                const float xr = xj + rj;
                const float one = xj == 1.f ? 1.f : 0.f;
                l0 = -rj * sp;
                lpd = fmaf(one, ar, fmaf(xj, ap, -xr * sp));
                gd0 = fmaf(-xr, p, xj);
                gd1 = l0 + one;
                r[j] = rj;
            } else {
                const float al = a[iD][j];
                const float lam = fast_ex2(al * kLog2e);
                l0 = -lam;
                lpd = fmaf(xj, al, -lam);
                gd0 = xj - lam;
                r[j] = 0.f;
            }
            const float e0 = fast_ex2(l0 * kLog2e);
            const float sz = tp + e0;
            const float inv = fast_rcp(sz * up);
            const float inv_up = inv * sz, inv_s = inv * up;
            const float pi = tp * inv_up;
            const float wz = e0 * inv_s;
            const bool pos = xj > 0.f;
            accA += pos ? lpd : 0.f;
            accB -= fast_lg2(pos ? inv_up : sz * inv_up);
            g[0][j] = ((pos ? 0.f : 1.f - wz) - pi) * gsv;
            const float w = pos ? gsv : wz * gsv;
            g[1][j] = gd0 * w;
            g[2][j] = gd1 * w;
        }
    }
}

// Forward-only form of fused_fast8 (evaluation passes): log p only, no reciprocal.
template <int KIND>
__device__ __forceinline__ void fused_fwd8(const float (&x)[8], const float (&a)[3][8], float &accA, float &accB,
                                           float (&r)[8]) {
    using T = Lik<KIND>;
    constexpr int iD = T::ZI ? 1 : 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float xj = x[j];
        float l0, lpd;          // log p_d(0) and log p_d(x) of the wrapped distribution (natural log)
        if (T::NB) {
            const float ap = a[iD][j], ar = a[iD + 1][j];
            const float t = fast_ex2(ap * kLog2e);
            const float rj = fast_ex2(ar * kLog2e);
            const float lg = fast_lg2(1.f + t);
            const float xr = xj + rj;
            const float one = xj == 1.f ? 1.f : 0.f;
            r[j] = rj;
            if (!T::ZI) {
                accA = fmaf(xj, ap, accA);
                accA = fmaf(one, ar, accA);
                accB = fmaf(xr, lg, accB);
                continue;
            }
            const float sp = lg * kLn2;
            l0 = -rj * sp;
            lpd = fmaf(one, ar, fmaf(xj, ap, -xr * sp));
        } else {
            const float al = a[iD][j];
            const float lam = fast_ex2(al * kLog2e);
            r[j] = 0.f;
            if (!T::ZI) {
                accA = fmaf(xj, al, accA) - lam;
                continue;
            }
            l0 = -lam;
            lpd = fmaf(xj, al, -lam);
        }
        const float api = a[0][j];
        const float lgp = fast_lg2(1.f + fast_ex2(api * kLog2e));
        const float lgu = fast_lg2(1.f + fast_ex2((l0 - api) * kLog2e));
        const bool pos = xj > 0.f;
        accA += pos ? lpd : api;
        accB += pos ? lgp : lgp - lgu;
    }
}

}  // namespace scvae
