// Continuous / binary reconstruction distributions: the `-r` choices of the reference
// besides the count family -- gaussian, softplus ("modified") gaussian, log-normal, gamma, bernoulli,
// lomax, exponentially modified gaussian (DU:31-73, :125-245; distributions/lomax.py:177-247,
// distributions/exponentially_modified_normal.py:196-234).
//
// Same contract as the count-likelihood kernels: one CTA per (sample, cell) row streams the target
// row and the P head rows (pre-activations a, `head_stride` columns apart) with 128-bit loads,
// applies the head's activation and clip (VAE:2466-2489: theta = clip(act(a), lo + tiny, hi - tiny)),
// sums log p over the genes with a fixed-order block reduction and, backward, writes
// go * d log p / d a (clip gradients: zero outside the clip, as tf.clip_by_value).  HBM-bound:
// (1 + P) * 4 bytes per (cell, gene) forward, (1 + 2 P) * 4 forward + backward.
#include "common.cuh"

namespace scvae {

constexpr float kTiny = 1.17549435e-38f;      // numpy.finfo(float32).tiny
constexpr float kLog2Pi = 1.8378770664093453f;

template <int KIND>
struct Cont;
template <> struct Cont<SCVAE_LIK_GAUSSIAN> { static constexpr int P = 2; };
template <> struct Cont<SCVAE_LIK_SOFTPLUS_GAUSSIAN> { static constexpr int P = 2; };
template <> struct Cont<SCVAE_LIK_LOG_NORMAL> { static constexpr int P = 2; };
template <> struct Cont<SCVAE_LIK_GAMMA> { static constexpr int P = 2; };
template <> struct Cont<SCVAE_LIK_BERNOULLI> { static constexpr int P = 1; };
template <> struct Cont<SCVAE_LIK_LOMAX> { static constexpr int P = 2; };
template <> struct Cont<SCVAE_LIK_EMG> { static constexpr int P = 3; };

// softplus(a) clipped to [tiny, inf) with its derivative w.r.t. a (0 where the clip is active)
__device__ __forceinline__ float softplus_clip(float a, float &d) {
    const float sp = fmaxf(a, 0.f) + log1pf(__expf(-fabsf(a)));
    d = sp < kTiny ? 0.f : __frcp_rn(1.f + __expf(-a));
    return fmaxf(sp, kTiny);
}
// digamma(z), z > 0: recurrence up to z >= 6, then the asymptotic series
__device__ __forceinline__ float digammaf_pos(float z) {
    float acc = 0.f;
    while (z < 6.f) {
        acc -= __frcp_rn(z);
        z += 1.f;
    }
    const float iz = __frcp_rn(z), iz2 = iz * iz;
    return acc + logf(z) - 0.5f * iz - iz2 * (0.083333336f - iz2 * (0.008333334f - iz2 * 0.003968254f));
}

// log p(x | a) and d log p / d a[h] of one (cell, gene) term
template <int KIND, bool BWD>
__device__ __forceinline__ float cont_elem(float x, const float (&a)[3], float (&g)[3]) {
    if (KIND == SCVAE_LIK_GAUSSIAN) {
        // Normal(mu, exp(clip(log_sigma, -3, 3)))
        const float ls = fminf(fmaxf(a[1], -3.f), 3.f);
        const float inv = __expf(-ls);
        const float zz = (x - a[0]) * inv;
        if (BWD) {
            g[0] = zz * inv;
            g[1] = (fabsf(a[1]) > 3.f) ? 0.f : zz * zz - 1.f;
        }
        return -0.5f * zz * zz - ls - 0.5f * kLog2Pi;
    }
    if (KIND == SCVAE_LIK_SOFTPLUS_GAUSSIAN) {
        // Normal(mean, sqrt(softplus(s))): v = softplus(s) is the variance
        const float s = a[1];
        const float v = fmaxf(s, 0.f) + log1pf(__expf(-fabsf(s)));
        const float dlt = x - a[0];
        const float iv = __frcp_rn(v);
        if (BWD) {
            g[0] = dlt * iv;
            g[1] = 0.5f * (dlt * dlt * iv - 1.f) * iv * __frcp_rn(1.f + __expf(-s));
        }
        return -0.5f * dlt * dlt * iv - 0.5f * logf(v) - 0.5f * kLog2Pi;
    }
    if (KIND == SCVAE_LIK_LOG_NORMAL) {
        // LogNormal(mean, sqrt(variance)), variance = clip(softplus(a1), tiny, inf):
        // Normal(mean, s).log_prob(log x) - log x  (non-finite for x <= 0, as in the reference)
        float dv;
        const float v = softplus_clip(a[1], dv);
        const float lx = logf(x);
        const float dlt = lx - a[0];
        const float iv = __frcp_rn(v);
        if (BWD) {
            g[0] = dlt * iv;
            g[1] = 0.5f * (dlt * dlt * iv - 1.f) * iv * dv;
        }
        return -0.5f * dlt * dlt * iv - 0.5f * logf(v) - 0.5f * kLog2Pi - lx;
    }
    if (KIND == SCVAE_LIK_GAMMA) {
        // Gamma(concentration, rate): xlogy(c - 1, x) - rate x - lgamma(c) + c log(rate)
        float dc, dr;
        const float c = softplus_clip(a[0], dc), r = softplus_clip(a[1], dr);
        const float lx = logf(x), lr = logf(r);
        const float xly = (c == 1.f) ? 0.f : (c - 1.f) * lx;
        if (BWD) {
            g[0] = (lx + lr - digammaf_pos(c)) * dc;
            g[1] = (c * __frcp_rn(r) - x) * dr;
        }
        return xly - r * x - lgammaf(c) + c * lr;
    }
    if (KIND == SCVAE_LIK_BERNOULLI) {
        // -sigmoid_cross_entropy_with_logits(labels = x, logits = a)
        const float e = __expf(-fabsf(a[0]));
        if (BWD) g[0] = x - (a[0] >= 0.f ? __frcp_rn(1.f + e) : e * __frcp_rn(1.f + e));
        return -(fmaxf(a[0], 0.f) - a[0] * x + log1pf(e));
    }
    if (KIND == SCVAE_LIK_LOMAX) {
        // Lomax(exp(lc), exp(lsc)), both logs clipped to [-10, 10]
        const float lc = fminf(fmaxf(a[0], -10.f), 10.f), lsc = fminf(fmaxf(a[1], -10.f), 10.f);
        const float c = __expf(lc);
        const float q = x * __expf(-lsc);           // x / scale
        const float l1p = log1pf(q);
        if (BWD) {
            g[0] = (fabsf(a[0]) > 10.f) ? 0.f : 1.f - c * l1p;
            g[1] = (fabsf(a[1]) > 10.f) ? 0.f : (c + 1.f) * q * __frcp_rn(1.f + q) - 1.f;
        }
        return -(c + 1.f) * l1p - (lsc - lc);
    }
    // exponentially modified gaussian: loc, scale = softplus, rate = softplus
    //   u = rate (x - loc), v = rate scale, w = (v^2 - u) / (sqrt(2) v)
    //   log p = -u + v^2 / 2 + log(max(erfc(w), tiny)) - log 2 + log rate
    float dsc, drt;
    const float loc = a[0];
    const float sc = softplus_clip(a[1], dsc), rt = softplus_clip(a[2], drt);
    const float u = rt * (x - loc), v = rt * sc;
    const float w = (v * v - u) * __frcp_rn(1.41421356f * v);
    const float er = erfcf(w);
    const float erc = fmaxf(er, kTiny);
    if (BWD) {
        // d log erfc(w) / d w = -2 / sqrt(pi) exp(-w^2) / erfc(w)  (0 where the clip is active)
        const float dlw = er < kTiny ? 0.f : -1.12837917f * __expf(-w * w) * __frcp_rn(erc);
        // w = (v^2 - u) / (sqrt2 v):  dw/du = -1 / (sqrt2 v),  dw/dv = (v^2 + u) / (sqrt2 v^2)
        const float is2v = __frcp_rn(1.41421356f * v);
        const float dl_du = -1.f - dlw * is2v;
        const float dl_dv = v + dlw * (v * v + u) * is2v * __frcp_rn(v);
        g[0] = dl_du * (-rt);                                       // u = rate (x - loc)
        g[1] = dl_dv * rt * dsc;                                    // v = rate scale
        g[2] = (dl_du * (x - loc) + dl_dv * sc + __frcp_rn(rt)) * drt;
    }
    return -u + 0.5f * v * v + logf(erc) - 0.69314718f + logf(rt);
}

// (mean, variance) of one term (evaluate path)
template <int KIND>
__device__ __forceinline__ void cont_moments(const float (&a)[3], float &m, float &v) {
    float d;
    if (KIND == SCVAE_LIK_GAUSSIAN) {
        m = a[0];
        const float s = __expf(fminf(fmaxf(a[1], -3.f), 3.f));
        v = s * s;
    } else if (KIND == SCVAE_LIK_SOFTPLUS_GAUSSIAN) {
        m = a[0];
        v = fmaxf(a[1], 0.f) + log1pf(__expf(-fabsf(a[1])));
    } else if (KIND == SCVAE_LIK_LOG_NORMAL) {
        const float var = softplus_clip(a[1], d);
        m = __expf(a[0] + 0.5f * var);
        v = (__expf(var) - 1.f) * __expf(2.f * a[0] + var);
    } else if (KIND == SCVAE_LIK_GAMMA) {
        const float c = softplus_clip(a[0], d), r = softplus_clip(a[1], d);
        m = c / r;
        v = m / r;
    } else if (KIND == SCVAE_LIK_BERNOULLI) {
        m = __frcp_rn(1.f + __expf(-a[0]));
        v = m * (1.f - m);
    } else if (KIND == SCVAE_LIK_LOMAX) {
        const float c = __expf(fminf(fmaxf(a[0], -10.f), 10.f)), s = __expf(fminf(fmaxf(a[1], -10.f), 10.f));
        // allow_nan_stats (the reference's default): nan / inf where the moment does not exist
        m = c > 1.f ? s / (c - 1.f) : NAN;
        v = c > 2.f ? s * s * (c - 1.f) / ((c - 1.f) * (c - 1.f) * (c - 2.f)) : (c > 1.f ? INFINITY : NAN);
    } else {
        const float sc = softplus_clip(a[1], d), rt = softplus_clip(a[2], d);
        m = a[0] + __frcp_rn(rt);
        v = sc * sc + __frcp_rn(rt * rt);
    }
}

constexpr int kContThreads = 256;

template <int KIND, bool BWD, bool VEC>
__global__ void __launch_bounds__(kContThreads)
continuous_kernel(const float *__restrict__ t, int64_t ldt, int t_rows, const float *__restrict__ a, int64_t lda,
                  int64_t head_stride, int G, const float *__restrict__ go, float go_scalar, float *__restrict__ da,
                  int64_t ldda, int64_t dhead_stride, float *__restrict__ logp) {
    constexpr int P = Cont<KIND>::P;
    constexpr int W = VEC ? 4 : 1;
    __shared__ float red[32];
    const int64_t row = blockIdx.x;
    const float *trp = t + (row % t_rows) * ldt;
    const float *arp = a + row * lda;
    float *drp = BWD ? da + row * ldda : nullptr;
    const float gscale = BWD ? (go ? go[row] : go_scalar) : 0.f;
    float acc = 0.f;
    for (int c0 = threadIdx.x * W; c0 < G; c0 += kContThreads * W) {
        float x[W], av[3][W], gv[3][W];
        if (VEC) {
            const float4 q = ldg_stream4(trp + c0);
            x[0] = q.x; x[W > 1 ? 1 : 0] = q.y; x[W > 2 ? 2 : 0] = q.z; x[W > 3 ? 3 : 0] = q.w;
#pragma unroll
            for (int h = 0; h < P; ++h) {
                const float4 p4 = ldg_stream4(arp + h * head_stride + c0);
                av[h][0] = p4.x; av[h][W > 1 ? 1 : 0] = p4.y; av[h][W > 2 ? 2 : 0] = p4.z; av[h][W > 3 ? 3 : 0] = p4.w;
            }
        } else {
            x[0] = __ldg(trp + c0);
#pragma unroll
            for (int h = 0; h < P; ++h) av[h][0] = __ldg(arp + h * head_stride + c0);
        }
#pragma unroll
        for (int j = 0; j < W; ++j) {
            const float aj[3] = {av[0][j], P > 1 ? av[1][j] : 0.f, P > 2 ? av[2][j] : 0.f};
            float g[3] = {0.f, 0.f, 0.f};
            acc += cont_elem<KIND, BWD>(x[j], aj, g);
            if (BWD) {
                gv[0][j] = g[0] * gscale;
                gv[1][j] = g[1] * gscale;
                gv[2][j] = g[2] * gscale;
            }
        }
        if (BWD) {
#pragma unroll
            for (int h = 0; h < P; ++h) {
                if (VEC) stg_stream4(drp + h * dhead_stride + c0, make_float4(gv[h][0], gv[h][W > 1 ? 1 : 0], gv[h][W > 2 ? 2 : 0], gv[h][W > 3 ? 3 : 0]));
                else drp[h * dhead_stride + c0] = gv[h][0];
            }
        }
    }
    const float total = block_sum(acc, red);
    if (threadIdx.x == 0 && logp) logp[row] = total;
}

template <int KIND>
__global__ void continuous_moments_kernel(const float *__restrict__ a, int64_t lda, int64_t head_stride, int B, int G,
                                          int RS, int K, const float *__restrict__ y, int64_t ldy,
                                          float *__restrict__ p_x_mean, float *__restrict__ p_x_stddev,
                                          float *__restrict__ stddev_of_mean, int64_t ldo) {
    constexpr int P = Cont<KIND>::P;
    const int g = blockIdx.y * blockDim.x + threadIdx.x;
    const int b = blockIdx.x;
    if (g >= G) return;
    const float inv = 1.f / (float)RS;
    float mean_tot = 0.f, var_of_mean = 0.f, mean_of_var = 0.f;
    for (int k = 0; k < K; ++k) {
        const float w = y ? y[(int64_t)b * ldy + k] : 1.f;
        float ms = 0.f, vs = 0.f;
        for (int s = 0; s < RS; ++s) {
            const float *ap = a + ((int64_t)(k * RS + s) * B + b) * lda + g;
            const float av[3] = {ap[0], P > 1 ? ap[head_stride] : 0.f, P > 2 ? ap[2 * head_stride] : 0.f};
            float m, v;
            cont_moments<KIND>(av, m, v);
            ms += m;
            vs += v;
        }
        const float pm = ms * inv * w;      // y-weighted per-k mean (GMVAE:3323-3329, quirk Q7)
        float dev = 0.f;
        for (int s = 0; s < RS; ++s) {
            const float *ap = a + ((int64_t)(k * RS + s) * B + b) * lda + g;
            const float av[3] = {ap[0], P > 1 ? ap[head_stride] : 0.f, P > 2 ? ap[2 * head_stride] : 0.f};
            float m, v;
            cont_moments<KIND>(av, m, v);
            dev += (m - pm) * (m - pm);
        }
        mean_tot += pm;
        var_of_mean += dev * inv * w;
        mean_of_var += vs * inv * w;
    }
    const int64_t o = (int64_t)b * ldo + g;
    if (p_x_mean) p_x_mean[o] = mean_tot;
    if (p_x_stddev) p_x_stddev[o] = sqrtf(var_of_mean + mean_of_var);
    if (stddev_of_mean) stddev_of_mean[o] = sqrtf(var_of_mean);
}

template <int KIND>
static int launch_cont(const float *t, int64_t ldt, int t_rows, const float *a, int64_t lda, int64_t head_stride, int M,
                       int G, const float *go, float go_scalar, float *da, int64_t ldda, int64_t dhead_stride,
                       float *logp, cudaStream_t s) {
    bool vec = (G % 4 == 0) && aligned16(t) && aligned16(a) && ldt % 4 == 0 && lda % 4 == 0 && head_stride % 4 == 0;
    if (da) vec = vec && aligned16(da) && ldda % 4 == 0 && dhead_stride % 4 == 0;
#define GO(BW, VC)                                                                                                  \
    continuous_kernel<KIND, BW, VC><<<M, kContThreads, 0, s>>>(t, ldt, t_rows, a, lda, head_stride, G, go, go_scalar, \
                                                               da, ldda, dhead_stride, logp)
    if (da) {
        if (vec) GO(true, true);
        else GO(true, false);
    } else {
        if (vec) GO(false, true);
        else GO(false, false);
    }
#undef GO
    SCVAE_CHECK_LAUNCH("continuous_likelihood");
    return 0;
}

}  // namespace scvae

using namespace scvae;

extern "C" int scvae_continuous_num_heads(int kind) {
    switch (kind) {
        case SCVAE_LIK_BERNOULLI: return 1;
        case SCVAE_LIK_GAUSSIAN:
        case SCVAE_LIK_SOFTPLUS_GAUSSIAN:
        case SCVAE_LIK_LOG_NORMAL:
        case SCVAE_LIK_GAMMA:
        case SCVAE_LIK_LOMAX: return 2;
        case SCVAE_LIK_EMG: return 3;
    }
    return -1;
}

extern "C" int scvae_continuous_likelihood(int kind, const float *t, int64_t ldt, int t_rows, const float *a,
                                           int64_t lda, int64_t head_stride, int M, int G, const float *go,
                                           float go_scalar, float *da, int64_t ldda, int64_t dhead_stride,
                                           float *logp, void *stream) {
    SCVAE_CHECK_ARG(t && a && M >= 0 && G > 0 && t_rows > 0 && (logp || da), "continuous_likelihood: bad arguments");
    if (M == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    switch (kind) {
#define CASE(KK) \
    case KK: return launch_cont<KK>(t, ldt, t_rows, a, lda, head_stride, M, G, go, go_scalar, da, ldda, dhead_stride, logp, s);
        CASE(SCVAE_LIK_GAUSSIAN)
        CASE(SCVAE_LIK_SOFTPLUS_GAUSSIAN)
        CASE(SCVAE_LIK_LOG_NORMAL)
        CASE(SCVAE_LIK_GAMMA)
        CASE(SCVAE_LIK_BERNOULLI)
        CASE(SCVAE_LIK_LOMAX)
        CASE(SCVAE_LIK_EMG)
#undef CASE
    }
    set_error("continuous_likelihood: unknown kind %d", kind);
    return 1;
}

extern "C" int scvae_continuous_moments(int kind, const float *a, int64_t lda, int64_t head_stride, int B, int G,
                                        int RS, int K, const float *y, int64_t ldy, float *p_x_mean,
                                        float *p_x_stddev, float *stddev_of_mean, int64_t ldo, void *stream) {
    SCVAE_CHECK_ARG(a && B > 0 && G > 0 && RS > 0 && K > 0, "continuous_moments: bad arguments");
    SCVAE_CHECK_ARG(K == 1 || y, "continuous_moments: K > 1 needs cluster weights y");
    dim3 grid(B, (G + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
    switch (kind) {
#define CASE(KK)                                                                                          \
    case KK:                                                                                              \
        continuous_moments_kernel<KK><<<grid, 256, 0, s>>>(a, lda, head_stride, B, G, RS, K, y, ldy,      \
                                                           p_x_mean, p_x_stddev, stddev_of_mean, ldo);    \
        break;
        CASE(SCVAE_LIK_GAUSSIAN)
        CASE(SCVAE_LIK_SOFTPLUS_GAUSSIAN)
        CASE(SCVAE_LIK_LOG_NORMAL)
        CASE(SCVAE_LIK_GAMMA)
        CASE(SCVAE_LIK_BERNOULLI)
        CASE(SCVAE_LIK_LOMAX)
        CASE(SCVAE_LIK_EMG)
#undef CASE
        default:
            set_error("continuous_moments: unknown kind %d", kind);
            return 1;
    }
    SCVAE_CHECK_LAUNCH("continuous_moments");
    return 0;
}
