// Piecewise-categorical ("categorised") count likelihoods, the `-k` option of the reference
// (CAT:210-274, head P_K VAE:2507-2532): the counts 0 .. k_max - 1 get categorical probabilities
// of their own, counts >= k_max the last class times the count distribution shifted by k_max:
//     log p(x) = log softmax(c)[min(x, k_max)] + [x >= k_max] log p_dist(x - k_max)
// Heads (all at `head_stride` columns from each other): the P heads of the count distribution
// first, then the k_max + 1 class logits (class-major; the reference's P_K variable is
// class-minor, the engine permutes on import / export).  One CTA per (sample, cell) row.
#include "likelihood_math.cuh"

namespace scvae {

constexpr int kMaxClasses = 16;     // k_max + 1

template <int KIND, bool BWD>
__global__ void __launch_bounds__(256)
piecewise_kernel(const float *__restrict__ t, int64_t ldt, int t_rows, const float *__restrict__ a, int64_t lda,
                 int64_t head_stride, int G, int k_max, const float *__restrict__ go, float go_scalar,
                 float *__restrict__ da, int64_t ldda, int64_t dhead_stride, float *__restrict__ logp) {
    using T = Lik<KIND>;
    constexpr int P = T::P;
    __shared__ float red[32];
    const int64_t row = blockIdx.x;
    const float *trp = t + (row % t_rows) * ldt;
    const float *arp = a + row * lda;
    float *drp = BWD ? da + row * ldda : nullptr;
    const float gscale = BWD ? (go ? go[row] : go_scalar) : 0.f;
    const int K1 = k_max + 1;
    const float kf = (float)k_max;
    float acc = 0.f;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        const float x = trp[g];
        // categorical part: log softmax(c)[class], class = clip(x, 0, k_max)
        float c[kMaxClasses];
        float mx = -INFINITY;
        for (int k = 0; k < K1; ++k) {
            c[k] = arp[(int64_t)(P + k) * head_stride + g];
            mx = fmaxf(mx, c[k]);
        }
        float se = 0.f;
        for (int k = 0; k < K1; ++k) se += __expf(c[k] - mx);
        const float lse = mx + __logf(se);
        const int cls = (int)fminf(fmaxf(x, 0.f), kf);
        float ccls = c[0];
        for (int k = 1; k < K1; ++k) ccls = (k == cls) ? c[k] : ccls;
        acc += ccls - lse;
        if (BWD) {
            for (int k = 0; k < K1; ++k)
                drp[(int64_t)(P + k) * dhead_stride + g] = gscale * ((k == cls ? 1.f : 0.f) - __expf(c[k] - lse));
        }
        // count distribution on the shifted count, only for x >= k_max (CAT:262-268)
        const bool tail = !(x < kf);
        float xs[1] = {tail ? x - kf : 0.f}, av[3][1], gv[3][1];
        for (int h = 0; h < P; ++h) av[h][0] = arp[(int64_t)h * head_stride + g];
        float lpd = 0.f;
        lik_group<KIND, BWD, 1>(xs, av, /*has_const=*/false, lpd, gv);
        acc += tail ? lpd : 0.f;
        if (BWD) {
            for (int h = 0; h < P; ++h) drp[(int64_t)h * dhead_stride + g] = tail ? gscale * gv[h][0] : 0.f;
        }
    }
    const float total = block_sum(acc, red);
    if (threadIdx.x == 0 && logp) logp[row] = total;
}

// mean / variance of the wrapped count distribution from its pre-activations (as likelihood.cu)
template <int KIND>
__device__ __forceinline__ void dist_moments(const float (&a)[3], float &m, float &v) {
    using T = Lik<KIND>;
    constexpr int iD = T::ZI ? 1 : 0;
    if (T::NB) {
        const float ap = fmaxf(a[iD], kLogitFloor);
        const float r = __expf(fminf(fmaxf(a[iD + 1], -10.f), 10.f));
        const float ea = __expf(ap);
        m = r * ea;
        v = m * (1.f + ea);
    } else {
        m = __expf(fminf(fmaxf(a[iD], -10.f), 10.f));
        v = m;
    }
    if (T::ZI) {
        const float api = fmaxf(a[0], kLogitFloor);
        const float q = 1.f - __frcp_rn(1.f + __expf(-api));
        const float zm = q * m;
        v = q * (v + m * m) - zm * zm;
        m = zm;
    }
}

// Categorised mean / variance (CAT:210-247) averaged over the RS samples as VAE:2665-2713 and,
// for the GMVAE (K > 1, y = q(y|x) [B, K]), marginalised over the clusters with the y-weighted
// per-cluster mean of GMVAE:3312-3386 (as moments_kernel in likelihood.cu).  Rows of a are
// ordered (k, sample, cell).
template <int KIND>
__global__ void piecewise_moments_kernel(const float *__restrict__ a, int64_t lda, int64_t head_stride, int B, int G,
                                         int RS, int K, const float *__restrict__ y, int64_t ldy, int k_max,
                                         float *__restrict__ p_x_mean, float *__restrict__ p_x_stddev,
                                         float *__restrict__ stddev_of_mean, int64_t ldo) {
    constexpr int P = Lik<KIND>::P;
    const int g = blockIdx.y * blockDim.x + threadIdx.x;
    const int b = blockIdx.x;
    if (g >= G) return;
    const int K1 = k_max + 1;
    const float kf = (float)k_max, inv = 1.f / (float)RS;
    auto moments = [&](int row, float &mean, float &var) {
        const float *ap = a + ((int64_t)row * B + b) * lda + g;
        float c[kMaxClasses], mx = -INFINITY, se = 0.f;
        for (int k = 0; k < K1; ++k) {
            c[k] = ap[(int64_t)(P + k) * head_stride];
            mx = fmaxf(mx, c[k]);
        }
        for (int k = 0; k < K1; ++k) se += __expf(c[k] - mx);
        float av[3] = {ap[0], P > 1 ? ap[head_stride] : 0.f, P > 2 ? ap[2 * head_stride] : 0.f};
        float dm, dv;
        dist_moments<KIND>(av, dm, dv);
        float m1 = 0.f, m2 = 0.f;
        for (int k = 0; k < k_max; ++k) {
            const float pk = __expf(c[k] - mx) / se;
            m1 += (float)k * pk;
            m2 += (float)(k * k) * pk;
        }
        const float pK = __expf(c[k_max] - mx) / se;
        mean = m1 + pK * (dm + kf);
        var = m2 + pK * (2.f * kf * dm + dv + dm * dm + kf * kf) - mean * mean;
    };
    float mean_tot = 0.f, var_of_mean = 0.f, mean_of_var = 0.f;
    for (int k = 0; k < K; ++k) {
        const float w = y ? y[(int64_t)b * ldy + k] : 1.f;
        float ms = 0.f, vs = 0.f;
        for (int s = 0; s < RS; ++s) {
            float m, v;
            moments(k * RS + s, m, v);
            ms += m;
            vs += v;
        }
        const float pm = ms * inv * w;
        float dev = 0.f;
        for (int s = 0; s < RS; ++s) {
            float m, v;
            moments(k * RS + s, m, v);
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

}  // namespace scvae

using namespace scvae;

extern "C" int scvae_piecewise_likelihood(int kind, int k_max, const float *t, int64_t ldt, int t_rows, const float *a,
                                          int64_t lda, int64_t head_stride, int M, int G, const float *go,
                                          float go_scalar, float *da, int64_t ldda, int64_t dhead_stride,
                                          float *logp, void *stream) {
    SCVAE_CHECK_ARG(t && a && M > 0 && G > 0 && t_rows > 0, "piecewise_likelihood: bad arguments");
    SCVAE_CHECK_ARG(k_max >= 1 && k_max + 1 <= kMaxClasses, "piecewise_likelihood: k_max must be in [1, %d]",
                    kMaxClasses - 1);
    SCVAE_CHECK_ARG(logp || da, "piecewise_likelihood: neither logp nor da requested");
    cudaStream_t s = (cudaStream_t)stream;
    switch (kind) {
#define CASE(KK)                                                                                               \
    case KK:                                                                                                   \
        if (da)                                                                                                \
            piecewise_kernel<KK, true><<<M, 256, 0, s>>>(t, ldt, t_rows, a, lda, head_stride, G, k_max, go,    \
                                                         go_scalar, da, ldda, dhead_stride, logp);             \
        else                                                                                                   \
            piecewise_kernel<KK, false><<<M, 256, 0, s>>>(t, ldt, t_rows, a, lda, head_stride, G, k_max,       \
                                                          nullptr, 0.f, nullptr, 0, 0, logp);                  \
        break;
        CASE(SCVAE_LIK_POISSON)
        CASE(SCVAE_LIK_NB)
        CASE(SCVAE_LIK_ZIP)
        CASE(SCVAE_LIK_ZINB)
#undef CASE
        default:
            set_error("piecewise_likelihood: unknown kind %d", kind);
            return 1;
    }
    SCVAE_CHECK_LAUNCH("piecewise_likelihood");
    return 0;
}

extern "C" int scvae_piecewise_moments(int kind, int k_max, const float *a, int64_t lda, int64_t head_stride, int B,
                                       int G, int RS, int K, const float *y, int64_t ldy, float *p_x_mean,
                                       float *p_x_stddev, float *stddev_of_mean, int64_t ldo, void *stream) {
    SCVAE_CHECK_ARG(a && B > 0 && G > 0 && RS > 0 && K > 0, "piecewise_moments: bad arguments");
    SCVAE_CHECK_ARG(K == 1 || y, "piecewise_moments: K > 1 needs cluster weights y");
    SCVAE_CHECK_ARG(k_max >= 1 && k_max + 1 <= kMaxClasses, "piecewise_moments: k_max out of range");
    dim3 grid(B, (G + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
    switch (kind) {
#define CASE(KK)                                                                                          \
    case KK:                                                                                              \
        piecewise_moments_kernel<KK><<<grid, 256, 0, s>>>(a, lda, head_stride, B, G, RS, K, y, ldy, k_max, \
                                                          p_x_mean, p_x_stddev, stddev_of_mean, ldo);     \
        break;
        CASE(SCVAE_LIK_POISSON)
        CASE(SCVAE_LIK_NB)
        CASE(SCVAE_LIK_ZIP)
        CASE(SCVAE_LIK_ZINB)
#undef CASE
        default:
            set_error("piecewise_moments: unknown kind %d", kind);
            return 1;
    }
    SCVAE_CHECK_LAUNCH("piecewise_moments");
    return 0;
}
