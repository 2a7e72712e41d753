// Constrained Poisson likelihood (DU "constrained poisson", VAE:2492-2496): one head, softmax over
// the genes of a cell, rate = N_b * clip(softmax(a)_g, tiny, 1 - tiny) with N_b the count sum of
// the cell (fed as `count_sum_parameter`, VAE:1018-1019):
//     log p(x | a, N) = sum_g [ x_g log(N s_g) - N s_g - lgamma(1 + x_g) ]
//     d/da_j          = m_j (x_j - N s_j) - s_j sum_g m_g (x_g - N s_g)      (m = clip mask)
// Unlike the other count likelihoods this one couples all genes of a row (the softmax), so it is a
// row kernel: one CTA per (sample, cell) row, three streaming passes over the row (log-sum-exp;
// log p and the coupling sum; gradient), the row staying L2-resident between passes.
#include "common.cuh"

namespace scvae {

constexpr float kLogTiny = -87.33654475f;    // log(float32 tiny): lower clip of the softmax output

template <bool BWD>
__global__ void __launch_bounds__(256)
constrained_poisson_kernel(const float *__restrict__ t, int64_t ldt, int t_rows, const float *__restrict__ a,
                           int64_t lda, int M, int G, const float *__restrict__ count_sum,
                           const float *__restrict__ row_const, const float *__restrict__ go, float go_scalar,
                           float *__restrict__ da, int64_t ldda, float *__restrict__ logp,
                           float *__restrict__ lse_out) {
    __shared__ float red[32];
    const int m = blockIdx.x;
    const int tr = m % t_rows;
    const float *ar = a + (int64_t)m * lda;
    const float *xr = t + (int64_t)tr * ldt;
    const float N = count_sum[tr];
    // ---- pass 1: log-sum-exp of the row (online max / sum per thread, then block combine) ----
    float mx = -INFINITY, sm = 0.f;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        const float v = ar[g];
        if (v > mx) {
            sm = sm * __expf(mx - v) + 1.f;
            mx = v;
        } else {
            sm += __expf(v - mx);
        }
    }
    float bmx = warp_max(mx);
    {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        __syncthreads();
        if (lane == 0) red[wid] = bmx;
        __syncthreads();
        float r = lane < (int)(blockDim.x >> 5) ? red[lane] : -INFINITY;
        bmx = warp_max(r);
        bmx = __shfl_sync(0xffffffffu, bmx, 0);
    }
    const float tot = block_sum(mx == -INFINITY ? 0.f : sm * __expf(mx - bmx), red);
    const float lse = bmx + __logf(tot);
    if (lse_out && threadIdx.x == 0) lse_out[m] = lse;
    // ---- pass 2: log p and the coupling sum S1 = sum_g m_g (x_g - N s_g) ----------------------
    const float logN = N > 0.f ? __logf(N) : 0.f;
    float lp = 0.f, s1 = 0.f;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        const float x = xr[g];
        const float ls = ar[g] - lse;
        const bool in = ls >= kLogTiny;
        const float lsc = in ? ls : kLogTiny;
        const float rate = N * __expf(lsc);
        lp += (x > 0.f ? x * (logN + lsc) : 0.f) - rate;      // (x = 0: 0 log 0 taken as 0)
        if (!row_const && x > 0.f) lp -= lgammaf(1.f + x);
        if (in) s1 += x - rate;
    }
    lp = block_sum(lp, red);
    s1 = block_sum(s1, red);
    if (logp && threadIdx.x == 0) logp[m] = lp - (row_const ? row_const[tr] : 0.f);
    if (!BWD) return;
    // ---- pass 3: gradient -------------------------------------------------------------------
    const float gm = go ? go[m] : go_scalar;
    float *dr = da + (int64_t)m * ldda;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        const float x = xr[g];
        const float ls = ar[g] - lse;
        const float s = __expf(ls);
        const float own = ls >= kLogTiny ? x - N * s : 0.f;
        dr[g] = gm * (own - s * s1);
    }
}

// p_x_mean, p_x_stddev, stddev_of_p_x_given_z_mean (VAE:2665-2713) for rate = N softmax(a):
// Poisson mean = variance = rate; lse [RS * B] from the forward kernel.
__global__ void constrained_poisson_moments_kernel(const float *__restrict__ a, int64_t lda,
                                                   const float *__restrict__ lse,
                                                   const float *__restrict__ count_sum, int B, int G, int RS,
                                                   float *__restrict__ p_x_mean, float *__restrict__ p_x_stddev,
                                                   float *__restrict__ stddev_of_mean, int64_t ldo) {
    const int g = blockIdx.y * blockDim.x + threadIdx.x;
    const int b = blockIdx.x;
    if (g >= G) return;
    const float N = count_sum[b], inv = 1.f / (float)RS;
    float ms = 0.f;
    for (int s = 0; s < RS; ++s) {
        const int64_t m = (int64_t)s * B + b;
        ms += N * __expf(fmaxf(a[m * lda + g] - lse[m], kLogTiny));
    }
    const float mean = ms * inv;
    float dev = 0.f;
    for (int s = 0; s < RS; ++s) {
        const int64_t m = (int64_t)s * B + b;
        const float r = N * __expf(fmaxf(a[m * lda + g] - lse[m], kLogTiny));
        dev += (r - mean) * (r - mean);
    }
    const int64_t o = (int64_t)b * ldo + g;
    if (p_x_mean) p_x_mean[o] = mean;
    if (p_x_stddev) p_x_stddev[o] = sqrtf(dev * inv + mean);
    if (stddev_of_mean) stddev_of_mean[o] = sqrtf(dev * inv);
}

// The same moments marginalised over the K clusters of the GMVAE (GMVAE:3312-3386, with the
// y-weighted per-cluster mean of quirk Q7): rows ordered (k, sample, cell), y = q(y|x) (B, ldy).
__global__ void constrained_poisson_mixture_moments_kernel(
    const float *__restrict__ a, int64_t lda, const float *__restrict__ lse,
    const float *__restrict__ count_sum, int B, int G, int RS, int K, const float *__restrict__ y,
    int64_t ldy, float *__restrict__ p_x_mean, float *__restrict__ p_x_stddev,
    float *__restrict__ stddev_of_mean, int64_t ldo) {
    const int g = blockIdx.y * blockDim.x + threadIdx.x;
    const int b = blockIdx.x;
    if (g >= G) return;
    const float N = count_sum[b], inv = 1.f / (float)RS;
    float mean_total = 0.f, mean_of_var = 0.f, var_of_mean = 0.f;
    for (int k = 0; k < K; ++k) {
        const float w = y[(int64_t)b * ldy + k];
        float ms = 0.f;
        for (int s = 0; s < RS; ++s) {
            const int64_t m = ((int64_t)k * RS + s) * B + b;
            ms += N * __expf(fmaxf(a[m * lda + g] - lse[m], kLogTiny));
        }
        const float c = w * ms * inv;          // y-weighted mean of E[x|z_k] over the samples
        float dev = 0.f;
        for (int s = 0; s < RS; ++s) {
            const int64_t m = ((int64_t)k * RS + s) * B + b;
            const float r = N * __expf(fmaxf(a[m * lda + g] - lse[m], kLogTiny));
            dev += (r - c) * (r - c);
        }
        mean_total += c;
        mean_of_var += c;                      // Poisson: Var[x|z] = E[x|z]
        var_of_mean += w * dev * inv;
    }
    const int64_t o = (int64_t)b * ldo + g;
    if (p_x_mean) p_x_mean[o] = mean_total;
    if (p_x_stddev) p_x_stddev[o] = sqrtf(mean_of_var + var_of_mean);
    if (stddev_of_mean) stddev_of_mean[o] = sqrtf(var_of_mean);
}

}  // namespace scvae

using namespace scvae;

extern "C" int scvae_constrained_poisson(const float *t, int64_t ldt, int t_rows, const float *a, int64_t lda, int M,
                                         int G, const float *count_sum, const float *row_const, const float *go,
                                         float go_scalar, float *da, int64_t ldda, float *logp, float *lse,
                                         void *stream) {
    SCVAE_CHECK_ARG(t && a && count_sum && M > 0 && G > 0 && t_rows > 0, "constrained_poisson: bad arguments");
    SCVAE_CHECK_ARG(logp || da, "constrained_poisson: neither logp nor da requested");
    cudaStream_t s = (cudaStream_t)stream;
    if (da)
        constrained_poisson_kernel<true><<<M, 256, 0, s>>>(t, ldt, t_rows, a, lda, M, G, count_sum, row_const, go,
                                                           go_scalar, da, ldda, logp, lse);
    else
        constrained_poisson_kernel<false><<<M, 256, 0, s>>>(t, ldt, t_rows, a, lda, M, G, count_sum, row_const,
                                                            nullptr, 0.f, nullptr, 0, logp, lse);
    SCVAE_CHECK_LAUNCH("constrained_poisson");
    return 0;
}

extern "C" int scvae_constrained_poisson_moments(const float *a, int64_t lda, const float *lse,
                                                 const float *count_sum, int B, int G, int RS, float *p_x_mean,
                                                 float *p_x_stddev, float *stddev_of_mean, int64_t ldo,
                                                 void *stream) {
    SCVAE_CHECK_ARG(a && lse && count_sum && B > 0 && G > 0 && RS > 0, "constrained_poisson_moments: bad arguments");
    dim3 grid(B, (G + 255) / 256);
    constrained_poisson_moments_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, lda, lse, count_sum, B, G, RS,
                                                                               p_x_mean, p_x_stddev,
                                                                               stddev_of_mean, ldo);
    SCVAE_CHECK_LAUNCH("constrained_poisson_moments");
    return 0;
}

extern "C" int scvae_constrained_poisson_mixture_moments(
    const float *a, int64_t lda, const float *lse, const float *count_sum, int B, int G, int RS,
    int K, const float *y, int64_t ldy, float *p_x_mean, float *p_x_stddev, float *stddev_of_mean,
    int64_t ldo, void *stream) {
    SCVAE_CHECK_ARG(a && lse && count_sum && y && B > 0 && G > 0 && RS > 0 && K > 0 && ldy >= K,
                    "constrained_poisson_mixture_moments: bad arguments");
    dim3 grid(B, (G + 255) / 256);
    constrained_poisson_mixture_moments_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        a, lda, lse, count_sum, B, G, RS, K, y, ldy, p_x_mean, p_x_stddev, stddev_of_mean, ldo);
    SCVAE_CHECK_LAUNCH("constrained_poisson_mixture_moments");
    return 0;
}
