// a9/a10: the Gaussian-mixture VAE specific pieces (GMVAE:2788-3434).
//
// The K cluster-conditional passes q(z|x,y=k) / p(x|z_k) share weights, so the engine batches
// them as K consecutive row groups of one tall matrix (rows ordered (k, sample, cell)); the dense
// layers, batch norm (per group) and the count likelihood are the same kernels as for the VAE.
// This file holds what is new:
//   * the one-hot concat [x, e_k] of GMVAE:2942-2947 as a per-group row offset
//     (x W_x + W_y[k] + b) and its gradient reductions -- x W_x is computed once, not K times;
//   * q(y|x) softmax, the softplus-Gaussian reparameterisation with the SAMPLED KL
//     log q(z_k) - log p(z_k|y=k) (GMVAE:3270-3289) and its backward;
//   * the y-marginalised bound with KL_y (+ free nats) and the gradient w.r.t. the logits
//     (GMVAE:3242-3261, 3388-3410).
#include "common.cuh"

namespace scvae {

// ---- per-group row offset ----------------------------------------------------------------
// y[(k*B + b), c] = x[b, c] + t[k, c], c < H;  y may be written augmented (ldy > H).
__global__ void group_offset_fwd_kernel(const float *__restrict__ x, int64_t ldx, const float *__restrict__ t,
                                        int64_t ldt, int K, int B, int H, float *__restrict__ y, int64_t ldy) {
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t row = blockIdx.x;  // k*B + b
    if (c >= H) return;
    const int k = (int)(row / B), b = (int)(row % B);
    y[row * ldy + c] = x[(int64_t)b * ldx + c] + t[(int64_t)k * ldt + c];
}
// dx[b, c] = sum_k dy[k*B+b, c]
__global__ void group_offset_bwd_x_kernel(const float *__restrict__ dy, int64_t lddy, int K, int B, int H,
                                          float *__restrict__ dx, int64_t lddx) {
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    const int b = blockIdx.x;
    if (c >= H) return;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc += dy[((int64_t)k * B + b) * lddy + c];
    dx[(int64_t)b * lddx + c] = acc;
}
// dt[k, c] = sum_b dy[k*B+b, c]   (32 columns x 8 row lanes per CTA, one CTA per (k, col block))
__global__ void __launch_bounds__(256)
group_offset_bwd_t_kernel(const float *__restrict__ dy, int64_t lddy, int B, int H, float *__restrict__ dt,
                          int64_t lddt, int accumulate) {
    __shared__ float sh[8][33];
    const int c = blockIdx.y * 32 + threadIdx.x;
    const int k = blockIdx.x;
    float acc = 0.f;
    if (c < H)
        for (int b = threadIdx.y; b < B; b += 8) acc += dy[((int64_t)k * B + b) * lddy + c];
    sh[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && c < H) {
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) tot += sh[i][threadIdx.x];
        float *o = dt + (int64_t)k * lddt + c;
        *o = accumulate ? *o + tot : tot;
    }
}

// ---- q(y|x): softmax over K clusters -------------------------------------------------------
// y = softmax(logits), logy = log_softmax(logits); one warp per cell.
__global__ void __launch_bounds__(128)
softmax_kernel(const float *__restrict__ logits, int64_t ldl, int B, int K, float *__restrict__ y,
               float *__restrict__ logy) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= B) return;
    const float *lr = logits + (int64_t)b * ldl;
    float mx = -INFINITY;
    for (int k = lane; k < K; k += 32) mx = fmaxf(mx, lr[k]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int k = lane; k < K; k += 32) se += expf(lr[k] - mx);
    se = warp_sum(se);
    const float lse = mx + logf(se);
    for (int k = lane; k < K; k += 32) {
        const float ly = lr[k] - lse;
        logy[(int64_t)b * K + k] = ly;
        y[(int64_t)b * K + k] = expf(ly);
    }
}

// ---- softplus-Gaussian latent with sampled KL ---------------------------------------------
__device__ __forceinline__ float softplus_f(float s) { return fmaxf(s, 0.f) + log1pf(expf(-fabsf(s))); }
__device__ __forceinline__ float sigmoid_f(float s) { return 1.f / (1.f + expf(-s)); }

// qh (K*B, ldq): [mean | softplus_scale] of q(z|x,y=k); pz (K, 2L) contiguous: the same for
// p(z|y=k).  eps (K*RS*B, L) rows ordered (k, rs, b).  z (K*RS*B, ldz) augmented.
// klz[(k*RS + rs)*B + b] = sum_l log q(z) - log p(z|k).  One warp per output row.
__global__ void __launch_bounds__(128)
gmvae_latent_fwd_kernel(const float *__restrict__ qh, int64_t ldq, const float *__restrict__ pz, int K, int B,
                        int L, int RS, const float *__restrict__ eps, float *__restrict__ z, int64_t ldz,
                        float *__restrict__ klz, float *__restrict__ kl_elem) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);  // (k*RS + rs)*B + b
    if (row >= (int64_t)K * RS * B) return;
    const int b = (int)(row % B);
    const int k = (int)(row / ((int64_t)RS * B));
    const float *q = qh + ((int64_t)k * B + b) * ldq;
    const float *p = pz + (int64_t)k * 2 * L;
    float kl = 0.f;
    for (int l = lane; l < L; l += 32) {
        const float mq = q[l], sq = sqrtf(softplus_f(q[L + l]));
        const float mp = p[l], sp = sqrtf(softplus_f(p[L + l]));
        const float e = eps[row * L + l];
        const float zz = mq + sq * e;
        const float u = (zz - mp) / sp;
        const float t = -0.5f * e * e - logf(sq) + 0.5f * u * u + logf(sp);
        kl += t;
        z[row * ldz + l] = zz;
        if (kl_elem) kl_elem[row * L + l] = t;
    }
    for (int c = L + lane; c < ldz; c += 32) z[row * ldz + c] = (c == L) ? 1.f : 0.f;
    kl = warp_sum(kl);
    if (lane == 0) klz[row] = kl;
}

// Backward.  coef[(k*RS+rs)*B + b] = d loss / d klz of that row.  dz: decoder gradient.
// dqh (K*B, lddq): gradient w.r.t. [mean | softplus_scale] pre-activations (sums over rs).
// One warp per (k, b).
__global__ void __launch_bounds__(128)
gmvae_latent_bwd_q_kernel(const float *__restrict__ qh, int64_t ldq, const float *__restrict__ pz, int K, int B,
                          int L, int RS, const float *__restrict__ eps, const float *__restrict__ dz,
                          int64_t lddz, const float *__restrict__ coef, float *__restrict__ dqh, int64_t lddq) {
    const int lane = threadIdx.x & 31;
    const int64_t kb = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (kb >= (int64_t)K * B) return;
    const int b = (int)(kb % B), k = (int)(kb / B);
    const float *q = qh + kb * ldq;
    const float *p = pz + (int64_t)k * 2 * L;
    for (int l = lane; l < L; l += 32) {
        const float mq = q[l], s_raw = q[L + l];
        const float sq = sqrtf(softplus_f(s_raw));
        const float mp = p[l], sp = sqrtf(softplus_f(p[L + l]));
        float dmu = 0.f, dsig = 0.f;
        for (int rs = 0; rs < RS; ++rs) {
            const int64_t row = ((int64_t)k * RS + rs) * B + b;
            const float e = eps[row * L + l];
            const float c = coef[row];
            const float u = (mq + sq * e - mp) / sp;
            const float g = dz[row * lddz + l] + c * u / sp;   // total d loss / d z
            dmu += g;
            dsig += g * e - c / sq;
        }
        float *o = dqh + kb * lddq;
        o[l] = dmu;
        o[L + l] = dsig * sigmoid_f(s_raw) / (2.f * sq);       // d sqrt(softplus(s)) / d s
    }
}

// Gradient w.r.t. the prior table pz (K, 2L): one CTA per cluster, fixed-order reduction.
__global__ void __launch_bounds__(256)
gmvae_latent_bwd_p_kernel(const float *__restrict__ qh, int64_t ldq, const float *__restrict__ pz, int K, int B,
                          int L, int RS, const float *__restrict__ eps, const float *__restrict__ coef,
                          float *__restrict__ dpz) {
    __shared__ float sh_m[8][33], sh_s[8][33];
    const int k = blockIdx.x;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float *p = pz + (int64_t)k * 2 * L;
    for (int l0 = 0; l0 < L; l0 += 32) {
        const int l = l0 + tx;
        float gm = 0.f, gs = 0.f;
        if (l < L) {
            const float mp = p[l], sp_raw = p[L + l];
            const float sp = sqrtf(softplus_f(sp_raw));
            for (int64_t r = ty; r < (int64_t)RS * B; r += 8) {
                const int b = (int)(r % B);
                const int64_t row = (int64_t)k * RS * B + r;
                const float *q = qh + ((int64_t)k * B + b) * ldq;
                const float sq = sqrtf(softplus_f(q[L + l]));
                const float u = (q[l] + sq * eps[row * L + l] - mp) / sp;
                const float c = coef[row];
                gm += -c * u / sp;
                gs += c * (1.f - u * u) / sp;
            }
            gs *= sigmoid_f(sp_raw) / (2.f * sp);
        }
        sh_m[ty][tx] = gm;
        sh_s[ty][tx] = gs;
        __syncthreads();
        if (ty == 0 && l < L) {
            float a = 0.f, c2 = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { a += sh_m[i][tx]; c2 += sh_s[i][tx]; }
            dpz[(int64_t)k * 2 * L + l] = a;
            dpz[(int64_t)k * 2 * L + L + l] = c2;
        }
        __syncthreads();
    }
}

// ---- bound -----------------------------------------------------------------------------------
// out[6] = {lower_bound, lower_bound_weighted, reconstruction_error, kl_divergence_z,
//           kl_divergence_y, kl_divergence_y used in the weighted bound (free nats)}.
// dlogits (nullable, (B, K) contiguous): d(-lower_bound_weighted)/d q(y|x) logits;
// dpy_logits (nullable, [K]): gradient w.r.t. learned prior logits.
// ll_sum / klz_sum: (K, B) = mean over samples of log p(x|z_k) / KL_z,k (not yet y-weighted).
__global__ void __launch_bounds__(256)
gmvae_bound_kernel(const float *__restrict__ y, const float *__restrict__ logy, const float *__restrict__ logp,
                   const float *__restrict__ klz, const float *__restrict__ log_py, int K, int RS, int B,
                   float weight, float free_nats_proportion, int uniform_prior, float *__restrict__ out,
                   float *__restrict__ dlogits, float *__restrict__ dpy_logits, float *__restrict__ ll_mean,
                   float *__restrict__ klz_mean) {
    __shared__ float red[32];
    __shared__ float s_use_kly;
    const float inv_rs = 1.f / (float)RS, inv_b = 1.f / (float)B;
    float s_re = 0.f, s_klz = 0.f, s_kly = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float kly = 0.f;
        for (int k = 0; k < K; ++k) {
            float ll = 0.f, kz = 0.f;
            for (int rs = 0; rs < RS; ++rs) {
                const int64_t row = ((int64_t)k * RS + rs) * B + b;
                ll += logp[row];
                kz += klz[row];
            }
            ll *= inv_rs;
            kz *= inv_rs;
            ll_mean[(int64_t)k * B + b] = ll;
            klz_mean[(int64_t)k * B + b] = kz;
            const float yk = y[(int64_t)b * K + k], ly = logy[(int64_t)b * K + k];
            s_re += yk * ll;
            s_klz += yk * kz;
            kly += yk * (ly - log_py[k]);     // uniform prior: log K - H[q]
        }
        s_kly += kly;
    }
    const float re = block_sum(s_re, red) * inv_b;
    const float kz = block_sum(s_klz, red) * inv_b;
    const float ky = block_sum(s_kly, red) * inv_b;
    // threshold = proportion * H[p(y)] (GMVAE:3260-3261), from the prior on the device: a learnt
    // prior changes every step and the step is replayed from a CUDA graph
    float free_nats_threshold = 0.f, prior_entropy = 0.f;
    if (free_nats_proportion > 0.f) {
        for (int k = 0; k < K; ++k) prior_entropy -= expf(log_py[k]) * log_py[k];
        free_nats_threshold = free_nats_proportion * prior_entropy;
    }
    const bool use_kly = !(free_nats_threshold > 0.f) || ky > free_nats_threshold;
    const float ky_mod = use_kly ? ky : free_nats_threshold;
    if (threadIdx.x == 0) {
        out[0] = re - (kz + ky);
        out[1] = re - weight * (kz + ky_mod);
        out[2] = re;
        out[3] = kz;
        out[4] = ky;
        out[5] = ky_mod;
        s_use_kly = use_kly ? 1.f : 0.f;
    }
    __syncthreads();
    if (!dlogits) return;
    const float wy = weight * s_use_kly;
    // loss = -re + weight*(klz + kly_mod);  d loss / d y_bk = (-ll + weight*klz)/B  (direct)
    //                                       + wy * (logy - log_py + 1)/B           (KL_y)
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float dot = 0.f;
        for (int k = 0; k < K; ++k) {
            const float yk = y[(int64_t)b * K + k];
            const float g = (-ll_mean[(int64_t)k * B + b] + weight * klz_mean[(int64_t)k * B + b] +
                             wy * (logy[(int64_t)b * K + k] - log_py[k] + 1.f)) * inv_b;
            dot += yk * g;
        }
        for (int k = 0; k < K; ++k) {
            const float yk = y[(int64_t)b * K + k];
            const float g = (-ll_mean[(int64_t)k * B + b] + weight * klz_mean[(int64_t)k * B + b] +
                             wy * (logy[(int64_t)b * K + k] - log_py[k] + 1.f)) * inv_b;
            dlogits[(int64_t)b * K + k] = yk * (g - dot);
        }
    }
    if (dpy_logits && !uniform_prior) {
        // d/d prior logits of wy * mean_b sum_k y_bk (logy - log_softmax(prior)_k) = wy (p_k - mean_b y_bk)
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            float ym = 0.f;
            for (int b = 0; b < B; ++b) ym += y[(int64_t)b * K + k];
            float g = wy * (expf(log_py[k]) - ym * inv_b);
            // below the threshold the loss holds weight * proportion * H[p(y)] instead of KL_y, and
            // a learnt prior receives its gradient: dH / d logit_k = -p_k (log p_k + H)
            if (s_use_kly == 0.f) g = -weight * free_nats_proportion * expf(log_py[k]) * (log_py[k] + prior_entropy);
            dpy_logits[k] = g;
        }
    }
}

// go[(k*RS+rs)*B + b] = -y[b,k]/(B RS)   (d loss / d log p(x|z_k));  coef = +weight*y/(B RS).
__global__ void gmvae_row_coef_kernel(const float *__restrict__ y, int K, int RS, int B, float weight,
                                      float *__restrict__ go, float *__restrict__ coef) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (int64_t)K * RS * B) return;
    const int b = (int)(row % B);
    const int k = (int)(row / ((int64_t)RS * B));
    const float v = y[(int64_t)b * K + k] / ((float)B * (float)RS);
    go[row] = -v;
    coef[row] = weight * v;
}

// z_mean[b, l] = sum_k y[b,k] mean_k[b, l]   (GMVAE:2896-2899)
__global__ void gmvae_z_mean_kernel(const float *__restrict__ qh, int64_t ldq, const float *__restrict__ y, int K,
                                    int B, int L, float *__restrict__ z_mean) {
    const int l = blockIdx.y * blockDim.x + threadIdx.x;
    const int b = blockIdx.x;
    if (l >= L) return;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc += y[(int64_t)b * K + k] * qh[((int64_t)k * B + b) * ldq + l];
    z_mean[(int64_t)b * L + l] = acc;
}

}  // namespace scvae

using namespace scvae;

extern "C" int scvae_group_offset_fwd(const float *x, int64_t ldx, const float *t, int64_t ldt, int K, int B,
                                      int H, float *y, int64_t ldy, void *stream) {
    SCVAE_CHECK_ARG(x && t && y && K > 0 && B > 0 && H > 0, "group_offset_fwd: bad arguments");
    const dim3 grid(K * B, (H + 127) / 128);
    group_offset_fwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(x, ldx, t, ldt, K, B, H, y, ldy);
    SCVAE_CHECK_LAUNCH("group_offset_fwd");
    return 0;
}

extern "C" int scvae_group_offset_bwd(const float *dy, int64_t lddy, int K, int B, int H, float *dx,
                                      int64_t lddx, float *dt, int64_t lddt, int accumulate_dt,
                                      void *stream) {
    SCVAE_CHECK_ARG(dy && K > 0 && B > 0 && H > 0, "group_offset_bwd: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    if (dx) {
        const dim3 grid(B, (H + 127) / 128);
        group_offset_bwd_x_kernel<<<grid, 128, 0, s>>>(dy, lddy, K, B, H, dx, lddx);
        SCVAE_CHECK_LAUNCH("group_offset_bwd_x");
    }
    if (dt) {
        const dim3 grid(K, (H + 31) / 32);
        group_offset_bwd_t_kernel<<<grid, dim3(32, 8), 0, s>>>(dy, lddy, B, H, dt, lddt, accumulate_dt);
        SCVAE_CHECK_LAUNCH("group_offset_bwd_t");
    }
    return 0;
}

extern "C" int scvae_softmax_fwd(const float *logits, int64_t ldl, int B, int K, float *y, float *logy,
                                 void *stream) {
    SCVAE_CHECK_ARG(logits && y && logy && B > 0 && K > 0, "softmax_fwd: bad arguments");
    softmax_kernel<<<(B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(logits, ldl, B, K, y, logy);
    SCVAE_CHECK_LAUNCH("softmax_fwd");
    return 0;
}

extern "C" int scvae_gmvae_latent_fwd(const float *qh, int64_t ldq, const float *pz, int K, int B, int L,
                                      int RS, const float *eps, float *z, int64_t ldz, float *klz,
                                      float *kl_elem, void *stream) {
    SCVAE_CHECK_ARG(qh && pz && eps && z && klz && K > 0 && B > 0 && L > 0 && RS > 0,
                    "gmvae_latent_fwd: bad arguments");
    const int64_t rows = (int64_t)K * RS * B;
    gmvae_latent_fwd_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, (cudaStream_t)stream>>>(
        qh, ldq, pz, K, B, L, RS, eps, z, ldz, klz, kl_elem);
    SCVAE_CHECK_LAUNCH("gmvae_latent_fwd");
    return 0;
}

extern "C" int scvae_gmvae_latent_bwd(const float *qh, int64_t ldq, const float *pz, int K, int B, int L,
                                      int RS, const float *eps, const float *dz, int64_t lddz,
                                      const float *coef, float *dqh, int64_t lddq, float *dpz,
                                      void *stream) {
    SCVAE_CHECK_ARG(qh && pz && eps && dz && coef && dqh && dpz, "gmvae_latent_bwd: NULL pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t kb = (int64_t)K * B;
    gmvae_latent_bwd_q_kernel<<<(unsigned)((kb + 3) / 4), 128, 0, s>>>(qh, ldq, pz, K, B, L, RS, eps, dz, lddz,
                                                                      coef, dqh, lddq);
    SCVAE_CHECK_LAUNCH("gmvae_latent_bwd_q");
    gmvae_latent_bwd_p_kernel<<<K, 256, 0, s>>>(qh, ldq, pz, K, B, L, RS, eps, coef, dpz);
    SCVAE_CHECK_LAUNCH("gmvae_latent_bwd_p");
    return 0;
}

extern "C" int scvae_gmvae_row_coefficients(const float *y, int K, int RS, int B, float weight, float *go,
                                            float *coef, void *stream) {
    SCVAE_CHECK_ARG(y && go && coef && K > 0 && RS > 0 && B > 0, "gmvae_row_coefficients: bad arguments");
    const int64_t rows = (int64_t)K * RS * B;
    gmvae_row_coef_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(y, K, RS, B, weight,
                                                                                           go, coef);
    SCVAE_CHECK_LAUNCH("gmvae_row_coefficients");
    return 0;
}

extern "C" int scvae_gmvae_bound(const float *y, const float *logy, const float *logp, const float *klz,
                                 const float *log_py, int K, int RS, int B, float weight,
                                 float free_nats_proportion, int uniform_prior, float *out, float *dlogits,
                                 float *dpy_logits, float *ll_mean, float *klz_mean, void *stream) {
    SCVAE_CHECK_ARG(y && logy && logp && klz && log_py && out && ll_mean && klz_mean,
                    "gmvae_bound: NULL pointer");
    gmvae_bound_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(y, logy, logp, klz, log_py, K, RS, B, weight,
                                                            free_nats_proportion, uniform_prior, out, dlogits,
                                                            dpy_logits, ll_mean, klz_mean);
    SCVAE_CHECK_LAUNCH("gmvae_bound");
    return 0;
}

extern "C" int scvae_gmvae_z_mean(const float *qh, int64_t ldq, const float *y, int K, int B, int L,
                                  float *z_mean, void *stream) {
    SCVAE_CHECK_ARG(qh && y && z_mean && K > 0 && B > 0 && L > 0, "gmvae_z_mean: bad arguments");
    const dim3 grid(B, (L + 63) / 64);
    gmvae_z_mean_kernel<<<grid, 64, 0, (cudaStream_t)stream>>>(qh, ldq, y, K, B, L, z_mean);
    SCVAE_CHECK_LAUNCH("gmvae_z_mean");
    return 0;
}
