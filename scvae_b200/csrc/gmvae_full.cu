// f4: the full-covariance Gaussian mixture of the GMVAE (`-q "full-covariance gaussian mixture"`):
// q(z|x,y=k) and p(z|y=k) are multivariate Gaussians with a lower-triangular scale matrix
// (DU:75-93: L "locations" + L (L + 1) / 2 "scales" per distribution, softplus + clip at tiny, laid out
// by tfp.distributions.fill_triangular; MultivariateNormalTriL of multivariate_normal.py:90-150).
//   z = loc_q + S_q eps,   KL_z = log q(z) - log p(z|y=k)   (sampled, GMVAE:3270-3289)
//   log N(z; loc, S) = -1/2 |S^-1 (z - loc)|^2 - sum_i log S_ii - L/2 log 2 pi
// With z = loc_q + S_q eps the first term of log q is -1/2 |eps|^2 whatever the parameters; log p needs
// a triangular solve per (cluster, sample, cell) row.  One warp per row; the prior's K activated scale
// matrices are expanded once per step.
#include "common.cuh"

namespace scvae {

constexpr int kFullMaxL = 128;          // latent sizes up to 128 (warp-local vectors in shared memory)
constexpr float kTiny = 1.17549435e-38f;

__device__ __forceinline__ float fc_softplus(float s) { return fmaxf(s, 0.f) + log1pf(expf(-fabsf(s))); }
__device__ __forceinline__ float fc_sigmoid(float s) { return 1.f / (1.f + expf(-s)); }
// fill_triangular (TFP 0.7, lower): element (i, j), j <= i, of the L x L matrix comes from entry idx of the
// vector of T = L (L + 1) / 2 scales: reshape(concat(x[L:], reverse(x)), (L, L)).
__device__ __forceinline__ int tril_index(int i, int j, int L, int T) {
    const int f = i * L + j;
    return f < T - L ? L + f : 2 * T - L - 1 - f;
}
// activated scale: clip(softplus(raw), tiny, inf); its derivative w.r.t. raw
__device__ __forceinline__ float scale_of(float raw) { return fmaxf(fc_softplus(raw), kTiny); }
__device__ __forceinline__ float dscale_of(float raw) { return fc_softplus(raw) > kTiny ? fc_sigmoid(raw) : 0.f; }

// pl[k][i][j] = activated scale matrix of p(z|y=k) (dense L x L, zero above the diagonal)
__global__ void full_prior_kernel(const float *__restrict__ pz, int64_t ldp, int K, int L, float *__restrict__ pl) {
    const int T = L * (L + 1) / 2;
    const int k = blockIdx.x;
    for (int e = threadIdx.x; e < L * L; e += blockDim.x) {
        const int i = e / L, j = e - i * L;
        pl[(int64_t)k * L * L + e] = j <= i ? scale_of(pz[(int64_t)k * ldp + L + tril_index(i, j, L, T)]) : 0.f;
    }
}

// Forward: one warp per row (k, rs, b).  w (rows, L) keeps S_p^-1 (z - loc_p) for the backward.
__global__ void __launch_bounds__(128)
full_latent_fwd_kernel(const float *__restrict__ qh, int64_t ldq, const float *__restrict__ pz, int64_t ldp,
                       const float *__restrict__ pl, int K, int B, int L, int RS, const float *__restrict__ eps,
                       float *__restrict__ z, int64_t ldz, float *__restrict__ klz, float *__restrict__ w_out) {
    __shared__ float s_e[4][kFullMaxL], s_w[4][kFullMaxL];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * 4 + wp;
    if (row >= (int64_t)K * RS * B) return;
    const int T = L * (L + 1) / 2;
    const int b = (int)(row % B);
    const int k = (int)(row / ((int64_t)RS * B));
    const float *q = qh + ((int64_t)k * B + b) * ldq;
    const float *ploc = pz + (int64_t)k * ldp;
    const float *P = pl + (int64_t)k * L * L;
    float *e = s_e[wp], *w = s_w[wp];
    float ee = 0.f;
    for (int l = lane; l < L; l += 32) {
        const float v = eps[row * L + l];
        e[l] = v;
        ee = fmaf(v, v, ee);
    }
    __syncwarp();
    float log_dq = 0.f, log_dp = 0.f, ww = 0.f;
    // z_i = loc_i + sum_{j <= i} S_q[i][j] eps_j, then w_i = (z_i - loc_p_i - sum_{j < i} S_p[i][j] w_j) / S_p[i][i]
    for (int i = 0; i < L; ++i) {
        float acc = 0.f, sol = 0.f;
        for (int j = lane; j <= i; j += 32) {
            const float s = scale_of(q[L + tril_index(i, j, L, T)]);
            acc = fmaf(s, e[j], acc);
            if (j == i) log_dq += logf(s);
            else sol = fmaf(P[i * L + j], w[j], sol);
        }
        acc = warp_sum(acc);
        sol = warp_sum(sol);
        const float zi = q[i] + acc;
        const float pii = P[i * L + i];
        const float wi = (zi - ploc[i] - sol) / pii;
        if (lane == 0) {
            z[row * ldz + i] = zi;
            w[i] = wi;
            w_out[row * L + i] = wi;
            log_dp += logf(pii);
            ww = fmaf(wi, wi, ww);
        }
        __syncwarp();
    }
    for (int c = L + lane; c < ldz; c += 32) z[row * ldz + c] = (c == L) ? 1.f : 0.f;
    ee = warp_sum(ee);
    log_dq = warp_sum(log_dq);
    // KL = log q - log p = (-1/2 |eps|^2 - sum log S_q,ii) - (-1/2 |w|^2 - sum log S_p,ii)
    if (lane == 0) klz[row] = (-0.5f * ee - log_dq) - (-0.5f * ww - log_dp);
}

// Backward w.r.t. the q(z|x,y) head pre-activations: one warp per (k, b), looping its RS rows.
// Per row: solve S_p^T u = w; g = dz + coef u (total d loss / d z); d loc_q += g;
// d S_q[i][j] += g_i eps_j - [i == j] coef / S_q,ii; cu (rows, L) = coef u for the prior's gradient.
__global__ void __launch_bounds__(128)
full_latent_bwd_q_kernel(const float *__restrict__ qh, int64_t ldq, const float *__restrict__ pl, int K, int B, int L,
                         int RS, const float *__restrict__ eps, const float *__restrict__ dz, int64_t lddz,
                         const float *__restrict__ coef, const float *__restrict__ w_in, float *__restrict__ cu,
                         float *__restrict__ dqh, int64_t lddq) {
    __shared__ float s_e[4][kFullMaxL], s_g[4][kFullMaxL], s_u[4][kFullMaxL];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int64_t kb = (int64_t)blockIdx.x * 4 + wp;
    if (kb >= (int64_t)K * B) return;
    const int T = L * (L + 1) / 2;
    const int b = (int)(kb % B), k = (int)(kb / B);
    const float *q = qh + kb * ldq;
    const float *P = pl + (int64_t)k * L * L;
    float *e = s_e[wp], *g = s_g[wp], *u = s_u[wp];
    float *o = dqh + kb * lddq;
    for (int rs = 0; rs < RS; ++rs) {
        const int64_t row = ((int64_t)k * RS + rs) * B + b;
        const float c = coef[row];
        for (int l = lane; l < L; l += 32) e[l] = eps[row * L + l];
        __syncwarp();
        // backward substitution: u_i = (w_i - sum_{j > i} S_p[j][i] u_j) / S_p[i][i]
        for (int i = L - 1; i >= 0; --i) {
            float sol = 0.f;
            for (int j = i + 1 + lane; j < L; j += 32) sol = fmaf(P[j * L + i], u[j], sol);
            sol = warp_sum(sol);
            if (lane == 0) u[i] = (w_in[row * L + i] - sol) / P[i * L + i];
            __syncwarp();
        }
        for (int l = lane; l < L; l += 32) {
            const float gz = fmaf(c, u[l], dz[row * lddz + l]);
            g[l] = gz;
            cu[row * L + l] = c * u[l];
            o[l] = rs == 0 ? gz : o[l] + gz;
        }
        __syncwarp();
        for (int t = lane; t < L * L; t += 32) {
            const int i = t / L, j = t - i * L;
            if (j > i) continue;
            const int idx = tril_index(i, j, L, T);
            const float raw = q[L + idx];
            float d = g[i] * e[j];
            if (i == j) d -= c / scale_of(raw);
            d *= dscale_of(raw);
            o[L + idx] = rs == 0 ? d : o[L + idx] + d;
        }
        __syncwarp();
    }
}

// Gradient w.r.t. the prior's pre-activations pz (K, ldp) = [locations | scales]: one CTA per cluster,
// every entry a fixed-order sum over the cluster's rows (deterministic).
//   d loc_p = - sum_rows coef u;   d S_p[i][j] = [i == j] (sum_rows coef) / S_p,ii - sum_rows (coef u_i) w_j
__global__ void __launch_bounds__(256)
full_latent_bwd_p_kernel(const float *__restrict__ pz, int64_t ldp, const float *__restrict__ pl, int K, int B, int L,
                         int RS, const float *__restrict__ coef, const float *__restrict__ w_in,
                         const float *__restrict__ cu, float *__restrict__ dpz, int64_t lddp) {
    const int T = L * (L + 1) / 2;
    const int k = blockIdx.x;
    const int64_t r0 = (int64_t)k * RS * B, r1 = r0 + (int64_t)RS * B;
    const float *P = pl + (int64_t)k * L * L;
    __shared__ float s_c;
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int64_t r = r0; r < r1; ++r) s += coef[r];
        s_c = s;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < L * L + L; t += blockDim.x) {
        if (t >= L * L) {                 // locations
            const int i = t - L * L;
            float s = 0.f;
            for (int64_t r = r0; r < r1; ++r) s += cu[r * L + i];
            dpz[(int64_t)k * lddp + i] = -s;
            continue;
        }
        const int i = t / L, j = t - i * L;
        if (j > i) continue;
        float s = 0.f;
        for (int64_t r = r0; r < r1; ++r) s = fmaf(cu[r * L + i], w_in[r * L + j], s);
        float d = -s;
        if (i == j) d += s_c / P[i * L + i];
        const int idx = tril_index(i, j, L, T);
        dpz[(int64_t)k * lddp + L + idx] = d * dscale_of(pz[(int64_t)k * ldp + L + idx]);
    }
}

// Evaluation statistics of the posterior (GMVAE:2881-2893): per cluster, the mean over the cells of
// diag(S S^T) (q_z_variances) and of S S^T (q_z_covariances).  One CTA per (cluster, matrix row i).
__global__ void __launch_bounds__(128)
full_covariance_mean_kernel(const float *__restrict__ qh, int64_t ldq, int K, int B, int L, float *__restrict__ cov) {
    const int T = L * (L + 1) / 2;
    const int k = blockIdx.x, i = blockIdx.y;
    for (int j = threadIdx.x; j < L; j += blockDim.x) {
        const int m = min(i, j);
        float acc = 0.f;
        for (int b = 0; b < B; ++b) {
            const float *q = qh + ((int64_t)k * B + b) * ldq + L;
            float s = 0.f;
            for (int t = 0; t <= m; ++t) s = fmaf(scale_of(q[tril_index(i, t, L, T)]), scale_of(q[tril_index(j, t, L, T)]), s);
            acc += s;
        }
        cov[((int64_t)k * L + i) * L + j] = acc / (float)B;
    }
}

}  // namespace scvae

using namespace scvae;

extern "C" int scvae_gmvae_full_prior(const float *pz, int64_t ldp, int K, int L, float *pl, void *stream) {
    SCVAE_CHECK_ARG(pz && pl && K > 0 && L > 0 && L <= kFullMaxL && ldp >= L + L * (L + 1) / 2,
                    "gmvae_full_prior: bad arguments (latent size <= %d)", kFullMaxL);
    full_prior_kernel<<<K, 256, 0, (cudaStream_t)stream>>>(pz, ldp, K, L, pl);
    SCVAE_CHECK_LAUNCH("gmvae_full_prior");
    return 0;
}

extern "C" int scvae_gmvae_latent_full_fwd(const float *qh, int64_t ldq, const float *pz, int64_t ldp, const float *pl,
                                           int K, int B, int L, int RS, const float *eps, float *z, int64_t ldz,
                                           float *klz, float *w, void *stream) {
    SCVAE_CHECK_ARG(qh && pz && pl && eps && z && klz && w && K > 0 && B > 0 && RS > 0 && L > 0 && L <= kFullMaxL &&
                        ldz > L,
                    "gmvae_latent_full_fwd: bad arguments (latent size <= %d)", kFullMaxL);
    const int64_t rows = (int64_t)K * RS * B;
    full_latent_fwd_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, (cudaStream_t)stream>>>(qh, ldq, pz, ldp, pl, K, B, L,
                                                                                         RS, eps, z, ldz, klz, w);
    SCVAE_CHECK_LAUNCH("gmvae_latent_full_fwd");
    return 0;
}

extern "C" int scvae_gmvae_latent_full_bwd(const float *qh, int64_t ldq, const float *pz, int64_t ldp, const float *pl,
                                           int K, int B, int L, int RS, const float *eps, const float *dz,
                                           int64_t lddz, const float *coef, const float *w, float *cu, float *dqh,
                                           int64_t lddq, float *dpz, int64_t lddp, void *stream) {
    SCVAE_CHECK_ARG(qh && pz && pl && eps && dz && coef && w && cu && dqh && dpz && L > 0 && L <= kFullMaxL,
                    "gmvae_latent_full_bwd: bad arguments (latent size <= %d)", kFullMaxL);
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t kb = (int64_t)K * B;
    full_latent_bwd_q_kernel<<<(unsigned)((kb + 3) / 4), 128, 0, s>>>(qh, ldq, pl, K, B, L, RS, eps, dz, lddz, coef, w, cu,
                                                                     dqh, lddq);
    SCVAE_CHECK_LAUNCH("gmvae_latent_full_bwd_q");
    full_latent_bwd_p_kernel<<<K, 256, 0, s>>>(pz, ldp, pl, K, B, L, RS, coef, w, cu, dpz, lddp);
    SCVAE_CHECK_LAUNCH("gmvae_latent_full_bwd_p");
    return 0;
}

extern "C" int scvae_gmvae_full_covariance_mean(const float *qh, int64_t ldq, int K, int B, int L, float *cov,
                                                void *stream) {
    SCVAE_CHECK_ARG(qh && cov && K > 0 && B > 0 && L > 0 && L <= kFullMaxL, "gmvae_full_covariance_mean: bad arguments");
    full_covariance_mean_kernel<<<dim3(K, L), 128, 0, (cudaStream_t)stream>>>(qh, ldq, K, B, L, cov);
    SCVAE_CHECK_LAUNCH("gmvae_full_covariance_mean");
    return 0;
}
