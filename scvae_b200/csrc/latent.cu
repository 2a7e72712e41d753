// a3/a6: Gaussian posterior q(z|x) = N(mu, exp(log_sigma)), reparameterised sample and the
// analytic KL to N(0,1) (VAE:2243-2369, VAE:2624-2656, DU:31-50), plus the VAE bound
// (VAE:2715-2734, MU:129-137).  One warp per cell; everything here is O(B*L) and tiny next to
// the (cells x genes) kernels, so the goal is few launches, not bandwidth.
#include <curand_kernel.h>

#include "common.cuh"

namespace scvae {

constexpr int kRowsPerBlock = 4;  // 4 warps, one cell each

__global__ void __launch_bounds__(32 * kRowsPerBlock)
gaussian_latent_fwd_kernel(const float *__restrict__ ph, int64_t ldph, int B, int L, int RS,
                           const float *__restrict__ eps, int unit_variance, int deterministic,
                           float *__restrict__ z, int64_t ldz, float *__restrict__ kl_row,
                           float *__restrict__ kl_elem) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
    if (b >= B) return;
    const float *pr = ph + (int64_t)b * ldph;
    float kl = 0.f;
    for (int l = lane; l < L; l += 32) {
        const float mu = pr[l];
        const float ls = unit_variance ? 0.f : fminf(fmaxf(pr[L + l], -3.f), 3.f);
        const float sigma = __expf(ls);
        // tfp kl_divergence(Normal(mu, sigma), Normal(0, 1))
        const float k = 0.5f * mu * mu + 0.5f * (sigma * sigma - 1.f) - ls;
        kl += k;
        if (kl_elem) kl_elem[(int64_t)b * L + l] = k;
        if (deterministic) {
            z[(int64_t)b * ldz + l] = mu;
        } else {
            for (int s = 0; s < RS; ++s) {
                const int64_t m = (int64_t)s * B + b;
                z[m * ldz + l] = mu + sigma * eps[m * L + l];
            }
        }
    }
    // augmented columns: z[:, L] = 1, z[:, L+1:] = 0
    const int nrep = deterministic ? 1 : RS;
    for (int s = 0; s < nrep; ++s) {
        float *zr = z + ((int64_t)s * B + b) * ldz;
        for (int c = L + lane; c < ldz; c += 32) zr[c] = (c == L) ? 1.f : 0.f;
    }
    kl = warp_sum(kl);
    if (lane == 0 && kl_row) kl_row[b] = kl;
}

__global__ void __launch_bounds__(32 * kRowsPerBlock)
gaussian_latent_bwd_kernel(const float *__restrict__ ph, int64_t ldph, int B, int L, int RS,
                           const float *__restrict__ eps, int unit_variance,
                           const float *__restrict__ dz, int64_t lddz, float kl_coef,
                           float *__restrict__ dph, int64_t lddph) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
    if (b >= B) return;
    const float *pr = ph + (int64_t)b * ldph;
    float *dr = dph + (int64_t)b * lddph;
    for (int l = lane; l < L; l += 32) {
        const float mu = pr[l];
        const float raw = unit_variance ? 0.f : pr[L + l];
        const float ls = fminf(fmaxf(raw, -3.f), 3.f);
        const float sigma = __expf(ls);
        float dmu = 0.f, dse = 0.f;
        for (int s = 0; s < RS; ++s) {
            const int64_t m = (int64_t)s * B + b;
            const float d = dz[m * lddz + l];
            dmu += d;
            dse += d * eps[m * L + l];
        }
        dr[l] = dmu + kl_coef * mu;
        if (!unit_variance) {
            const float mask = (raw < -3.f || raw > 3.f) ? 0.f : 1.f;
            dr[L + l] = (dse * sigma + kl_coef * (sigma * sigma - 1.f)) * mask;
        }
    }
}

// lower_bound = mean_{s,b} logmeanexp_r(logp - kl); weighted variant; go = -softmax_r / (S B).
__global__ void __launch_bounds__(1024)
vae_bound_kernel(const float *__restrict__ logp, const float *__restrict__ kl_row, int R, int S,
                 int B, float weight, float *__restrict__ out, float *__restrict__ go) {
    __shared__ float red[32];
    const int SB = S * B;
    const float inv_sb = 1.f / (float)SB;
    float s_lb = 0.f, s_lbw = 0.f, s_lp = 0.f, s_kl = 0.f;
    for (int i = threadIdx.x; i < SB; i += blockDim.x) {
        const int b = i % B;
        const float kl = kl_row[b];
        float mx = -INFINITY, mxw = -INFINITY;
        for (int r = 0; r < R; ++r) {
            const float lp = logp[(int64_t)r * SB + i];
            mx = fmaxf(mx, lp - kl);
            mxw = fmaxf(mxw, lp - weight * kl);
            s_lp += lp;
        }
        float se = 0.f, sew = 0.f;
        for (int r = 0; r < R; ++r) {
            const float lp = logp[(int64_t)r * SB + i];
            se += expf(lp - kl - mx);
            sew += expf(lp - weight * kl - mxw);
        }
        s_lb += logf(se / (float)R) + mx;
        s_lbw += logf(sew / (float)R) + mxw;
        if (i < B) s_kl += kl;
        if (go) {
            for (int r = 0; r < R; ++r) {
                const float lp = logp[(int64_t)r * SB + i];
                go[(int64_t)r * SB + i] = -expf(lp - weight * kl - mxw) / sew * inv_sb;
            }
        }
    }
    const float lb = block_sum(s_lb, red);
    const float lbw = block_sum(s_lbw, red);
    const float lp = block_sum(s_lp, red);
    const float kl = block_sum(s_kl, red);
    if (threadIdx.x == 0) {
        out[0] = lb * inv_sb;
        out[1] = lbw * inv_sb;
        out[2] = lp * inv_sb / (float)R;
        out[3] = kl / (float)B;
    }
}

// ---- sampled (non-analytical) KL term (VAE:2628-2640): kl[r,s,b,l] = log q(z|x) - log p(z) at
// the drawn z = mu + sigma eps, i.e. 0.5 z^2 - 0.5 eps^2 - log_sigma (the 2 pi terms cancel).
// The default of every VAE latent distribution except the plain `gaussian` (VAE:186-192).
// One warp per cell.  kl_rows[RS*B] (sample-major, like logp) = sum_l; kl_elem (nullable, (B, L))
// = mean over the RS samples, so that its column mean is kl_divergence_neurons (VAE:2643-2646).
__global__ void __launch_bounds__(32 * kRowsPerBlock)
gaussian_sampled_kl_kernel(const float *__restrict__ ph, int64_t ldph, int B, int L, int RS,
                           const float *__restrict__ eps, int unit_variance, int deterministic,
                           float *__restrict__ kl_rows, float *__restrict__ kl_elem) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
    if (b >= B) return;
    const float *pr = ph + (int64_t)b * ldph;
    const int nrep = deterministic ? 1 : RS;
    for (int s = 0; s < nrep; ++s) {
        const int64_t m = (int64_t)s * B + b;
        float acc = 0.f;
        for (int l = lane; l < L; l += 32) {
            const float mu = pr[l];
            const float ls = unit_variance ? 0.f : fminf(fmaxf(pr[L + l], -3.f), 3.f);
            const float e = deterministic ? 0.f : eps[m * L + l];
            const float zv = mu + __expf(ls) * e;
            acc += 0.5f * zv * zv - 0.5f * e * e - ls;
        }
        acc = warp_sum(acc);
        if (lane == 0) kl_rows[m] = acc;
    }
    if (kl_elem) {
        const float inv = 1.f / (float)nrep;
        for (int l = lane; l < L; l += 32) {
            const float mu = pr[l];
            const float ls = unit_variance ? 0.f : fminf(fmaxf(pr[L + l], -3.f), 3.f);
            const float sigma = __expf(ls);
            float acc = 0.f;
            for (int s = 0; s < nrep; ++s) {
                const float e = deterministic ? 0.f : eps[((int64_t)s * B + b) * L + l];
                const float zv = mu + sigma * e;
                acc += 0.5f * zv * zv - 0.5f * e * e - ls;
            }
            kl_elem[(int64_t)b * L + l] = acc * inv;
        }
    }
}

// Backward of the reparameterised sample AND the sampled KL: with c_m = d loss / d kl_rows[m]
// (= -weight * go[m]; weight / (S B) when R == 1), z = mu + sigma eps:
//   d kl_m / d mu = z,   d kl_m / d log_sigma = z sigma eps - 1.
__global__ void __launch_bounds__(32 * kRowsPerBlock)
gaussian_sampled_kl_bwd_kernel(const float *__restrict__ ph, int64_t ldph, int B, int L, int RS,
                               const float *__restrict__ eps, int unit_variance,
                               const float *__restrict__ dz, int64_t lddz,
                               const float *__restrict__ go, float weight, float coef_scalar,
                               float *__restrict__ dph, int64_t lddph) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
    if (b >= B) return;
    const float *pr = ph + (int64_t)b * ldph;
    float *dr = dph + (int64_t)b * lddph;
    for (int l = lane; l < L; l += 32) {
        const float mu = pr[l];
        const float raw = unit_variance ? 0.f : pr[L + l];
        const float ls = fminf(fmaxf(raw, -3.f), 3.f);
        const float sigma = __expf(ls);
        float dmu = 0.f, dls = 0.f;
        for (int s = 0; s < RS; ++s) {
            const int64_t m = (int64_t)s * B + b;
            const float e = eps[m * L + l];
            const float c = go ? -weight * go[m] : coef_scalar;
            const float zv = mu + sigma * e;
            const float d = dz[m * lddz + l] + c * zv;      // total gradient w.r.t. z
            dmu += d;
            dls += d * sigma * e - c;
        }
        dr[l] = dmu;
        if (!unit_variance) {
            const float mask = (raw < -3.f || raw > 3.f) ? 0.f : 1.f;
            dr[L + l] = dls * mask;
        }
    }
}

// vae_bound_kernel with one KL value per (r, s, b) row (sampled KL).  out[3] = mean over all
// rows = sum_l kl_divergence_neurons[l] (VAE:2643-2652).
__global__ void __launch_bounds__(1024)
vae_bound_rows_kernel(const float *__restrict__ logp, const float *__restrict__ kl_rows, int R,
                      int S, int B, float weight, float *__restrict__ out,
                      float *__restrict__ go) {
    __shared__ float red[32];
    const int SB = S * B;
    const float inv_sb = 1.f / (float)SB;
    float s_lb = 0.f, s_lbw = 0.f, s_lp = 0.f, s_kl = 0.f;
    for (int i = threadIdx.x; i < SB; i += blockDim.x) {
        float mx = -INFINITY, mxw = -INFINITY;
        for (int r = 0; r < R; ++r) {
            const float lp = logp[(int64_t)r * SB + i];
            const float kl = kl_rows[(int64_t)r * SB + i];
            mx = fmaxf(mx, lp - kl);
            mxw = fmaxf(mxw, lp - weight * kl);
            s_lp += lp;
            s_kl += kl;
        }
        float se = 0.f, sew = 0.f;
        for (int r = 0; r < R; ++r) {
            const float lp = logp[(int64_t)r * SB + i];
            const float kl = kl_rows[(int64_t)r * SB + i];
            se += expf(lp - kl - mx);
            sew += expf(lp - weight * kl - mxw);
        }
        s_lb += logf(se / (float)R) + mx;
        s_lbw += logf(sew / (float)R) + mxw;
        if (go) {
            for (int r = 0; r < R; ++r) {
                const float lp = logp[(int64_t)r * SB + i];
                const float kl = kl_rows[(int64_t)r * SB + i];
                go[(int64_t)r * SB + i] = -expf(lp - weight * kl - mxw) / sew * inv_sb;
            }
        }
    }
    const float lb = block_sum(s_lb, red);
    const float lbw = block_sum(s_lbw, red);
    const float lp = block_sum(s_lp, red);
    const float kl = block_sum(s_kl, red);
    if (threadIdx.x == 0) {
        out[0] = lb * inv_sb;
        out[1] = lbw * inv_sb;
        out[2] = lp * inv_sb / (float)R;
        out[3] = kl * inv_sb / (float)R;
    }
}

__global__ void col_mean_kernel(const float *__restrict__ x, int64_t ldx, int rows, int cols,
                                float *__restrict__ out) {
    // one block of 32 x 8 threads per 32 columns; deterministic tree
    __shared__ float tile[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    float acc = 0.f;
    if (c < cols)
        for (int r = threadIdx.y; r < rows; r += 8) acc += x[(int64_t)r * ldx + c];
    tile[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += tile[i][threadIdx.x];
        out[c] = t / (float)rows;
    }
}

__global__ void fill_normal_kernel(float *__restrict__ out, int64_t n, uint64_t seed, uint64_t offset,
                                   const int64_t *__restrict__ offset_dev) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t base = tid * 4;
    if (base >= n) return;
    curandStatePhilox4_32_10_t st;
    if (offset_dev) offset += (uint64_t)(*offset_dev);
    // curand offsets count single 32-bit outputs and one call consumes a Philox block of four:
    // four outputs per unit of `offset`, so that consecutive offsets (steps, minibatches) draw
    // from disjoint blocks instead of overlapping ones
    curand_init(seed, (unsigned long long)tid, 4ull * offset, &st);
    const float4 v = curand_normal4(&st);
    const float vv[4] = {v.x, v.y, v.z, v.w};
    for (int j = 0; j < 4 && base + j < n; ++j) out[base + j] = vv[j];
}

}  // namespace scvae

namespace scvae {
// Decoder-input extras (VAE:2400-2441): one-hot batch index and/or normalised count sum of cell
// b = m % B, written behind the augmented ones column of the latent sample matrix.
__global__ void decoder_features_kernel(float *__restrict__ z, int64_t ldz, int M, int B, int col0,
                                        const float *__restrict__ batch_index, int n_batches,
                                        const float *__restrict__ count_sum) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int b = m % B;
    float *zr = z + (int64_t)m * ldz + col0;
    if (batch_index) {
        const int k = (int)batch_index[b];
        for (int c = 0; c < n_batches; ++c) zr[c] = (c == k) ? 1.f : 0.f;
    }
    if (count_sum) zr[batch_index ? n_batches : 0] = count_sum[b];
}
}  // namespace scvae

using namespace scvae;

extern "C" int scvae_decoder_features(float *z, int64_t ldz, int M, int B, int col0, const float *batch_index,
                                      int n_batches, const float *count_sum, void *stream) {
    SCVAE_CHECK_ARG(z && M > 0 && B > 0 && col0 >= 0, "decoder_features: bad arguments");
    SCVAE_CHECK_ARG((!batch_index || n_batches > 0) && col0 + (batch_index ? n_batches : 0) + (count_sum ? 1 : 0) <= ldz,
                    "decoder_features: the extra columns do not fit the row (ldz=%lld)", (long long)ldz);
    if (!batch_index && !count_sum) return 0;
    decoder_features_kernel<<<(M + 255) / 256, 256, 0, (cudaStream_t)stream>>>(z, ldz, M, B, col0, batch_index,
                                                                             n_batches, count_sum);
    SCVAE_CHECK_LAUNCH("decoder_features");
    return 0;
}

extern "C" int scvae_gaussian_latent_fwd(const float *ph, int64_t ldph, int B, int L, int RS,
                                         const float *eps, int unit_variance, int deterministic,
                                         float *z, int64_t ldz, float *kl_row, float *kl_elem,
                                         void *stream) {
    SCVAE_CHECK_ARG(ph && z && B > 0 && L > 0 && RS > 0, "gaussian_latent_fwd: bad arguments");
    SCVAE_CHECK_ARG(deterministic || eps, "gaussian_latent_fwd: eps is NULL");
    SCVAE_CHECK_ARG(ldz >= L && ldph >= (unit_variance ? L : 2 * L), "gaussian_latent_fwd: bad ld");
    const int blocks = (B + kRowsPerBlock - 1) / kRowsPerBlock;
    gaussian_latent_fwd_kernel<<<blocks, 32 * kRowsPerBlock, 0, (cudaStream_t)stream>>>(
        ph, ldph, B, L, RS, eps, unit_variance, deterministic, z, ldz, kl_row, kl_elem);
    SCVAE_CHECK_LAUNCH("gaussian_latent_fwd");
    return 0;
}

extern "C" int scvae_gaussian_latent_bwd(const float *ph, int64_t ldph, int B, int L, int RS,
                                         const float *eps, int unit_variance, const float *dz,
                                         int64_t lddz, float kl_coef, float *dph, int64_t lddph,
                                         void *stream) {
    SCVAE_CHECK_ARG(ph && eps && dz && dph && B > 0 && L > 0 && RS > 0,
                    "gaussian_latent_bwd: bad arguments");
    const int blocks = (B + kRowsPerBlock - 1) / kRowsPerBlock;
    gaussian_latent_bwd_kernel<<<blocks, 32 * kRowsPerBlock, 0, (cudaStream_t)stream>>>(
        ph, ldph, B, L, RS, eps, unit_variance, dz, lddz, kl_coef, dph, lddph);
    SCVAE_CHECK_LAUNCH("gaussian_latent_bwd");
    return 0;
}

extern "C" int scvae_vae_bound(const float *logp, const float *kl_row, int R, int S, int B,
                               float weight, float *out, float *go, void *stream) {
    SCVAE_CHECK_ARG(logp && kl_row && out && R > 0 && S > 0 && B > 0, "vae_bound: bad arguments");
    // one CTA (deterministic tree); 1024 threads keep the serial depth at S*B/1024 terms
    const int threads = (S * B >= 1024) ? 1024 : 256;
    vae_bound_kernel<<<1, threads, 0, (cudaStream_t)stream>>>(logp, kl_row, R, S, B, weight, out, go);
    SCVAE_CHECK_LAUNCH("vae_bound");
    return 0;
}

extern "C" int scvae_gaussian_sampled_kl(const float *ph, int64_t ldph, int B, int L, int RS,
                                         const float *eps, int unit_variance, int deterministic,
                                         float *kl_rows, float *kl_elem, void *stream) {
    SCVAE_CHECK_ARG(ph && kl_rows && B > 0 && L > 0 && RS > 0, "gaussian_sampled_kl: bad arguments");
    SCVAE_CHECK_ARG(deterministic || eps, "gaussian_sampled_kl: eps is NULL");
    SCVAE_CHECK_ARG(ldph >= (unit_variance ? L : 2 * L), "gaussian_sampled_kl: bad ld");
    const int blocks = (B + kRowsPerBlock - 1) / kRowsPerBlock;
    gaussian_sampled_kl_kernel<<<blocks, 32 * kRowsPerBlock, 0, (cudaStream_t)stream>>>(
        ph, ldph, B, L, RS, eps, unit_variance, deterministic, kl_rows, kl_elem);
    SCVAE_CHECK_LAUNCH("gaussian_sampled_kl");
    return 0;
}

extern "C" int scvae_gaussian_sampled_kl_bwd(const float *ph, int64_t ldph, int B, int L, int RS,
                                             const float *eps, int unit_variance, const float *dz,
                                             int64_t lddz, const float *go, float weight,
                                             float coef_scalar, float *dph, int64_t lddph,
                                             void *stream) {
    SCVAE_CHECK_ARG(ph && eps && dz && dph && B > 0 && L > 0 && RS > 0,
                    "gaussian_sampled_kl_bwd: bad arguments");
    const int blocks = (B + kRowsPerBlock - 1) / kRowsPerBlock;
    gaussian_sampled_kl_bwd_kernel<<<blocks, 32 * kRowsPerBlock, 0, (cudaStream_t)stream>>>(
        ph, ldph, B, L, RS, eps, unit_variance, dz, lddz, go, weight, coef_scalar, dph, lddph);
    SCVAE_CHECK_LAUNCH("gaussian_sampled_kl_bwd");
    return 0;
}

extern "C" int scvae_vae_bound_rows(const float *logp, const float *kl_rows, int R, int S, int B,
                                    float weight, float *out, float *go, void *stream) {
    SCVAE_CHECK_ARG(logp && kl_rows && out && R > 0 && S > 0 && B > 0,
                    "vae_bound_rows: bad arguments");
    const int threads = (S * B >= 1024) ? 1024 : 256;
    vae_bound_rows_kernel<<<1, threads, 0, (cudaStream_t)stream>>>(logp, kl_rows, R, S, B, weight,
                                                                   out, go);
    SCVAE_CHECK_LAUNCH("vae_bound_rows");
    return 0;
}

extern "C" int scvae_col_mean(const float *x, int64_t ldx, int rows, int cols, float *out,
                              void *stream) {
    SCVAE_CHECK_ARG(x && out && rows > 0 && cols > 0, "col_mean: bad arguments");
    col_mean_kernel<<<(cols + 31) / 32, dim3(32, 8), 0, (cudaStream_t)stream>>>(x, ldx, rows, cols, out);
    SCVAE_CHECK_LAUNCH("col_mean");
    return 0;
}

extern "C" int scvae_fill_normal(float *out, int64_t n, uint64_t seed, uint64_t offset,
                                 const int64_t *offset_dev, void *stream) {
    SCVAE_CHECK_ARG(out && n >= 0, "fill_normal: bad arguments");
    if (n == 0) return 0;
    const int64_t threads = (n + 3) / 4;
    fill_normal_kernel<<<(int)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out, n, seed, offset, offset_dev);
    SCVAE_CHECK_LAUNCH("fill_normal");
    return 0;
}
