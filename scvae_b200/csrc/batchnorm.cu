// a2: batch normalisation (center only) + ReLU around the dense layers (MU:62-74).
//
// tf.contrib.layers.batch_norm(center=True, scale=False), epsilon 1e-3, decay 0.999:
//   train: yhat = (y - mean_B) / sqrt(var_B + eps) + beta   (biased batch variance),
//          moving <- moving - (1 - decay)(moving - batch) with the Bessel-corrected variance;
//   eval : moving statistics.
// The statistics couple all rows of a minibatch, so the op is two tiny launches: per-row-chunk
// partial moments (count, mean, M2) and a finalise+apply pass in which every CTA re-combines the
// partials in a fixed order (Chan's formula: deterministic, no atomics, no cancellation).  Rows
// may form `groups` independent groups (one BN op per GMVAE cluster k, GMVAE:2859-2877).
// The output is written in the augmented layout (column H = 1, columns > H = 0).
#include "common.cuh"

namespace scvae {

constexpr int kBnRows = 64;   // rows per chunk
constexpr int kBnCols = 32;   // columns per CTA
constexpr int kBnLanes = 8;   // row lanes per CTA (blockDim = 32 x 8)

static inline int bn_chunks(int rows_per_group) { return (rows_per_group + kBnRows - 1) / kBnRows; }
static inline int round32(int h) { return (h + 31) & ~31; }

// scratch layout: part[3][groups*nchunks][H32] then var_unbiased[groups][H32]
struct BnScratch {
    float *p0, *p1, *p2, *var;
    int H32;
};
static inline BnScratch bn_scratch(float *scratch, int M, int H, int groups) {
    const int H32 = round32(H);
    const int64_t n = (int64_t)groups * bn_chunks(M / groups) * H32;
    return {scratch, scratch + n, scratch + 2 * n, scratch + 3 * n, H32};
}

// ---- forward -------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBnCols *kBnLanes)
bn_partial_kernel(const float *__restrict__ y, int64_t ldy, int n, int H, int nchunks, int H32,
                  float *__restrict__ pmean, float *__restrict__ pm2) {
    __shared__ float sh[kBnLanes][kBnCols + 1];
    const int c = blockIdx.y * kBnCols + threadIdx.x;
    const int chunk = blockIdx.x % nchunks, group = blockIdx.x / nchunks;
    const int r0 = chunk * kBnRows, r1 = min(r0 + kBnRows, n);
    const float *base = y + ((int64_t)group * n) * ldy;
    const bool ok = c < H;
    float s = 0.f;
    if (ok)
        for (int r = r0 + threadIdx.y; r < r1; r += kBnLanes) s += base[(int64_t)r * ldy + c];
    sh[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < kBnLanes; ++i) tot += sh[i][threadIdx.x];
    const float mean = tot / (float)(r1 - r0);
    __syncthreads();
    float q = 0.f;
    if (ok)
        for (int r = r0 + threadIdx.y; r < r1; r += kBnLanes) {
            const float d = base[(int64_t)r * ldy + c] - mean;
            q += d * d;
        }
    sh[threadIdx.y][threadIdx.x] = q;
    __syncthreads();
    if (threadIdx.y == 0 && ok) {
        float m2 = 0.f;
#pragma unroll
        for (int i = 0; i < kBnLanes; ++i) m2 += sh[i][threadIdx.x];
        const int64_t o = (int64_t)blockIdx.x * H32 + c;
        pmean[o] = mean;
        pm2[o] = m2;
    }
}

// Chan's pairwise update of (count, mean, M2) with another partial (nb, mb, qb).
__device__ __forceinline__ void chan_merge(float &cnt, float &mu, float &m2, float nb, float mb, float qb) {
    const float tot = cnt + nb;
    if (tot > 0.f) {
        const float delta = mb - mu;
        const float f = nb / tot;
        mu += delta * f;
        m2 += qb + delta * delta * (cnt * f);
        cnt = tot;
    }
}

// Combine the chunk partials of one group/column -> (mean, biased var).  The 8 row lanes of
// the CTA each fold every 8th partial, then lane 0 folds the 8 results: fixed order
// (deterministic), sequential depth nchunks/8 + 8 instead of nchunks.  All threads must call.
__device__ __forceinline__ void bn_combine(const float *pmean, const float *pm2, int group, int nchunks,
                                           int n, int H32, int c, bool ok, float (*sh)[3][kBnCols + 1],
                                           float &mean, float &var) {
    float cnt = 0.f, mu = 0.f, m2 = 0.f;
    if (ok) {
        for (int k = threadIdx.y; k < nchunks; k += kBnLanes) {
            const int64_t o = ((int64_t)group * nchunks + k) * H32 + c;
            const float nb = (float)(min((k + 1) * kBnRows, n) - k * kBnRows);
            chan_merge(cnt, mu, m2, nb, pmean[o], pm2[o]);
        }
    }
    sh[threadIdx.y][0][threadIdx.x] = cnt;
    sh[threadIdx.y][1][threadIdx.x] = mu;
    sh[threadIdx.y][2][threadIdx.x] = m2;
    __syncthreads();
    cnt = 0.f; mu = 0.f; m2 = 0.f;
#pragma unroll
    for (int i = 0; i < kBnLanes; ++i)
        chan_merge(cnt, mu, m2, sh[i][0][threadIdx.x], sh[i][1][threadIdx.x], sh[i][2][threadIdx.x]);
    mean = mu;
    var = cnt > 0.f ? m2 / cnt : 0.f;
}

__global__ void __launch_bounds__(kBnCols *kBnLanes)
bn_apply_kernel(const float *__restrict__ y, int64_t ldy, int n, int H, int nchunks, int H32,
                const float *__restrict__ pmean, const float *__restrict__ pm2,
                const float *__restrict__ beta, const float *__restrict__ moving_mean,
                const float *__restrict__ moving_var, int training, int relu,
                float *__restrict__ out, int64_t ldo, float *__restrict__ save_mean,
                float *__restrict__ save_rstd, float *__restrict__ var_unbiased, float *__restrict__ mov_mean_out,
                float *__restrict__ mov_var_out) {
    __shared__ float sh[kBnLanes][3][kBnCols + 1];
    const int c = blockIdx.y * kBnCols + threadIdx.x;
    const int chunk = blockIdx.x % nchunks, group = blockIdx.x / nchunks;
    const int r0 = chunk * kBnRows, r1 = min(r0 + kBnRows, n);
    const float *base = y + ((int64_t)group * n) * ldy;
    float *obase = out + ((int64_t)group * n) * ldo;
    const bool ok = c < H;
    float mean = 0.f, rstd = 1.f;
    if (training) {  // uniform branch: every thread of the CTA takes part in the combine
        float var;
        bn_combine(pmean, pm2, group, nchunks, n, H32, c, ok, sh, mean, var);
        rstd = rsqrtf(var + kBnEps);
        if (ok && chunk == 0 && threadIdx.y == 0) {
            save_mean[group * H + c] = mean;
            save_rstd[group * H + c] = rstd;
            const float vu = var * ((float)n / (float)max(n - 1, 1));
            var_unbiased[group * H32 + c] = vu;
            if (mov_mean_out) {   // single group: the moving-average update rides along
                mov_mean_out[c] -= (1.f - kBnDecay) * (mov_mean_out[c] - mean);
                mov_var_out[c] -= (1.f - kBnDecay) * (mov_var_out[c] - vu);
            }
        }
    } else if (ok) {
        mean = moving_mean[c];
        rstd = rsqrtf(moving_var[c] + kBnEps);
    }
    if (c >= ldo) return;
    if (!ok) {  // augmented columns
        const float v = (c == H) ? 1.f : 0.f;
        for (int r = r0 + threadIdx.y; r < r1; r += kBnLanes) obase[(int64_t)r * ldo + c] = v;
        return;
    }
    const float bt = beta[c];
    for (int r = r0 + threadIdx.y; r < r1; r += kBnLanes) {
        float v = (base[(int64_t)r * ldy + c] - mean) * rstd + bt;
        if (relu) v = fmaxf(v, 0.f);
        obase[(int64_t)r * ldo + c] = v;
    }
}

// moving <- moving - (1 - decay)(moving - batch), group after group (k = 0..groups-1).
__global__ void bn_moving_kernel(int H, int groups, int H32, const float *__restrict__ save_mean,
                                 const float *__restrict__ var_unbiased, float *__restrict__ moving_mean,
                                 float *__restrict__ moving_var) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= H) return;
    float mm = moving_mean[c], mv = moving_var[c];
    for (int k = 0; k < groups; ++k) {
        mm -= (1.f - kBnDecay) * (mm - save_mean[k * H + c]);
        mv -= (1.f - kBnDecay) * (mv - var_unbiased[k * H32 + c]);
    }
    moving_mean[c] = mm;
    moving_var[c] = mv;
}

// ---- backward ------------------------------------------------------------------------------
// dyhat = dout * relu'(out);  s1 = sum dyhat, s2 = sum dyhat * xhat  per group/column;
// dy = rstd (dyhat - s1/n - xhat s2/n);  dbeta = sum over everything of dyhat.
__global__ void __launch_bounds__(kBnCols *kBnLanes)
bn_bwd_partial_kernel(const float *__restrict__ dout, int64_t lddo, const float *__restrict__ y,
                      int64_t ldy, const float *__restrict__ out, int64_t ldo, int n, int H, int nchunks,
                      int H32, const float *__restrict__ save_mean, const float *__restrict__ save_rstd,
                      int relu, float *__restrict__ p1, float *__restrict__ p2) {
    __shared__ float sh1[kBnLanes][kBnCols + 1], sh2[kBnLanes][kBnCols + 1];
    const int c = blockIdx.y * kBnCols + threadIdx.x;
    const int chunk = blockIdx.x % nchunks, group = blockIdx.x / nchunks;
    const int r0 = chunk * kBnRows, r1 = min(r0 + kBnRows, n);
    const int64_t g0 = (int64_t)group * n;
    const bool ok = c < H;
    float s1 = 0.f, s2 = 0.f;
    if (ok) {
        const float mean = save_mean[group * H + c], rstd = save_rstd[group * H + c];
        for (int r = r0 + threadIdx.y; r < r1; r += kBnLanes) {
            const int64_t rr = g0 + r;
            float d = dout[rr * lddo + c];
            if (relu && !(out[rr * ldo + c] > 0.f)) d = 0.f;
            s1 += d;
            s2 += d * ((y[rr * ldy + c] - mean) * rstd);
        }
    }
    sh1[threadIdx.y][threadIdx.x] = s1;
    sh2[threadIdx.y][threadIdx.x] = s2;
    __syncthreads();
    if (threadIdx.y == 0 && ok) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int i = 0; i < kBnLanes; ++i) {
            a += sh1[i][threadIdx.x];
            b += sh2[i][threadIdx.x];
        }
        const int64_t o = (int64_t)blockIdx.x * H32 + c;
        p1[o] = a;
        p2[o] = b;
    }
}

__global__ void __launch_bounds__(kBnCols *kBnLanes)
bn_bwd_apply_kernel(const float *__restrict__ dout, int64_t lddo, const float *__restrict__ y,
                    int64_t ldy, const float *__restrict__ out, int64_t ldo, int n, int H, int groups,
                    int nchunks, int H32, const float *__restrict__ save_mean,
                    const float *__restrict__ save_rstd, int relu, const float *__restrict__ p1,
                    const float *__restrict__ p2, float *__restrict__ dy, int64_t lddy,
                    float *__restrict__ dbeta, int accumulate_dbeta) {
    const int c = blockIdx.y * kBnCols + threadIdx.x;
    const int chunk = blockIdx.x % nchunks, group = blockIdx.x / nchunks;
    const int r0 = chunk * kBnRows, r1 = min(r0 + kBnRows, n);
    const int64_t g0 = (int64_t)group * n;
    __shared__ float sh[kBnLanes][3][kBnCols + 1];
    const bool ok = c < H;
    float s1 = 0.f, s2 = 0.f, tot = 0.f;
    if (ok) {
        for (int k = threadIdx.y; k < nchunks; k += kBnLanes) {
            const int64_t o = ((int64_t)group * nchunks + k) * H32 + c;
            s1 += p1[o];
            s2 += p2[o];
        }
        if (blockIdx.x == 0)
            for (int k = threadIdx.y; k < groups * nchunks; k += kBnLanes) tot += p1[(int64_t)k * H32 + c];
    }
    sh[threadIdx.y][0][threadIdx.x] = s1;
    sh[threadIdx.y][1][threadIdx.x] = s2;
    sh[threadIdx.y][2][threadIdx.x] = tot;
    __syncthreads();
    if (!ok) return;
    s1 = 0.f; s2 = 0.f; tot = 0.f;
#pragma unroll
    for (int i = 0; i < kBnLanes; ++i) {
        s1 += sh[i][0][threadIdx.x];
        s2 += sh[i][1][threadIdx.x];
        tot += sh[i][2][threadIdx.x];
    }
    if (blockIdx.x == 0 && threadIdx.y == 0) dbeta[c] = accumulate_dbeta ? dbeta[c] + tot : tot;
    const float mean = save_mean[group * H + c], rstd = save_rstd[group * H + c];
    const float inv_n = 1.f / (float)n;
    const float m1 = s1 * inv_n, m2 = s2 * inv_n;
    for (int r = r0 + threadIdx.y; r < r1; r += kBnLanes) {
        const int64_t rr = g0 + r;
        float d = dout[rr * lddo + c];
        if (relu && !(out[rr * ldo + c] > 0.f)) d = 0.f;
        const float xh = (y[rr * ldy + c] - mean) * rstd;
        dy[rr * lddy + c] = rstd * (d - m1 - xh * m2);
    }
}

// ---- no-BN variants -------------------------------------------------------------------------
__global__ void act_fwd_kernel(const float *__restrict__ y, int64_t ldy, int M, int H, int relu,
                               float *__restrict__ out, int64_t ldo) {
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    const int r = blockIdx.x;
    if (c >= ldo) return;
    float v;
    if (c < H) {
        v = y[(int64_t)r * ldy + c];
        if (relu) v = fmaxf(v, 0.f);
    } else {
        v = (c == H) ? 1.f : 0.f;
    }
    out[(int64_t)r * ldo + c] = v;
}

__global__ void act_bwd_kernel(const float *__restrict__ dout, int64_t lddo, const float *__restrict__ out,
                               int64_t ldo, int M, int H, int relu, float *__restrict__ dy, int64_t lddy) {
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    const int r = blockIdx.x;
    if (c >= H) return;
    float d = dout[(int64_t)r * lddo + c];
    if (relu && !(out[(int64_t)r * ldo + c] > 0.f)) d = 0.f;
    dy[(int64_t)r * lddy + c] = d;
}

}  // namespace scvae

using namespace scvae;

extern "C" int64_t scvae_bn_scratch_floats(int M, int H, int groups) {
    if (M <= 0 || H <= 0 || groups <= 0 || M % groups) return -1;
    const int64_t H32 = round32(H);
    return 3 * (int64_t)groups * bn_chunks(M / groups) * H32 + (int64_t)groups * H32;
}

extern "C" int scvae_bn_act_fwd(const float *y, int64_t ldy, int M, int H, int groups,
                                const float *beta, float *moving_mean, float *moving_var,
                                int training, int update_moving, int relu, float *out, int64_t ldo,
                                float *save_mean, float *save_rstd, float *scratch, void *stream) {
    SCVAE_CHECK_ARG(y && beta && moving_mean && moving_var && out, "bn_act_fwd: NULL pointer");
    SCVAE_CHECK_ARG(M > 0 && H > 0 && groups > 0 && M % groups == 0, "bn_act_fwd: bad shape");
    SCVAE_CHECK_ARG(ldy >= H && ldo >= H, "bn_act_fwd: bad leading dimension");
    SCVAE_CHECK_ARG(!training || (save_mean && save_rstd && scratch),
                    "bn_act_fwd: training needs save_mean/save_rstd/scratch");
    cudaStream_t s = (cudaStream_t)stream;
    const int n = M / groups, nchunks = bn_chunks(n);
    BnScratch sc = bn_scratch(scratch, M, H, groups);
    const dim3 block(kBnCols, kBnLanes);
    if (training) {
        const dim3 grid(groups * nchunks, (H + kBnCols - 1) / kBnCols);
        bn_partial_kernel<<<grid, block, 0, s>>>(y, ldy, n, H, nchunks, sc.H32, sc.p0, sc.p1);
        SCVAE_CHECK_LAUNCH("bn_partial");
    }
    const dim3 grid2(groups * nchunks, (int)((ldo + kBnCols - 1) / kBnCols));
    bn_apply_kernel<<<grid2, block, 0, s>>>(y, ldy, n, H, nchunks, sc.H32, sc.p0, sc.p1, beta,
                                            moving_mean, moving_var, training, relu, out, ldo,
                                            save_mean, save_rstd, sc.var,
                                            (training && update_moving && groups == 1) ? moving_mean : nullptr,
                                            moving_var);
    SCVAE_CHECK_LAUNCH("bn_apply");
    if (training && update_moving && groups > 1) {
        bn_moving_kernel<<<(H + 127) / 128, 128, 0, s>>>(H, groups, sc.H32, save_mean, sc.var,
                                                         moving_mean, moving_var);
        SCVAE_CHECK_LAUNCH("bn_moving");
    }
    return 0;
}

extern "C" int scvae_bn_act_bwd(const float *dout, int64_t lddo, const float *y, int64_t ldy,
                                const float *out, int64_t ldo, int M, int H, int groups,
                                const float *save_mean, const float *save_rstd, int relu, float *dy,
                                int64_t lddy, float *dbeta, int accumulate_dbeta, float *scratch,
                                void *stream) {
    SCVAE_CHECK_ARG(dout && y && out && save_mean && save_rstd && dy && dbeta && scratch,
                    "bn_act_bwd: NULL pointer");
    SCVAE_CHECK_ARG(M > 0 && H > 0 && groups > 0 && M % groups == 0, "bn_act_bwd: bad shape");
    cudaStream_t s = (cudaStream_t)stream;
    const int n = M / groups, nchunks = bn_chunks(n);
    BnScratch sc = bn_scratch(scratch, M, H, groups);
    const dim3 block(kBnCols, kBnLanes);
    const dim3 grid(groups * nchunks, (H + kBnCols - 1) / kBnCols);
    bn_bwd_partial_kernel<<<grid, block, 0, s>>>(dout, lddo, y, ldy, out, ldo, n, H, nchunks, sc.H32,
                                                 save_mean, save_rstd, relu, sc.p1, sc.p2);
    SCVAE_CHECK_LAUNCH("bn_bwd_partial");
    bn_bwd_apply_kernel<<<grid, block, 0, s>>>(dout, lddo, y, ldy, out, ldo, n, H, groups, nchunks,
                                               sc.H32, save_mean, save_rstd, relu, sc.p1, sc.p2, dy,
                                               lddy, dbeta, accumulate_dbeta);
    SCVAE_CHECK_LAUNCH("bn_bwd_apply");
    return 0;
}

extern "C" int scvae_act_fwd(const float *y, int64_t ldy, int M, int H, int relu, float *out,
                             int64_t ldo, void *stream) {
    SCVAE_CHECK_ARG(y && out && M > 0 && H > 0 && ldo >= H, "act_fwd: bad arguments");
    const dim3 grid(M, (int)((ldo + 127) / 128));
    act_fwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(y, ldy, M, H, relu, out, ldo);
    SCVAE_CHECK_LAUNCH("act_fwd");
    return 0;
}

extern "C" int scvae_act_bwd(const float *dout, int64_t lddo, const float *out, int64_t ldo, int M,
                             int H, int relu, float *dy, int64_t lddy, void *stream) {
    SCVAE_CHECK_ARG(dout && out && dy && M > 0 && H > 0, "act_bwd: bad arguments");
    const dim3 grid(M, (H + 127) / 128);
    act_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(dout, lddo, out, ldo, M, H, relu, dy, lddy);
    SCVAE_CHECK_LAUNCH("act_bwd");
    return 0;
}
