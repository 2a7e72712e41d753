// f3: dropout on the INPUT of a dense layer (MU:45-50, tf.contrib.layers.dropout): inverted
// dropout, x * keep_mask / keep_probability while training.  The reference draws one mask per
// site: every encoder / decoder layer, and one per posterior / likelihood head over the same
// activation (VAE:2286, :2487, :2516).  A site's mask is defined by a noise tensor and a
// threshold (keep <=> noise < threshold): standard-normal noise from scvae_fill_normal with the
// threshold at the normal quantile of the keep probability, or injected +-inf-like values in
// the parity tests.  The kernels work on the bias-augmented operand layout: the masked logical
// columns [0, n) skip the physical column `skip_col` (the ones column when extras follow it,
// VAE:2400-2441), every other column of the `width` stored ones is copied through.
#include "common.cuh"

namespace scvae {

__global__ void dropout_fwd_kernel(const float *__restrict__ x, int64_t ldx, int n, int skip_col,
                                   const float *__restrict__ noise, float threshold,
                                   float inv_keep, float *__restrict__ out, int64_t ldo,
                                   int width) {
    const int pc = blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t r = blockIdx.x;
    if (pc >= width) return;
    float v = x[r * ldx + pc];
    if (pc != skip_col) {
        const int c = pc < skip_col ? pc : pc - 1;
        if (c < n) v = (noise[r * n + c] < threshold) ? v * inv_keep : 0.f;
    }
    out[r * ldo + pc] = v;
}

// dx (+)= dsrc * keep_mask / keep on the masked columns (dsrc == NULL: in place on dx); columns
// outside the masked range are left untouched.
__global__ void dropout_bwd_kernel(float *__restrict__ dx, int64_t lddx, int n, int skip_col,
                                   const float *__restrict__ noise, float threshold,
                                   float inv_keep, const float *__restrict__ dsrc, int64_t ldds,
                                   int accumulate) {
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t r = blockIdx.x;
    if (c >= n) return;
    const int pc = c < skip_col ? c : c + 1;
    const float d = dsrc ? dsrc[r * ldds + pc] : dx[r * lddx + pc];
    const float g = (noise[r * n + c] < threshold) ? d * inv_keep : 0.f;
    dx[r * lddx + pc] = (accumulate && dsrc) ? dx[r * lddx + pc] + g : g;
}

}  // namespace scvae

using namespace scvae;

extern "C" int scvae_dropout_fwd(const float *x, int64_t ldx, int rows, int n, int skip_col,
                                 const float *noise, float threshold, float keep, float *out,
                                 int64_t ldo, int width, void *stream) {
    SCVAE_CHECK_ARG(x && noise && out && rows > 0 && n > 0 && width > 0 && keep > 0.f,
                    "dropout_fwd: bad arguments");
    SCVAE_CHECK_ARG(ldx >= width && ldo >= width && skip_col >= 0 && skip_col <= n &&
                        n + (skip_col < n ? 1 : 0) <= width,
                    "dropout_fwd: bad layout");
    const dim3 grid(rows, (width + 127) / 128);
    dropout_fwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(x, ldx, n, skip_col, noise, threshold,
                                                              1.f / keep, out, ldo, width);
    SCVAE_CHECK_LAUNCH("dropout_fwd");
    return 0;
}

extern "C" int scvae_dropout_bwd(float *dx, int64_t lddx, int rows, int n, int skip_col,
                                 const float *noise, float threshold, float keep,
                                 const float *dsrc, int64_t ldds, int accumulate, void *stream) {
    SCVAE_CHECK_ARG(dx && noise && rows > 0 && n > 0 && keep > 0.f, "dropout_bwd: bad arguments");
    SCVAE_CHECK_ARG(skip_col >= 0 && skip_col <= n, "dropout_bwd: bad layout");
    const dim3 grid(rows, (n + 127) / 128);
    dropout_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(dx, lddx, n, skip_col, noise, threshold,
                                                              1.f / keep, dsrc, ldds, accumulate);
    SCVAE_CHECK_LAUNCH("dropout_bwd");
    return 0;
}
