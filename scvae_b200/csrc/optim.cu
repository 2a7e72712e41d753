// a8: elementwise gradient clip + TensorFlow-style Adam (VAE:2736-2770), one fused pass over
// the flat parameter buffer.  tf.train.AdamOptimizer: lr_t = lr sqrt(1-b2^t)/(1-b1^t),
// theta -= lr_t m / (sqrt(v) + eps)  (epsilon outside the bias correction).
#include <cuda_fp16.h>

#include "common.cuh"

namespace scvae {

__global__ void __launch_bounds__(256)
adam_clip_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                 float *__restrict__ v, int64_t n, const int64_t *__restrict__ step, float lr,
                 float beta1, float beta2, float eps, float clip, float gscale) {
    __shared__ float s_lr_t;
    if (threadIdx.x == 0) {
        const double t = (double)(*step + 1);
        s_lr_t = (float)((double)lr * sqrt(1.0 - pow((double)beta2, t)) / (1.0 - pow((double)beta1, t)));
    }
    __syncthreads();
    const float lr_t = s_lr_t;
    const float ob1 = 1.f - beta1, ob2 = 1.f - beta2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float gi = g[i] * gscale;
        gi = fminf(fmaxf(gi, -clip), clip);
        const float mi = beta1 * m[i] + ob1 * gi;
        const float vi = beta2 * v[i] + ob2 * gi * gi;
        m[i] = mi;
        v[i] = vi;
        p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
    }
}

// fp16 operand copy of an fp32 master tensor, zero padded to the destination width
__global__ void f32_to_f16_kernel(const float *__restrict__ src, int64_t lds, int cols, __half *__restrict__ dst,
                                  int64_t ldd, float scale) {
    const int64_t r = blockIdx.x;
    for (int c = threadIdx.x + blockIdx.y * blockDim.x; c < ldd; c += blockDim.x * gridDim.y)
        dst[r * ldd + c] = __float2half_rn(c < cols ? src[r * lds + c] * scale : 0.f);
}

__global__ void step_advance_kernel(int64_t *step) { *step += 1; }

}  // namespace scvae

extern "C" int scvae_adam_clip_step(float *param, const float *grad, float *m, float *v, int64_t n,
                                    const int64_t *step, float lr, float beta1, float beta2,
                                    float epsilon, float clip, float grad_scale, void *stream) {
    using namespace scvae;
    SCVAE_CHECK_ARG(param && grad && m && v && step && n >= 0, "adam_clip_step: bad arguments");
    if (n == 0) return 0;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    adam_clip_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, m, v, n, step, lr,
                                                                    beta1, beta2, epsilon, clip,
                                                                    grad_scale);
    SCVAE_CHECK_LAUNCH("adam_clip_step");
    return 0;
}

extern "C" int scvae_step_advance(int64_t *step, void *stream) {
    using namespace scvae;
    SCVAE_CHECK_ARG(step, "step_advance: NULL");
    step_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step);
    SCVAE_CHECK_LAUNCH("step_advance");
    return 0;
}

extern "C" int scvae_f32_to_f16(const float *src, int64_t lds, int64_t rows, int cols, void *dst, int64_t ldd,
                                float scale, void *stream) {
    using namespace scvae;
    SCVAE_CHECK_ARG(src && dst && rows > 0 && cols > 0 && ldd >= cols, "f32_to_f16: bad arguments");
    const dim3 grid((unsigned)rows, (unsigned)((ldd + 1023) / 1024));
    f32_to_f16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, lds, cols, (__half *)dst, ldd, scale);
    SCVAE_CHECK_LAUNCH("f32_to_f16");
    return 0;
}
