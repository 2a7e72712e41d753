// a8: elementwise gradient clip + TensorFlow-style Adam (VAE:2736-2770), one fused pass over
// the flat parameter buffer.  tf.train.AdamOptimizer: lr_t = lr sqrt(1-b2^t)/(1-b1^t),
// theta -= lr_t m / (sqrt(v) + eps)  (epsilon outside the bias correction).
#include <cuda_fp16.h>

#include "common.cuh"

namespace scvae {

struct AdamShadows {
    scvae_shadow s[SCVAE_MAX_SHADOWS];
    int n;
};

// fp16 copies (and rounding remainders) of the parameters just updated: 4 consecutive floats of one
// row of a shadowed block -> 4 halves (8-byte store).  i: flat offset of the first float.
__device__ __forceinline__ void write_shadows(const AdamShadows &sh, int64_t i, const float4 &p4) {
    for (int k = 0; k < sh.n; ++k) {
        const scvae_shadow &s = sh.s[k];
        if (i < s.lo || i >= s.hi) continue;
        // (a shadowed block holds < 2^32 floats: 32-bit divisions, the kernel stays memory-bound)
        const uint32_t rel = (uint32_t)(i - s.lo);
        const uint32_t r = rel / (uint32_t)s.src_ld;
        const int c = (int)(rel - r * (uint32_t)s.src_ld);
        if (c >= s.dst_ld) continue;                    // (src_ld, dst_ld % 4 == 0: all four or none)
        const uint32_t blk = s.src_block_rows > r ? 0u : r / (uint32_t)s.src_block_rows;
        const int64_t row = (int64_t)blk * s.dst_block_rows + (r - blk * (uint32_t)s.src_block_rows);
        const float x[4] = {p4.x, p4.y, p4.z, p4.w};
        __align__(8) __half h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float xv = (c + j < s.cols) ? x[j] : 0.f;
            h[j] = __float2half_rn(xv);
            l[j] = __float2half_rn(xv - __half2float(h[j]));
        }
        *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(s.hi16) + row * s.dst_ld + c) = *reinterpret_cast<const uint2 *>(h);
        if (s.lo16)
            *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(s.lo16) + row * s.dst_ld + c) = *reinterpret_cast<const uint2 *>(l);
    }
}

__global__ void __launch_bounds__(256)
adam_clip_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                 float *__restrict__ v, int64_t n, int64_t *__restrict__ step, float lr,
                 float beta1, float beta2, float eps, float clip, float gscale,
                 const float *__restrict__ scalars, const AdamShadows sh, int *__restrict__ advance_counter,
                 int advance_total) {
    pdl_wait();          // (programmatic dependent launch: common.cuh)
    pdl_trigger();
    __shared__ float s_lr_t;
    if (threadIdx.x == 0) {
        const double t = (double)(*step + 1);
        const double lr_eff = (double)lr * (scalars ? (double)scalars[0] : 1.0);
        s_lr_t = adam_step_size(lr_eff, beta1, beta2, t);
    }
    __syncthreads();
    const float lr_t = s_lr_t;
    const float ob1 = 1.f - beta1, ob2 = 1.f - beta2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    auto update = [&](float &pi, float gi, float &mi, float &vi) {
        gi = fminf(fmaxf(gi * gscale, -clip), clip);
        mi = beta1 * mi + ob1 * gi;
        vi = beta2 * vi + ob2 * gi * gi;
        pi = pi - lr_t * mi / (sqrtf(vi) + eps);
    };
    const bool vec = aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v);
    const int64_t n4 = vec ? (n >> 2) : 0;
    for (int64_t i = tid; i < n4; i += stride) {
        float4 p4 = reinterpret_cast<float4 *>(p)[i];
        const float4 g4 = reinterpret_cast<const float4 *>(g)[i];
        float4 m4 = reinterpret_cast<float4 *>(m)[i];
        float4 v4 = reinterpret_cast<float4 *>(v)[i];
        update(p4.x, g4.x, m4.x, v4.x);
        update(p4.y, g4.y, m4.y, v4.y);
        update(p4.z, g4.z, m4.z, v4.z);
        update(p4.w, g4.w, m4.w, v4.w);
        reinterpret_cast<float4 *>(p)[i] = p4;
        reinterpret_cast<float4 *>(m)[i] = m4;
        reinterpret_cast<float4 *>(v)[i] = v4;
        if (sh.n) write_shadows(sh, i << 2, p4);
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += stride) update(p[i], g[i], m[i], v[i]);
    // the step counter advances when the LAST CTA of the `advance_total` CTAs that share the counter
    // (all optimiser launches of this step, on whichever streams) is done: every CTA read *step at
    // its start, so none can see the new value
    if (advance_counter) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(advance_counter, 1) == advance_total - 1) {
                *step += 1;
                *advance_counter = 0;
                __threadfence();
            }
        }
    }
}

// fp16 operand copy of an fp32 master tensor, zero padded to the destination width.
// One thread converts 8 consecutive columns (one 128-bit store); ldd % 8 == 0.
__global__ void __launch_bounds__(256)
f32_to_f16_kernel(const float *__restrict__ src, int64_t lds, int64_t rows, int cols, __half *__restrict__ dst,
                  int64_t ldd, float scale) {
    const int64_t groups = ldd >> 3;
    const int64_t total = rows * groups;
    const bool vec = (lds & 3) == 0 && aligned16(src);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / groups;
        const int c = (int)(i % groups) << 3;
        const float *sp = src + r * lds + c;
        float v[8];
        if (vec && c + 8 <= cols) {
            const float4 a = *reinterpret_cast<const float4 *>(sp);
            const float4 b = *reinterpret_cast<const float4 *>(sp + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (c + j < cols) ? sp[j] : 0.f;
        }
        uint4 pk;
        __half2 h0 = __floats2half2_rn(v[0] * scale, v[1] * scale), h1 = __floats2half2_rn(v[2] * scale, v[3] * scale);
        __half2 h2 = __floats2half2_rn(v[4] * scale, v[5] * scale), h3 = __floats2half2_rn(v[6] * scale, v[7] * scale);
        pk.x = *reinterpret_cast<uint32_t *>(&h0);
        pk.y = *reinterpret_cast<uint32_t *>(&h1);
        pk.z = *reinterpret_cast<uint32_t *>(&h2);
        pk.w = *reinterpret_cast<uint32_t *>(&h3);
        *reinterpret_cast<uint4 *>(dst + r * ldd + c) = pk;
    }
}

// hi = fp16(scale x), lo = fp16(scale x - hi): the pair carries ~22 mantissa bits of x.
__global__ void __launch_bounds__(256)
f32_to_f16_split_kernel(const float *__restrict__ src, int64_t lds, int64_t rows, int cols, __half *__restrict__ hi,
                        __half *__restrict__ lo, int64_t ldd, float scale) {
    const int64_t groups = ldd >> 3;
    const int64_t total = rows * groups;
    const bool vec = (lds & 3) == 0 && aligned16(src);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / groups;
        const int c = (int)(i % groups) << 3;
        const float *sp = src + r * lds + c;
        float v[8];
        if (vec && c + 8 <= cols) {
            const float4 a = *reinterpret_cast<const float4 *>(sp);
            const float4 b = *reinterpret_cast<const float4 *>(sp + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (c + j < cols) ? sp[j] : 0.f;
        }
        __align__(16) __half h[8], l[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float x = v[j] * scale;
            h[j] = __float2half_rn(x);
            l[j] = __float2half_rn(x - __half2float(h[j]));
        }
        *reinterpret_cast<uint4 *>(hi + r * ldd + c) = *reinterpret_cast<const uint4 *>(h);
        *reinterpret_cast<uint4 *>(lo + r * ldd + c) = *reinterpret_cast<const uint4 *>(l);
    }
}

__global__ void step_advance_kernel(int64_t *step) { *step += 1; }

}  // namespace scvae

static int64_t adam_blocks(int64_t n) {
    int64_t blocks = ((n >> 2) + 255) / 256;       // one float4 per thread and trip
    // at most four CTAs per SM, all resident at once: the double-precision step-size prologue of a CTA
    // is then paid once per launch, not once per wave
    if (blocks > 148 * 4) blocks = 148 * 4;
    return blocks < 1 ? 1 : blocks;
}

extern "C" int scvae_adam_clip_ctas(int64_t n) { return (int)adam_blocks(n); }

extern "C" int scvae_adam_clip_step(float *param, const float *grad, float *m, float *v, int64_t n,
                                    int64_t *step, float lr, float beta1, float beta2,
                                    float epsilon, float clip, float grad_scale, const float *scalars,
                                    const scvae_shadow *shadows, int n_shadows, int *advance_counter,
                                    int advance_total, void *stream) {
    using namespace scvae;
    SCVAE_CHECK_ARG(param && grad && m && v && step && n >= 0, "adam_clip_step: bad arguments");
    SCVAE_CHECK_ARG(n_shadows >= 0 && n_shadows <= SCVAE_MAX_SHADOWS && (n_shadows == 0 || shadows),
                    "adam_clip_step: at most %d shadows", SCVAE_MAX_SHADOWS);
    if (n == 0) return 0;
    AdamShadows sh;
    sh.n = n_shadows;
    for (int k = 0; k < n_shadows; ++k) {
        const scvae_shadow &s = shadows[k];
        SCVAE_CHECK_ARG(s.hi16 && s.lo % 4 == 0 && s.hi % 4 == 0 && s.src_ld % 4 == 0 && s.dst_ld % 4 == 0 &&
                            s.hi - s.lo < (int64_t)1 << 32 && s.src_ld < (int64_t)1 << 31 &&
                            s.src_block_rows > 0 && s.dst_block_rows > 0 && aligned16(param) &&
                            (reinterpret_cast<uintptr_t>(s.hi16) & 7u) == 0,
                        "adam_clip_step: bad shadow %d", k);
        sh.s[k] = s;
    }
    launch_pdl(kPdlAdam, adam_clip_kernel, dim3((unsigned)::adam_blocks(n)), dim3(256), 0, (cudaStream_t)stream,
               param, grad, m, v, n, step, lr, beta1, beta2, epsilon, clip, grad_scale, scalars, sh, advance_counter,
               advance_total);
    SCVAE_CHECK_LAUNCH("adam_clip_step");
    return 0;
}

extern "C" int scvae_f32_to_f16_split(const float *src, int64_t lds, int64_t rows, int cols, void *hi, void *lo,
                                      int64_t ldd, float scale, void *stream) {
    using namespace scvae;
    SCVAE_CHECK_ARG(src && hi && lo && rows > 0 && cols > 0 && ldd >= cols, "f32_to_f16_split: bad arguments");
    SCVAE_CHECK_ARG(ldd % 8 == 0 && aligned16(hi) && aligned16(lo), "f32_to_f16_split: destination rows must be 16-byte multiples");
    int64_t blocks = (rows * (ldd >> 3) + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    f32_to_f16_split_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, lds, rows, cols, (__half *)hi,
                                                                              (__half *)lo, ldd, scale);
    SCVAE_CHECK_LAUNCH("f32_to_f16_split");
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
    SCVAE_CHECK_ARG(ldd % 8 == 0 && aligned16(dst), "f32_to_f16: destination rows must be 16-byte multiples");
    int64_t blocks = (rows * (ldd >> 3) + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    f32_to_f16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, lds, rows, cols, (__half *)dst, ldd,
                                                                        scale);
    SCVAE_CHECK_LAUNCH("f32_to_f16");
    return 0;
}
