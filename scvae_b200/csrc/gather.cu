// a1: minibatch gather.  Replaces the host-side x_train[idx].toarray() (VAE:994-998,
// GMVAE:1078-1082): the CSR count matrix lives in HBM and each step densifies B rows into the
// augmented (B, ldx) fp32 layout the GEMMs and the likelihood kernels read.  HBM-write-bound.
#include "common.cuh"

namespace scvae {

__global__ void __launch_bounds__(256)
csr_densify_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                   const float *__restrict__ values, const int64_t *__restrict__ rows, int G,
                   float *__restrict__ x, int64_t ldx, float *__restrict__ row_const, int rebase,
                   uint16_t *__restrict__ t16, int64_t ldt16) {
    __shared__ float red[32];
    const int b = blockIdx.x;
    const int64_t row = rows ? rows[b] : b;
    float *xr = x + (int64_t)b * ldx;
    // zero fill (+ the augmented ones column)
    if ((ldx & 3) == 0 && aligned16(x)) {
        float4 *x4 = reinterpret_cast<float4 *>(xr);
        const int n4 = (int)(ldx >> 2);
        for (int i = threadIdx.x; i < n4; i += blockDim.x) {
            float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const int c = i << 2;
            if (G >= c && G < c + 4) (&z.x)[G - c] = 1.f;
            x4[i] = z;
        }
    } else {
        for (int i = threadIdx.x; i < ldx; i += blockDim.x) xr[i] = (i == G) ? 1.f : 0.f;
    }
    if (t16) {  // 16-bit copy of the counts for the fused likelihood heads (zero fill)
        uint4 *t4 = reinterpret_cast<uint4 *>(t16 + (int64_t)b * ldt16);
        const int n8 = (int)(ldt16 >> 3);
        for (int i = threadIdx.x; i < n8; i += blockDim.x) t4[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    const int64_t base = rebase ? indptr[0] : 0;
    const int64_t s = indptr[row] - base, e = indptr[row + 1] - base;
    float acc = 0.f;
    for (int64_t i = s + threadIdx.x; i < e; i += blockDim.x) {
        const float v = values[i];
        const int c = indices[i];
        if (c >= 0 && c < G) {
            xr[c] = v;
            if (t16) t16[(int64_t)b * ldt16 + c] = (uint16_t)fminf(fmaxf(v, 0.f), 65535.f);
        }
        if (v > 0.f) acc += lgammaf(1.f + v);
    }
    if (row_const) {
        const float tot = block_sum(acc, red);
        if (threadIdx.x == 0) row_const[b] = tot;
    }
}

// dense fp32 counts -> u16 (clamped), zero padded to ldt16 columns
__global__ void f32_to_u16_kernel(const float *__restrict__ x, int64_t ldx, int G, uint16_t *__restrict__ t16,
                                  int64_t ldt16) {
    const int64_t r = blockIdx.x;
    for (int c = threadIdx.x + blockIdx.y * blockDim.x; c < ldt16; c += blockDim.x * gridDim.y)
        t16[r * ldt16 + c] = c < G ? (uint16_t)fminf(fmaxf(x[r * ldx + c], 0.f), 65535.f) : (uint16_t)0;
}

}  // namespace scvae

extern "C" int scvae_csr_densify(const int64_t *indptr, const int32_t *indices, const float *values,
                                 const int64_t *rows, int B, int G, float *x, int64_t ldx,
                                 float *row_const, int rebase, void *t16, int64_t ldt16, void *stream) {
    using namespace scvae;
    SCVAE_CHECK_ARG(indptr && indices && values && x, "csr_densify: NULL pointer");
    SCVAE_CHECK_ARG(B >= 0 && G > 0 && ldx >= G, "csr_densify: bad shape (B=%d G=%d ldx=%lld)", B, G,
                    (long long)ldx);
    SCVAE_CHECK_ARG(!t16 || (ldt16 % 8 == 0 && ldt16 >= G && aligned16(t16)), "csr_densify: bad t16 layout");
    if (B == 0) return 0;
    csr_densify_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(indptr, indices, values, rows, G, x, ldx,
                                                            row_const, rebase, (uint16_t *)t16, ldt16);
    SCVAE_CHECK_LAUNCH("csr_densify");
    return 0;
}

extern "C" int scvae_f32_to_u16(const float *x, int64_t ldx, int64_t rows, int G, void *t16, int64_t ldt16,
                                void *stream) {
    using namespace scvae;
    SCVAE_CHECK_ARG(x && t16 && rows > 0 && G > 0 && ldt16 >= G, "f32_to_u16: bad arguments");
    const dim3 grid((unsigned)rows, (unsigned)((ldt16 + 1023) / 1024));
    f32_to_u16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, ldx, G, (uint16_t *)t16, ldt16);
    SCVAE_CHECK_LAUNCH("f32_to_u16");
    return 0;
}
