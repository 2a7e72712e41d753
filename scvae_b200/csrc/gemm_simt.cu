// Exact-fp32 GEMM (FFMA, shared-memory tiled) for the small dense layers of the model
// (hidden->latent heads, latent->hidden, GMVAE cluster heads: MU:53-59) and as the fp32
// ground truth the tcgen05 kernel is checked against.  Any shape, any alignment, all three
// operand layouts of the forward / dgrad / wgrad products.
#include "common.cuh"

namespace scvae {

constexpr int TM = 64, TN = 64, TK = 16;

// C[m,n] = sum_k A(m,k) B(n,k) with generic element strides.
__global__ void __launch_bounds__(256)
gemm_f32_kernel(int M, int N, int K, const float *__restrict__ A, int64_t sam, int64_t sak,
                const float *__restrict__ B, int64_t sbn, int64_t sbk, float *__restrict__ C,
                int64_t ldc, int accumulate) {
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 4 x 4 outputs each
    float acc[4][4] = {};

    // loader mapping: if the k stride is 1 walk k fastest, else walk m/n fastest (coalescing)
    const bool a_kfast = (sak == 1), b_kfast = (sbk == 1);
    for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + i * 256;  // 1024 elements per operand tile
            int mm, kk;
            if (a_kfast) { kk = e & 15; mm = e >> 4; } else { mm = e & 63; kk = e >> 6; }
            const int gm = m0 + mm, gk = k0 + kk;
            As[kk][mm] = (gm < M && gk < K) ? A[(int64_t)gm * sam + (int64_t)gk * sak] : 0.f;
            int nn, kb;
            if (b_kfast) { kb = e & 15; nn = e >> 4; } else { nn = e & 63; kb = e >> 6; }
            const int gn = n0 + nn, gkb = k0 + kb;
            Bs[kb][nn] = (gn < N && gkb < K) ? B[(int64_t)gn * sbn + (int64_t)gkb * sbk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float *c = C + (int64_t)gm * ldc + gn;
            *c = accumulate ? *c + acc[i][j] : acc[i][j];
        }
    }
}

}  // namespace scvae

extern "C" int scvae_gemm_f32(int layout, int M, int N, int K, const float *A, int64_t lda,
                              const float *B, int64_t ldb, float *C, int64_t ldc, int accumulate,
                              void *stream) {
    using namespace scvae;
    SCVAE_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0, "gemm_f32: bad arguments");
    int64_t sam, sak, sbn, sbk;
    switch (layout) {
        case SCVAE_GEMM_NT: sam = lda; sak = 1; sbn = ldb; sbk = 1; break;
        case SCVAE_GEMM_NN: sam = lda; sak = 1; sbn = 1; sbk = ldb; break;
        case SCVAE_GEMM_TN: sam = 1; sak = lda; sbn = 1; sbk = ldb; break;
        default: set_error("gemm_f32: unknown layout %d", layout); return 1;
    }
    const dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM);
    SCVAE_CHECK_ARG(grid.y <= 65535, "gemm_f32: M too large for this kernel (%d)", M);
    gemm_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(M, N, K, A, sam, sak, B, sbn, sbk, C, ldc,
                                                            accumulate);
    SCVAE_CHECK_LAUNCH("gemm_f32");
    return 0;
}
