// J1: the latency-bound middle of a VAE training step as TWO persistent kernels.
//
// Between the two gene-axis products of a step (first encoder layer, likelihood heads) the
// reference runs, per `dense_layer` (MU:53-74) FC -> batch_norm -> ReLU, then the posterior heads,
// the reparameterised sample and the analytic KL (VAE:2280-2369, :2624-2627) and the decoder
// layers -- all on (cells x ~100) tensors.  As separate launches that is ~10 kernels forward and
// ~12 backward of 4-10 us each, on the critical path.  Here each direction is ONE kernel:
//
//   vae_mid_fwd:  [split-K partials of x W1^T] -> BN -> ReLU -> (FC -> BN -> ReLU)* -> posterior FC
//                 -> clip / sample / KL -> (FC -> BN -> ReLU)* -> fp16 operand of the fused heads
//   vae_mid_bwd:  [gene-range partials of the decoder gradient, log p partials] -> ELBO, first-order
//                 fp16 correction of log p -> (BN/ReLU bwd -> wgrad -> dgrad)* -> sample/KL bwd ->
//                 posterior wgrad/dgrad -> (BN/ReLU bwd -> wgrad -> dgrad)* -> fp16 dY1 for the first
//                 layer's weight-gradient product; weight-gradient partials reduced in fixed order
//
// Every dense product is the GEMM of its layer with the batch norm, activation and sample fused
// behind it ("epilogue"): a CTA owns a slab of <= 64 cells, keeps the slab's activations in shared
// memory (column-major, so that forward, dgrad and wgrad all stream them with 128-bit loads) and
// multiplies with exact fp32 FFMA against the layer's weights staged in shared memory -- these
// products are ~0.1 % of the step's FLOPs, and exact fp32 removes the tf32 truncation bias
// (-1e-3 relative on mu, -2e-3 on KL) that the tensor-core path had here.  Batch statistics couple
// all cells: per-CTA (count, mean, M2) partials, one grid barrier, then every CTA folds the
// partials in the same fixed order (Chan) -- deterministic, no atomics on data.
//
// The kernels are latency-, not throughput-bound, so the code is organised around global-memory
// round trips: every phase issues ALL of its global loads before consuming any of them (fully
// unrolled, predicated batches), and loads that do not depend on other CTAs (the next product's
// weights, the next layer's stored activations) are issued ahead of the grid barrier or the
// arithmetic they would otherwise wait behind.
#include <cuda_fp16.h>
#include <curand_kernel.h>

#include "common.cuh"

namespace scvae {

constexpr int kMidThreads = 256;
constexpr int kMidCols = 128;       // widest activation held in shared memory
constexpr int kMidRows = 64;        // most cells per CTA
constexpr int kMidWP = 132;         // row pitch of the staged weight tile: 4 mod 32 (conflict-free 128-bit rows)
constexpr int kMidWIt = kMidCols * (kMidCols / 4) / kMidThreads;   // most float4 of a weight tile per thread (16)
constexpr int kMidSlots = 2 * SCVAE_MID_MAX_LAYERS;                // batch-normed layers (encoder + decoder)
constexpr long long kMidSpinLimit = 4000000000ll;   // ~2 s of clock64 ticks: bounded barrier wait

typedef scvae_mid_layer MidLayer;
typedef scvae_mid_desc MidDesc;

// row pitch of the activation buffers: a multiple of 4 (128-bit rows) that is 4 mod 32, so that
// the 8 lanes of a 128-bit shared-memory phase reading 8 consecutive columns hit distinct banks
__host__ __device__ static inline int mid_pitch(int rows) {
    const int r4 = ((rows < 32 ? 32 : rows) + 3) & ~3;      // the product tile spans >= 32 rows
    return r4 + ((36 - r4 % 32) % 32);
}
// shared memory: 4 activation buffers, the weight tile, per-column scratch, normalisation vectors
__host__ __device__ static inline int mid_smem_floats(int rows) {
    return 4 * kMidCols * mid_pitch(rows) + kMidCols * kMidWP + 8 * kMidCols + kMidSlots * 3 * kMidCols + 64;
}

struct MidCtx {
    float *act[4];    // activation buffers [128][RP] (column-major: column c of cell r at c RP + r)
    float *sw;        // staged weights [<=128][kMidWP], natural layout sw[n][k] = W[row0 + n][k]
    float *stat;      // [8][128] per-column scratch
    float *bnv;       // [slots][3][128]: mean, rstd, beta of every batch-normed layer
    float *red;       // 64 floats
    int RP;           // activation row pitch
    int r0, nr;       // first cell / cells of this CTA
    int nr4;
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Grid-wide barrier (all CTAs are co-resident: grid <= number of SMs, one CTA per SM).  The
// counter's top bit flips once per barrier, so no reset between barriers or launches is needed.
// The wait is bounded: a timeout sets *err and lets the kernel finish (results then invalid).
__device__ __forceinline__ void mid_grid_sync(unsigned *bar, int *err) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned add = (blockIdx.x == 0) ? (0x80000000u - (gridDim.x - 1)) : 1u;
        __threadfence();
        const unsigned old = atomicAdd(bar, add);
        const long long t0 = clock64();
        while (((old ^ ld_acquire_u32(bar)) & 0x80000000u) == 0) {
            if (clock64() - t0 > kMidSpinLimit) {
                atomicExch(err, 1);
                break;
            }
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ int mid_rows_of(int cta, int rows_per_cta, int B) {
    const int lo = cta * rows_per_cta;
    return max(0, min(rows_per_cta, B - lo));
}

// ---- batched slab access: one global round trip per slab -------------------------------------------
// element i of a slab of `ncol` columns: column i % ncol of cell i / ncol (lanes along the column
// index: coalesced global access).  SLAB_IT (a constexpr of the enclosing function) bounds the
// elements per thread: 128 columns x pitch / 256 threads.
#define MID_SLAB(total, ncol, v, LOAD)                                            \
    float v[SLAB_IT];                                                             \
    _Pragma("unroll") for (int it_ = 0; it_ < SLAB_IT; ++it_) {                   \
        const int i = threadIdx.x + it_ * kMidThreads;                            \
        const int col = i % (ncol), r = i / (ncol);                               \
        (void)col; (void)r;                                                       \
        v[it_] = (i < (total)) ? (LOAD) : 0.f;                                    \
    }
#define MID_SLAB_FOR(total, ncol, v, BODY)                                        \
    _Pragma("unroll") for (int it_ = 0; it_ < SLAB_IT; ++it_) {                   \
        const int i = threadIdx.x + it_ * kMidThreads;                            \
        if (i < (total)) {                                                        \
            const int col = i % (ncol), r = i / (ncol);                           \
            const float x = v[it_];                                               \
            (void)col; (void)r; (void)x;                                          \
            BODY                                                                  \
        }                                                                         \
    }
template <int TM>
struct MidTraits {
    static constexpr int kSlabIt = kMidCols * (TM == 4 ? 36 : 68) / kMidThreads;   // 18 or 34
};

// ---- weight staging: sw[n][k] = W[row0 + n][k], n < N, k < round4(Kc), zeros beyond Kc ----------------
// issue (loads into registers) and commit (stores to shared memory) are separate so that the loads can
// be in flight across a barrier or another phase's arithmetic.
struct WRegs {
    float4 v[kMidWIt];
};
__device__ __forceinline__ void stage_w_issue(WRegs &w, const float *__restrict__ W, int64_t ldw, int row0, int N, int Kc) {
    const int K4 = (Kc + 3) >> 2;
#pragma unroll
    for (int it = 0; it < kMidWIt; ++it) {
        const int i = threadIdx.x + it * kMidThreads;
        if (i < N * K4) {
            const int n = i / K4, k = (i % K4) << 2;
            w.v[it] = __ldg(reinterpret_cast<const float4 *>(W + (int64_t)(row0 + n) * ldw + k));   // ldw % 4 == 0
        }
    }
}
__device__ __forceinline__ void stage_w_commit(const MidCtx &c, const WRegs &w, int N, int Kc) {
    const int K4 = (Kc + 3) >> 2;
#pragma unroll
    for (int it = 0; it < kMidWIt; ++it) {
        const int i = threadIdx.x + it * kMidThreads;
        if (i < N * K4) {
            const int n = i / K4, k = (i % K4) << 2;
            float4 v = w.v[it];
            if (k + 1 >= Kc) v.y = 0.f;
            if (k + 2 >= Kc) v.z = 0.f;
            if (k + 3 >= Kc) v.w = 0.f;
            *reinterpret_cast<float4 *>(c.sw + n * kMidWP + k) = v;
        }
    }
}
__device__ __forceinline__ void stage_w(const MidCtx &c, const float *__restrict__ W, int64_t ldw, int row0, int N, int Kc) {
    WRegs w;
    stage_w_issue(w, W, ldw, row0, N, Kc);
    stage_w_commit(c, w, N, Kc);
}

// ---- forward product tile: acc[i][j] = sum_k A[k][ty TM + i] * sw[tx + 32 j][k] ---------------------
// (weights in their natural (out, in) layout; the 8 lanes of a 128-bit phase read 8 consecutive rows of
// sw, 4 mod 32 floats apart: conflict free).  A must be finite (zero) on the rows [K, round4(K)).
template <int TM>
__device__ __forceinline__ void mm_fwd(const float *__restrict__ A, int RP, const float *__restrict__ sw, int K, int N,
                                       float (&acc)[TM][4]) {
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
    const float *a = A + ty * TM;
    const int K4 = (K + 3) & ~3;
#pragma unroll 2
    for (int k = 0; k < K4; k += 4) {
        float4 wv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            wv[j] = (tx + 32 * j < N) ? *reinterpret_cast<const float4 *>(sw + (tx + 32 * j) * kMidWP + k)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            float av[TM];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                const float4 t = *reinterpret_cast<const float4 *>(a + (k + kk) * RP + i);
                av[i] = t.x; av[i + 1] = t.y; av[i + 2] = t.z; av[i + 3] = t.w;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float wk = kk == 0 ? wv[j].x : kk == 1 ? wv[j].y : kk == 2 ? wv[j].z : wv[j].w;
#pragma unroll
                for (int i = 0; i < TM; ++i) acc[i][j] = fmaf(av[i], wk, acc[i][j]);
            }
        }
    }
}
// acc -> C[n][r] (shared, column-major) and, when Y is given, -> Y[r0 + r][n] (HBM, row-major);
// column n of thread (ty, tx), slot j is tx + 32 j.
template <int TM>
__device__ __forceinline__ void mm_fwd_store(const MidCtx &c, float *__restrict__ C, int N, const float (&acc)[TM][4],
                                             float *__restrict__ Y, int64_t ldy) {
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = tx + 32 * j;
        if (n < N) {
#pragma unroll
            for (int i = 0; i < TM; i += 4)
                *reinterpret_cast<float4 *>(C + n * c.RP + ty * TM + i) =
                    make_float4(acc[i][j], acc[i + 1][j], acc[i + 2][j], acc[i + 3][j]);
            if (Y) {
#pragma unroll
                for (int i = 0; i < TM; ++i)
                    if (ty * TM + i < c.nr) Y[(int64_t)(c.r0 + ty * TM + i) * ldy + n] = acc[i][j];
            }
        }
    }
}
// C[n][r] = sum_k A[k][r] W[n][k] for the staged weights (shared -> shared [+ HBM]).
template <int TM>
__device__ __forceinline__ void mid_product(const MidCtx &c, const float *A, int K, float *C, int N, float *Y, int64_t ldy) {
    float acc[TM][4] = {};
    mm_fwd<TM>(A, c.RP, c.sw, K, N, acc);
    mm_fwd_store<TM>(c, C, N, acc, Y, ldy);
}

// ---- dgrad product tile: acc[i][j] += sum_n G[n][ty TM + i] * sw[n][tx 4 + j] -----------------------
template <int TM>
__device__ __forceinline__ void mm_dgrad(const float *__restrict__ G, int RP, const float *__restrict__ sw, int Nred,
                                         float (&acc)[TM][4]) {
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
    const float *a = G + ty * TM;
    const float *w = sw + tx * 4;
#pragma unroll 4
    for (int n = 0; n < Nred; ++n) {
        const float4 wv = *reinterpret_cast<const float4 *>(w + n * kMidWP);
        float av[TM];
#pragma unroll
        for (int i = 0; i < TM; i += 4) {
            const float4 t = *reinterpret_cast<const float4 *>(a + n * RP + i);
            av[i] = t.x; av[i + 1] = t.y; av[i + 2] = t.z; av[i + 3] = t.w;
        }
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            acc[i][0] = fmaf(av[i], wv.x, acc[i][0]);
            acc[i][1] = fmaf(av[i], wv.y, acc[i][1]);
            acc[i][2] = fmaf(av[i], wv.z, acc[i][2]);
            acc[i][3] = fmaf(av[i], wv.w, acc[i][3]);
        }
    }
}
template <int TM>
__device__ __forceinline__ void mm_dgrad_store(const MidCtx &c, float *__restrict__ C, int Kout, const float (&acc)[TM][4]) {
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int k = tx * 4 + j;
        if (k < Kout) {
#pragma unroll
            for (int i = 0; i < TM; i += 4)
                *reinterpret_cast<float4 *>(C + k * c.RP + ty * TM + i) =
                    make_float4(acc[i][j], acc[i + 1][j], acc[i + 2][j], acc[i + 3][j]);
        }
    }
}

// ---- weight-gradient partial of this CTA: out[n][k] = sum_{r < nr} G[n][r] * I[k][r] ----------
// thread (ty, tx) of a 16 x 16 grid owns n = ty + 16 i, k = tx + 16 j.  out: (N, Kp) row-major
// in the workspace (Kp = padded input width = leading dimension of the weight).
__device__ __noinline__ void mid_wgrad(const MidCtx &c, const float *__restrict__ G, int N, const float *__restrict__ I,
                                       int Kp, float *__restrict__ out) {
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const int ni = (N - ty + 15) >> 4, kj = (Kp - tx + 15) >> 4;    // valid i / j counts
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int r = 0; r < c.nr4; r += 4) {
        float4 g[8], x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            g[i] = i < ni ? *reinterpret_cast<const float4 *>(G + (ty + 16 * i) * c.RP + r) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            x[j] = j < kj ? *reinterpret_cast<const float4 *>(I + (tx + 16 * j) * c.RP + r) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                acc[i][j] = fmaf(g[i].x, x[j].x, fmaf(g[i].y, x[j].y, fmaf(g[i].z, x[j].z, fmaf(g[i].w, x[j].w, acc[i][j]))));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (i < ni && j < kj) out[(int64_t)(ty + 16 * i) * Kp + tx + 16 * j] = acc[i][j];
}

__device__ __forceinline__ void zero_cols(const MidCtx &c, float *dst, int col0, int col1) {
    for (int i = col0 * c.RP + threadIdx.x; i < col1 * c.RP; i += kMidThreads) dst[i] = 0.f;
}
// augmented ones column (valid cells only) at `col`, zero columns behind it up to `col_end`
__device__ __forceinline__ void set_aug_cols(const MidCtx &c, float *dst, int col, int col_end) {
    for (int i = col * c.RP + threadIdx.x; i < col_end * c.RP; i += kMidThreads)
        dst[i] = (i < (col + 1) * c.RP && i - col * c.RP < c.nr) ? 1.f : 0.f;
}

__device__ __forceinline__ void chan_merge2(float &cnt, float &mu, float &m2, float nb, float mb, float qb) {
    const float tot = cnt + nb;
    if (tot > 0.f) {
        const float delta = mb - mu;
        const float f = nb / tot;
        mu += delta * f;
        m2 += qb + delta * delta * (cnt * f);
        cnt = tot;
    }
}

// Per-CTA partial statistics of Y[c][r] (columns < N over this CTA's cells): (mean, M2) -> workspace.
__device__ __forceinline__ void mid_bn_partial(const MidCtx &c, const float *Y, int N, float *ws_stat) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int col = warp; col < N; col += kMidThreads / 32) {
        float s = 0.f;
        for (int r = lane; r < c.nr; r += 32) s += Y[col * c.RP + r];
        s = warp_sum(s);
        const float mean = c.nr > 0 ? s / (float)c.nr : 0.f;
        float q = 0.f;
        for (int r = lane; r < c.nr; r += 32) {
            const float dv = Y[col * c.RP + r] - mean;
            q += dv * dv;
        }
        q = warp_sum(q);
        if (lane == 0) {
            ws_stat[((int64_t)blockIdx.x * 2 + 0) * kMidCols + col] = mean;
            ws_stat[((int64_t)blockIdx.x * 2 + 1) * kMidCols + col] = q;
        }
    }
}
// Every CTA folds all partials in the same fixed order: warp w takes the CTAs k = w mod 8 (4 columns
// per lane, their loads independent and batched: the fold costs ~2 L2 round trips, not one per
// partial), then the 8 per-warp results in warp order.  `scratch`: [8 warps][3][128] floats of idle
// shared memory.  Result: s_mean / s_rstd [N]; CTA 0 writes the saved statistics and moving averages.
__device__ __forceinline__ void mid_bn_fold(const MidCtx &c, const MidDesc &d, const MidLayer &l, const float *ws_stat,
                                            float *scratch, float *s_mean, float *s_rstd) {
    const int N = l.n_out;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x;
    float cnt[4] = {0.f, 0.f, 0.f, 0.f}, mu[4] = {0.f, 0.f, 0.f, 0.f}, m2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 10
    for (int k = warp; k < G; k += kMidThreads / 32) {
        const float nb = (float)mid_rows_of(k, d.rows_per_cta, d.B);
        float mb[4], qb[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int col = lane + 32 * q;
            mb[q] = col < N ? __ldcg(ws_stat + ((int64_t)k * 2 + 0) * kMidCols + col) : 0.f;
            qb[q] = col < N ? __ldcg(ws_stat + ((int64_t)k * 2 + 1) * kMidCols + col) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) chan_merge2(cnt[q], mu[q], m2[q], nb, mb[q], qb[q]);
    }
    float *x = scratch;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        x[(warp * 3 + 0) * kMidCols + lane + 32 * q] = cnt[q];
        x[(warp * 3 + 1) * kMidCols + lane + 32 * q] = mu[q];
        x[(warp * 3 + 2) * kMidCols + lane + 32 * q] = m2[q];
    }
    __syncthreads();
    const int col = threadIdx.x;
    if (col < N) {
        float cn = 0.f, mean_all = 0.f, q_all = 0.f;
        for (int w = 0; w < kMidThreads / 32; ++w)
            chan_merge2(cn, mean_all, q_all, x[(w * 3 + 0) * kMidCols + col], x[(w * 3 + 1) * kMidCols + col],
                        x[(w * 3 + 2) * kMidCols + col]);
        const float var = cn > 0.f ? q_all / cn : 0.f;
        const float rstd = rsqrtf(var + kBnEps);
        s_mean[col] = mean_all;
        s_rstd[col] = rstd;
        if (blockIdx.x == 0) {
            l.mean[col] = mean_all;
            l.rstd[col] = rstd;
            if (d.update_moving) {      // Bessel-corrected variance (tf fused batch norm)
                const float vu = var * ((float)d.B / (float)max(d.B - 1, 1));
                l.moving_mean[col] -= (1.f - kBnDecay) * (l.moving_mean[col] - mean_all);
                l.moving_var[col] -= (1.f - kBnDecay) * (l.moving_var[col] - vu);
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void mid_setup(MidCtx &c, const MidDesc &d, float *smem) {
    const int rows = d.rows_per_cta;
    c.RP = mid_pitch(rows);
    for (int b = 0; b < 4; ++b) c.act[b] = smem + b * kMidCols * c.RP;
    c.sw = c.act[3] + kMidCols * c.RP;
    c.stat = c.sw + kMidCols * kMidWP;
    c.bnv = c.stat + 8 * kMidCols;
    c.red = c.bnv + kMidSlots * 3 * kMidCols;
    c.r0 = blockIdx.x * rows;
    c.nr = mid_rows_of(blockIdx.x, rows, d.B);
    c.nr4 = (c.nr + 3) & ~3;
}

// workspace map (floats): [BN slots][grid][2][128] | [grid][4] bound partials | dW partials
__device__ __forceinline__ float *ws_stat_slot(const MidDesc &d, int slot) {
    return d.workspace + (int64_t)slot * gridDim.x * 2 * kMidCols;
}
__device__ __forceinline__ float *ws_bound(const MidDesc &d) {
    return d.workspace + (int64_t)kMidSlots * gridDim.x * 2 * kMidCols;
}
__device__ __forceinline__ float *ws_dw_base(const MidDesc &d) { return ws_bound(d) + (int64_t)gridDim.x * 4; }

// mean / rstd / beta of layer slot `s` (encoder layers first, then decoder layers) in shared memory
__device__ __forceinline__ float *bnv_mean(const MidCtx &c, int s) { return c.bnv + (s * 3 + 0) * kMidCols; }
__device__ __forceinline__ float *bnv_rstd(const MidCtx &c, int s) { return c.bnv + (s * 3 + 1) * kMidCols; }
__device__ __forceinline__ float *bnv_beta(const MidCtx &c, int s) { return c.bnv + (s * 3 + 2) * kMidCols; }

// Normalisation vectors of every layer -> shared memory.  mode 0: beta only (training forward: the
// batch statistics follow from the folds), 1: + moving statistics (evaluation), 2: + saved batch
// statistics (backward).  Layers without batch norm get (0, 1, 0): the identity.
__device__ __forceinline__ void mid_load_bnv(const MidCtx &c, const MidDesc &d, int mode) {
    const int nl = d.n_enc + d.n_dec;
    for (int i = threadIdx.x; i < nl * kMidCols; i += kMidThreads) {
        const int s = i / kMidCols, col = i % kMidCols;
        const MidLayer &l = s < d.n_enc ? d.enc[s] : d.dec[s - d.n_enc];
        float mean = 0.f, rstd = 1.f, beta = 0.f;
        if (l.beta && col < l.n_out) {
            beta = l.beta[col];
            if (mode == 1) {
                mean = l.moving_mean[col];
                rstd = rsqrtf(l.moving_var[col] + kBnEps);
            } else if (mode == 2) {
                mean = l.mean[col];
                rstd = l.rstd[col];
            }
        }
        bnv_mean(c, s)[col] = mean;
        bnv_rstd(c, s)[col] = rstd;
        bnv_beta(c, s)[col] = beta;
    }
}

// =============================================================================================
// forward
// =============================================================================================
// One layer's normalisation + ReLU on Y (shared) -> H (shared), batch statistics through the grid
// barrier (H doubles as the scratch of the fold: it is written only afterwards).
__device__ __forceinline__ void mid_fwd_bn_relu(const MidCtx &c, const MidDesc &d, const MidLayer &l, int slot,
                                                const float *Y, float *H, bool training, int aug_end) {
    const int N = l.n_out;
    float *s_mean = bnv_mean(c, slot), *s_rstd = bnv_rstd(c, slot), *s_beta = bnv_beta(c, slot);
    if (l.beta && training) {
        mid_bn_partial(c, Y, N, ws_stat_slot(d, slot));
        mid_grid_sync(d.barrier, d.error);
        mid_bn_fold(c, d, l, ws_stat_slot(d, slot), H, s_mean, s_rstd);
    }
    for (int i = threadIdx.x; i < N * c.RP; i += kMidThreads) {
        const int col = i / c.RP, r = i - col * c.RP;
        float v = 0.f;
        if (r < c.nr) v = fmaxf((Y[i] - s_mean[col]) * s_rstd[col] + s_beta[col], 0.f);
        H[i] = v;
    }
    set_aug_cols(c, H, N, aug_end);
}

template <int TM>
__global__ void __launch_bounds__(kMidThreads, 1) vae_mid_fwd_kernel(const MidDesc d) {
    constexpr int SLAB_IT = MidTraits<TM>::kSlabIt;
    extern __shared__ __align__(16) float mid_smem[];
    MidCtx c;
    mid_setup(c, d, mid_smem);
    const bool training = d.training != 0;
    const int L = d.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *E = c.act[3];       // noise slab [l][r]

    // ---- prologue: every load that depends on no other CTA, in ONE round trip ------------------------
    // first-layer pre-activations (the tensor-core product's split-K partials, folded in fixed order:
    // this is that GEMM's epilogue), the noise, the normalisation vectors of every layer, and the
    // weights of the first product of this kernel.
    const MidLayer &l0 = d.enc[0];
    const int N0 = l0.n_out;
    const MidLayer &first = d.n_enc > 1 ? d.enc[1] : d.post;
    const int firstN = d.n_enc > 1 ? first.n_out : L;
    WRegs wr;
    stage_w_issue(wr, first.w, first.ldw, 0, firstN, first.k_in);
    mid_load_bnv(c, d, training ? 0 : 1);
    const bool use_eps = !d.deterministic;
    if (use_eps && d.generate_eps) {
        uint64_t offset = d.offset;
        if (d.offset_dev) offset += (uint64_t)(*d.offset_dev);
        const int64_t e0 = (int64_t)c.r0 * L, e1 = (int64_t)(c.r0 + c.nr) * L;
        for (int64_t q = (e0 >> 2) + threadIdx.x; q * 4 < e1; q += kMidThreads) {
            curandStatePhilox4_32_10_t st;            // same stream as scvae_fill_normal
            curand_init(d.seed, (unsigned long long)q, 4ull * offset, &st);
            const float4 v = curand_normal4(&st);
            const float vv[4] = {v.x, v.y, v.z, v.w};
            for (int j = 0; j < 4; ++j) {
                const int64_t e = q * 4 + j;
                if (e >= e0 && e < e1) {
                    d.eps[e] = vv[j];
                    E[(int)((e - e0) % L) * c.RP + (int)((e - e0) / L)] = vv[j];
                }
            }
        }
    }
    {
        float *Y = c.act[1];
        const int total = N0 * c.RP;
        const int etotal = (use_eps && !d.generate_eps) ? L * c.RP : 0;
        MID_SLAB(etotal, L, ev, (r < c.nr ? d.eps[(int64_t)(c.r0 + r) * L + col] : 0.f))
        // (split-K partials: one batch per slice, summed in slice order)
        MID_SLAB(total, N0, y, (r < c.nr ? __ldcg(d.y1_parts + (int64_t)(c.r0 + r) * d.y1_ld + col) : 0.f))
        for (int s = 1; s < d.y1_nsplit; ++s) {
            MID_SLAB(total, N0, part,
                     (r < c.nr ? __ldcg(d.y1_parts + (int64_t)s * d.y1_slice + (int64_t)(c.r0 + r) * d.y1_ld + col) : 0.f))
#pragma unroll
            for (int it = 0; it < SLAB_IT; ++it) y[it] += part[it];
        }
        MID_SLAB_FOR(total, N0, y, {
            const float val = x * d.y1_alpha;
            if (r < c.nr && l0.y) l0.y[(int64_t)(c.r0 + r) * l0.ldy + col] = val;
            Y[col * c.RP + r] = val;
        })
        MID_SLAB_FOR(etotal, L, ev, { E[col * c.RP + r] = x; })
    }
    stage_w_commit(c, wr, firstN, first.k_in);
    __syncthreads();

    int cur = 0;          // buffer holding the current activation H
    mid_fwd_bn_relu(c, d, l0, 0, c.act[1], c.act[0], training, ((N0 + 1 + 3) & ~3));
    __syncthreads();
    for (int i = 1; i < d.n_enc; ++i) {
        const MidLayer &l = d.enc[i];
        const int yb = (cur + 1) % 3, hb = (cur + 2) % 3;
        if (i > 1) {
            stage_w(c, l.w, l.ldw, 0, l.n_out, l.k_in);
            __syncthreads();
        }
        mid_product<TM>(c, c.act[cur], l.k_in, c.act[yb], l.n_out, l.y, l.ldy);
        __syncthreads();
        mid_fwd_bn_relu(c, d, l, i, c.act[yb], c.act[hb], training, ((l.n_out + 1 + 3) & ~3));
        cur = hb;
        __syncthreads();
    }
    // ---- posterior heads: [mu | log_sigma] = h W^T (no batch norm, VAE:2268-2289) ----------------
    const int mb = (cur + 1) % 3, lb = (cur + 2) % 3;
    {
        const MidLayer &l = d.post;
        for (int part = 0; part < 2; ++part) {
            if (part > 0 || d.n_enc > 1) {
                stage_w(c, l.w, l.ldw, part * L, L, l.k_in);
                __syncthreads();
            }
            mid_product<TM>(c, c.act[cur], l.k_in, c.act[part ? lb : mb], L, nullptr, 0);
            __syncthreads();
        }
    }
    // the decoder's first weights travel while the sample is formed
    const MidLayer &dl0 = d.dec[0];
    stage_w_issue(wr, dl0.w, dl0.ldw, 0, dl0.n_out, dl0.k_in);
    // ---- sample and KL (VAE:2353-2369, :2624-2627), one element per thread and pass ----------------
    {
        float *sMu = c.act[mb], *sLs = c.act[lb], *sZ = c.act[cur];
        const int Kzp = (dl0.k_in + 3) & ~3;
        for (int i = threadIdx.x; i < L * c.RP; i += kMidThreads) {
            const int l = i / c.RP, r = i - l * c.RP;
            float zv = 0.f, k = 0.f;
            if (r < c.nr) {
                const int64_t row = c.r0 + r;
                const float mu = sMu[i];
                const float raw = sLs[i];
                const float ls = fminf(fmaxf(raw, -3.f), 3.f);
                const float sigma = __expf(ls);
                // tfp kl_divergence(Normal(mu, sigma), Normal(0, 1))
                k = 0.5f * mu * mu + 0.5f * (sigma * sigma - 1.f) - ls;
                zv = d.deterministic ? mu : mu + sigma * E[i];
                if (d.kl_elem) d.kl_elem[row * L + l] = k;
                // HBM copies for the backward pass / the shells: [mu | raw log_sigma], z
                d.ph[row * d.ldph + l] = mu;
                d.ph[row * d.ldph + L + l] = raw;
                d.z[row * d.ldz + l] = zv;
            }
            sZ[i] = zv;
            sLs[i] = k;           // (log_sigma is consumed: its buffer now holds the KL terms)
        }
        // augmented column, decoder-input extras (VAE:2400-2441), zero padding
        for (int i = threadIdx.x; i < (Kzp - L) * c.RP; i += kMidThreads) {
            const int col = L + i / c.RP, r = i % c.RP;
            float v = 0.f;
            if (r < c.nr) {
                const int64_t row = c.r0 + r;
                if (col == L) v = 1.f;
                else if (d.batch_index && col - (L + 1) < d.n_batches)
                    v = ((int)d.batch_index[row] == col - (L + 1)) ? 1.f : 0.f;
                else if (d.count_sum && col == L + 1 + (d.batch_index ? d.n_batches : 0))
                    v = d.count_sum[row];
                if (col < d.ldz) d.z[row * d.ldz + col] = v;
            }
            sZ[col * c.RP + r] = v;
        }
        __syncthreads();
        for (int r = warp; r < c.nr; r += kMidThreads / 32) {      // per-cell KL, fixed order
            float kl = 0.f;
            for (int l = lane; l < L; l += 32) kl += sLs[l * c.RP + r];
            kl = warp_sum(kl);
            if (lane == 0 && d.kl_row) d.kl_row[c.r0 + r] = kl;
        }
    }
    stage_w_commit(c, wr, dl0.n_out, dl0.k_in);
    __syncthreads();
    // ---- decoder layers ---------------------------------------------------------------------------
    for (int j = 0; j < d.n_dec; ++j) {
        const MidLayer &l = d.dec[j];
        const int yb = (cur + 1) % 3, hb = (cur + 2) % 3;
        if (j > 0) {
            stage_w(c, l.w, l.ldw, 0, l.n_out, l.k_in);
            __syncthreads();
        }
        mid_product<TM>(c, c.act[cur], l.k_in, c.act[yb], l.n_out, l.y, l.ldy);
        __syncthreads();
        mid_fwd_bn_relu(c, d, l, d.n_enc + j, c.act[yb], c.act[hb], training, ((l.n_out + 1 + 3) & ~3));
        cur = hb;
        __syncthreads();
    }
    // ---- operand of the fused likelihood heads: fp16, augmented, zero padded to ldd16 columns ----
    {
        const int N = d.dec[d.n_dec - 1].n_out;
        const float *H = c.act[cur];
        const int groups = (int)(d.ldd16 >> 3);
        __half *out = reinterpret_cast<__half *>(d.d16);
        for (int i = threadIdx.x; i < c.nr * groups; i += kMidThreads) {
            const int r = i / groups, c0 = (i % groups) << 3;
            __align__(16) __half h[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) h[j] = __float2half_rn((c0 + j <= N) ? H[(c0 + j) * c.RP + r] : 0.f);   // column N = 1
            *reinterpret_cast<uint4 *>(out + (int64_t)(c.r0 + r) * d.ldd16 + c0) = *reinterpret_cast<const uint4 *>(h);
        }
        if (d.h_last) {
            float *hl = d.h_last;
            for (int i = threadIdx.x; i < c.nr * (int)d.ldh_last; i += kMidThreads) {
                const int r = i / (int)d.ldh_last, col = i % (int)d.ldh_last;
                hl[(int64_t)(c.r0 + r) * d.ldh_last + col] = col <= N ? H[col * c.RP + r] : 0.f;
            }
        }
    }
}

// =============================================================================================
// backward
// =============================================================================================
// Batched load of a layer's stored pre-activations -> X[c][r] = xhat (normalised) and / or
// H[c][r] = relu(xhat + beta), with the layer's (mean, rstd, beta) from shared-memory slot `s`.
template <int TM>
__device__ __forceinline__ void mid_recompute(const MidCtx &c, const MidLayer &l, int s, float *X, float *H) {
    constexpr int SLAB_IT = MidTraits<TM>::kSlabIt;
    const int N = l.n_out;
    const int total = N * c.RP;
    const float *m = bnv_mean(c, s), *rs = bnv_rstd(c, s), *bt = bnv_beta(c, s);
    MID_SLAB(total, N, y, (r < c.nr ? l.y[(int64_t)(c.r0 + r) * l.ldy + col] : 0.f))
    MID_SLAB_FOR(total, N, y, {
        float xh = 0.f;
        float h = 0.f;
        if (r < c.nr) {
            xh = (x - m[col]) * rs[col];
            h = fmaxf(xh + bt[col], 0.f);
        }
        if (X) X[col * c.RP + r] = xh;
        if (H) H[col * c.RP + r] = h;
    })
}

// In place on Gd[c][r] (d loss / d activation): ReLU mask, batch-norm backward.  X holds xhat.
// dy = rstd (g - s1/B - xhat s2/B);  dbeta = s1  (s1 = sum g, s2 = sum g xhat over ALL cells).
// `scratch`: [8 warps][2][128] floats of idle shared memory.
__device__ __forceinline__ void mid_bn_relu_bwd(const MidCtx &c, const MidDesc &d, const MidLayer &l, int s, float *Gd,
                                                const float *X, float *ws_stat, float *scratch) {
    const int N = l.n_out;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float *rs = bnv_rstd(c, s), *bt = bnv_beta(c, s);
    // g = dH * [H > 0], H = relu(xhat + beta)
    for (int i = threadIdx.x; i < N * c.RP; i += kMidThreads) {
        const int col = i / c.RP, r = i - col * c.RP;
        float g = 0.f;
        if (r < c.nr && X[i] + bt[col] > 0.f) g = Gd[i];
        Gd[i] = g;
    }
    __syncthreads();
    if (!l.beta) return;
    for (int col = warp; col < N; col += kMidThreads / 32) {
        float s1 = 0.f, s2 = 0.f;
        for (int r = lane; r < c.nr; r += 32) {
            const float g = Gd[col * c.RP + r];
            s1 += g;
            s2 += g * X[col * c.RP + r];
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (lane == 0) {
            ws_stat[((int64_t)blockIdx.x * 2 + 0) * kMidCols + col] = s1;
            ws_stat[((int64_t)blockIdx.x * 2 + 1) * kMidCols + col] = s2;
        }
    }
    mid_grid_sync(d.barrier, d.error);
    float *t1 = c.stat, *t2 = c.stat + kMidCols;
    {
        // fixed-order fold, loads batched as in the forward statistics
        const int G = gridDim.x;
        float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 10
        for (int k = warp; k < G; k += kMidThreads / 32) {
            float v1[4], v2[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int col = lane + 32 * q;
                v1[q] = col < N ? __ldcg(ws_stat + ((int64_t)k * 2 + 0) * kMidCols + col) : 0.f;
                v2[q] = col < N ? __ldcg(ws_stat + ((int64_t)k * 2 + 1) * kMidCols + col) : 0.f;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                a1[q] += v1[q];
                a2[q] += v2[q];
            }
        }
        float *x = scratch;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            x[(warp * 2 + 0) * kMidCols + lane + 32 * q] = a1[q];
            x[(warp * 2 + 1) * kMidCols + lane + 32 * q] = a2[q];
        }
        __syncthreads();
        const int col = threadIdx.x;
        if (col < N) {
            float s1 = 0.f, s2 = 0.f;
            for (int w = 0; w < kMidThreads / 32; ++w) {
                s1 += x[(w * 2 + 0) * kMidCols + col];
                s2 += x[(w * 2 + 1) * kMidCols + col];
            }
            t1[col] = s1;
            t2[col] = s2;
            if (blockIdx.x == 0) l.dbeta[col] = s1;
        }
    }
    __syncthreads();
    const float inv_b = 1.f / (float)d.B;
    for (int i = threadIdx.x; i < N * c.RP; i += kMidThreads) {
        const int col = i / c.RP, r = i - col * c.RP;
        float v = 0.f;
        if (r < c.nr) v = rs[col] * (Gd[i] - t1[col] * inv_b - X[i] * t2[col] * inv_b);
        Gd[i] = v;
    }
    __syncthreads();
}

// dIn[k][r] = sum_n G[n][r] W[row0 + n][k], k < n_in, accumulated over `parts` row blocks of W.
// `pre`: the weights of part 0, issued by the caller (in flight across the work before this call).
template <int TM>
__device__ __forceinline__ void mid_dgrad(const MidCtx &c, const MidLayer &l, const float *G0, const float *G1,
                                          int rows_per_part, int parts, float *Out, const WRegs &pre) {
    float acc[TM][4] = {};
    for (int p = 0; p < parts; ++p) {
        if (p == 0) {
            stage_w_commit(c, pre, rows_per_part, l.n_in);
        } else {
            __syncthreads();
            stage_w(c, l.w, l.ldw, p * rows_per_part, rows_per_part, l.n_in);
        }
        __syncthreads();
        mm_dgrad<TM>(p ? G1 : G0, c.RP, c.sw, rows_per_part, acc);
    }
    __syncthreads();            // every thread is done reading G before Out (which may alias) is written
    mm_dgrad_store<TM>(c, Out, l.n_in, acc);
}

template <int TM>
__global__ void __launch_bounds__(kMidThreads, 1) vae_mid_bwd_kernel(const MidDesc d) {
    constexpr int SLAB_IT = MidTraits<TM>::kSlabIt;
    extern __shared__ __align__(16) float mid_smem[];
    MidCtx c;
    mid_setup(c, d, mid_smem);
    const int L = d.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float weight = d.kl_weight;
    if (d.scalars) weight *= d.scalars[1];
    const float kl_coef = weight / (float)d.B;
    int slot = 0;          // batch-norm slot in the workspace
    float *dw_ws = ws_dw_base(d);
    // per-layer offset of the weight-gradient partials: [layer][cta][n_out * ldw]
    auto dw_part = [&](float *&cursor, const MidLayer &l) {
        float *mine = cursor + (int64_t)blockIdx.x * l.n_out * l.ldw;
        cursor += (int64_t)gridDim.x * l.n_out * l.ldw;
        return mine;
    };

    // ---- phase 0: decoder-output gradient and log p from the gene-range partials of the fused
    // heads kernel; log p gets the first-order correction for the fp16 rounding of its operand:
    // log p(d) ~ log p(d16) + (d - d16) . d log p / d d, and dd = go_scalar * d log p / d d.
    const MidLayer &last = d.dec[d.n_dec - 1];
    const int s_last = d.n_enc + d.n_dec - 1;
    float *Gd = c.act[0], *X = c.act[1], *T = c.act[2], *F = c.act[3];
    WRegs wr;
    mid_load_bnv(c, d, 2);
    {
        const int N = last.n_out;
        const int total = N * c.RP;
        const __half *d16 = reinterpret_cast<const __half *>(d.d16);
        // per-cell scalars: threads 0..nr-1
        float lp = 0.f, klr = 0.f;
        if ((int)threadIdx.x < c.nr) {
            const int64_t row = c.r0 + threadIdx.x;
            for (int s = 0; s < d.logp_nsplit; ++s) lp += __ldcg(d.logp_parts + (int64_t)s * d.logp_slice + row);
            if (d.row_const) lp -= d.row_const[row];
            klr = d.kl_row[row];
        }
        MID_SLAB(total, N, y, (r < c.nr ? last.y[(int64_t)(c.r0 + r) * last.ldy + col] : 0.f))
        MID_SLAB(total, N, h16, (r < c.nr ? __half2float(d16[(int64_t)(c.r0 + r) * d.ldd16 + col]) : 0.f))
        MID_SLAB(total, N, g, (r < c.nr ? __ldcg(d.dd_parts + (int64_t)(c.r0 + r) * d.dd_ld + col) : 0.f))
        for (int s = 1; s < d.dd_nsplit; ++s) {
            MID_SLAB(total, N, part,
                     (r < c.nr ? __ldcg(d.dd_parts + (int64_t)s * d.dd_slice + (int64_t)(c.r0 + r) * d.dd_ld + col) : 0.f))
#pragma unroll
            for (int it = 0; it < SLAB_IT; ++it) g[it] += part[it];
        }
        __syncthreads();         // the normalisation vectors are in shared memory
        const float *m = bnv_mean(c, s_last), *rs = bnv_rstd(c, s_last), *bt = bnv_beta(c, s_last);
#pragma unroll
        for (int it = 0; it < SLAB_IT; ++it) {
            const int i = threadIdx.x + it * kMidThreads;
            if (i < total) {
                const int col = i % N, r = i / N;
                float xh = 0.f, corr = 0.f;
                if (r < c.nr) {
                    xh = (y[it] - m[col]) * rs[col];
                    const float h = fmaxf(xh + bt[col], 0.f);      // fp32 activation the heads saw in fp16
                    corr = (h - h16[it]) * g[it];
                }
                Gd[col * c.RP + r] = g[it];
                X[col * c.RP + r] = xh;
                T[col * c.RP + r] = corr;
            }
        }
        __syncthreads();
        const float inv_go = 1.f / d.go_scalar;
        for (int r = warp; r < c.nr; r += kMidThreads / 32) {
            float corr = 0.f;
            for (int col = lane; col < N; col += 32) corr += T[col * c.RP + r];
            corr = warp_sum(corr) * inv_go;
            if (lane == 0) c.stat[2 * kMidCols + r] = corr;
        }
        __syncthreads();
        float lp_sum = 0.f, kl_sum = 0.f;
        if ((int)threadIdx.x < c.nr) {
            lp += c.stat[2 * kMidCols + threadIdx.x];
            d.logp[c.r0 + threadIdx.x] = lp;
            lp_sum = lp;
            kl_sum = klr;
        }
        // per-CTA partial of the bound (fixed order: lanes, warps, then CTAs at the end)
        lp_sum = warp_sum(lp_sum);
        kl_sum = warp_sum(kl_sum);
        if (lane == 0) {
            c.red[warp] = lp_sum;
            c.red[8 + warp] = kl_sum;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float a = 0.f, b = 0.f;
            for (int w = 0; w < kMidThreads / 32; ++w) {
                a += c.red[w];
                b += c.red[8 + w];
            }
            ws_bound(d)[blockIdx.x * 4 + 0] = a;
            ws_bound(d)[blockIdx.x * 4 + 1] = b;
        }
        __syncthreads();
    }
    // ---- decoder layers, last to first -------------------------------------------------------------
    // invariant at the top: Gd = d loss / d activation of layer j, X = its xhat; T, F free
    for (int j = d.n_dec - 1; j >= 0; --j) {
        const MidLayer &l = d.dec[j];
        const int s = d.n_enc + j;
        // issued ahead of the barrier inside the batch-norm backward: the layer's dgrad weights and
        // its input -- the previous decoder activation (recomputed, with its xhat for the next
        // iteration) or the latent sample
        float *In = T, *Xprev = F;
        const int Kp = (int)l.ldw;
        stage_w_issue(wr, l.w, l.ldw, 0, l.n_out, l.n_in);
        if (j > 0) {
            mid_recompute<TM>(c, d.dec[j - 1], s - 1, Xprev, In);
            set_aug_cols(c, In, d.dec[j - 1].n_out, Kp);
        } else {
            const int zc = min(Kp, (int)d.ldz);
            const int total = zc * c.RP;
            MID_SLAB(total, zc, zv, (r < c.nr ? d.z[(int64_t)(c.r0 + r) * d.ldz + col] : 0.f))
            MID_SLAB_FOR(total, zc, zv, { In[col * c.RP + r] = x; })
            if (zc < Kp) zero_cols(c, In, zc, Kp);
        }
        mid_bn_relu_bwd(c, d, l, s, Gd, X, ws_stat_slot(d, slot++), c.sw);
        mid_wgrad(c, Gd, l.n_out, In, Kp, dw_part(dw_ws, l));
        __syncthreads();
        float *Out = X;      // xhat of this layer is dead: the dgrad output takes its buffer
        mid_dgrad<TM>(c, l, Gd, Gd, l.n_out, 1, Out, wr);
        __syncthreads();
        // rotate: Gd <- Out, X <- Xprev, free: old Gd (-> F), In (T stays T)
        float *oldG = Gd;
        Gd = Out;
        X = Xprev;
        F = oldG;
    }
    // ---- sample / KL backward (VAE:2353-2369 and the analytic KL): Gd = dZ[l][r] -----------------
    // dmu = dz + c mu;  dlog_sigma = (dz eps sigma + c (sigma^2 - 1)) [|raw| <= 3];  c = weight / B
    float *Gmu = Gd, *Gls = X;
    const MidLayer &pl = d.post;
    const MidLayer &prev = d.enc[d.n_enc - 1];
    stage_w_issue(wr, pl.w, pl.ldw, 0, L, pl.n_in);
    {
        const int total = L * c.RP;
        MID_SLAB(total, L, mu, (r < c.nr ? d.ph[(int64_t)(c.r0 + r) * d.ldph + col] : 0.f))
        MID_SLAB(total, L, raw, (r < c.nr ? d.ph[(int64_t)(c.r0 + r) * d.ldph + L + col] : 0.f))
        MID_SLAB(total, L, ep, ((r < c.nr && !d.deterministic) ? d.eps[(int64_t)(c.r0 + r) * L + col] : 0.f))
#pragma unroll
        for (int it = 0; it < SLAB_IT; ++it) {
            const int i = threadIdx.x + it * kMidThreads;
            if (i < total) {
                const int col = i % L, r = i / L;
                float gm = 0.f, gl = 0.f;
                if (r < c.nr) {
                    const float ls = fminf(fmaxf(raw[it], -3.f), 3.f);
                    const float sigma = __expf(ls);
                    const float dz = Gd[col * c.RP + r];
                    gm = dz + kl_coef * mu[it];
                    const float mask = (raw[it] < -3.f || raw[it] > 3.f) ? 0.f : 1.f;
                    gl = (dz * ep[it] * sigma + kl_coef * (sigma * sigma - 1.f)) * mask;
                }
                Gmu[col * c.RP + r] = gm;      // (the element this thread just read)
                Gls[col * c.RP + r] = gl;
            }
        }
    }
    // ---- posterior heads: wgrad of both halves, dgrad summed over them ----------------------------
    {
        float *In = T;
        const int Kp = (int)pl.ldw;
        mid_recompute<TM>(c, prev, d.n_enc - 1, F, In);      // F = xhat of the last encoder layer (kept)
        set_aug_cols(c, In, prev.n_out, Kp);
        __syncthreads();
        float *mine = dw_part(dw_ws, pl);
        mid_wgrad(c, Gmu, L, In, Kp, mine);
        mid_wgrad(c, Gls, L, In, Kp, mine + (int64_t)L * Kp);
        __syncthreads();
        mid_dgrad<TM>(c, pl, Gmu, Gls, L, 2, In, wr);           // d loss / d H of the last encoder layer
        __syncthreads();
        // Gd <- In (T's buffer); X <- F (xhat); free: Gmu's and Gls's buffers
        float *f0 = Gmu, *f1 = Gls;
        Gd = In;
        X = F;
        T = f0;
        F = f1;
    }
    // ---- encoder layers, last to second; the first one ends in fp16 dY1 for the big wgrad ----------
    for (int i = d.n_enc - 1; i >= 0; --i) {
        const MidLayer &l = d.enc[i];
        float *In = T, *Xprev = F;
        const int Kp = (int)l.ldw;
        if (i > 0) {
            stage_w_issue(wr, l.w, l.ldw, 0, l.n_out, l.n_in);
            mid_recompute<TM>(c, d.enc[i - 1], i - 1, Xprev, In);
            set_aug_cols(c, In, d.enc[i - 1].n_out, Kp);
        }
        mid_bn_relu_bwd(c, d, l, i, Gd, X, ws_stat_slot(d, slot++), c.sw);
        if (i == 0) break;
        mid_wgrad(c, Gd, l.n_out, In, Kp, dw_part(dw_ws, l));
        __syncthreads();
        float *Out = X;
        mid_dgrad<TM>(c, l, Gd, Gd, l.n_out, 1, Out, wr);
        __syncthreads();
        float *oldG = Gd;
        Gd = Out;
        X = Xprev;
        F = oldG;
    }
    // dY1 (B, H1) -> fp16 + rounding remainder, scaled into range, zero padded: operand of dW1 = dY1^T X
    {
        const int N = d.enc[0].n_out;
        const int groups = (int)(d.lddy1 >> 3);
        __half *out = reinterpret_cast<__half *>(d.dy1_16);
        __half *out_lo = reinterpret_cast<__half *>(d.dy1_16_lo);
        for (int i = threadIdx.x; i < c.nr * groups; i += kMidThreads) {
            const int r = i / groups, c0 = (i % groups) << 3;
            __align__(16) __half h[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float v = (c0 + j < N) ? Gd[(c0 + j) * c.RP + r] * d.dy1_scale : 0.f;
                h[j] = __float2half_rn(v);
                lo[j] = __float2half_rn(v - __half2float(h[j]));     // rounding remainder
            }
            *reinterpret_cast<uint4 *>(out + (int64_t)(c.r0 + r) * d.lddy1 + c0) = *reinterpret_cast<const uint4 *>(h);
            if (out_lo)
                *reinterpret_cast<uint4 *>(out_lo + (int64_t)(c.r0 + r) * d.lddy1 + c0) = *reinterpret_cast<const uint4 *>(lo);
        }
        if (d.dy1) {
            for (int i = threadIdx.x; i < c.nr * N; i += kMidThreads) {
                const int r = i / N, col = i % N;
                d.dy1[(int64_t)(c.r0 + r) * d.lddy1_f32 + col] = Gd[col * c.RP + r];
            }
        }
    }
    // ---- all partials are in the workspace: fold them in fixed order --------------------------------
    mid_grid_sync(d.barrier, d.error);
    {
        float *cursor = ws_dw_base(d);
        const int G = gridDim.x;
        auto reduce = [&](const MidLayer &l) {
            const int64_t n = (int64_t)l.n_out * l.ldw;
            for (int64_t e = (int64_t)blockIdx.x * kMidThreads + threadIdx.x; e < n; e += (int64_t)gridDim.x * kMidThreads) {
                float s = 0.f;
                int k = 0;
                for (; k + 16 <= G; k += 16) {        // 16 loads in flight, summed in CTA order
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = __ldcg(cursor + (int64_t)(k + j) * n + e);
#pragma unroll
                    for (int j = 0; j < 16; ++j) s += v[j];
                }
                for (; k < G; ++k) s += __ldcg(cursor + (int64_t)k * n + e);
                l.dw[e] = s;
            }
            cursor += (int64_t)gridDim.x * n;
        };
        for (int j = d.n_dec - 1; j >= 0; --j) reduce(d.dec[j]);
        reduce(d.post);
        for (int i = d.n_enc - 1; i >= 1; --i) reduce(d.enc[i]);
        if (blockIdx.x == 0) {
            for (int k = threadIdx.x; k < G; k += kMidThreads) {
                c.stat[k] = __ldcg(ws_bound(d) + k * 4 + 0);
                c.stat[kMidThreads + k] = __ldcg(ws_bound(d) + k * 4 + 1);
            }
            __syncthreads();
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            // lower bound (R = 1): mean_b(log p - KL); weighted; ENRE; KL (VAE:2715-2734)
            float lp = 0.f, kl = 0.f;
            for (int k = 0; k < G; ++k) {
                lp += c.stat[k];
                kl += c.stat[kMidThreads + k];
            }
            const float inv_b = 1.f / (float)d.B;
            d.bound[0] = (lp - kl) * inv_b;
            d.bound[1] = (lp - weight * kl) * inv_b;
            d.bound[2] = lp * inv_b;
            d.bound[3] = kl * inv_b;
        }
    }
}

static int mid_check(const MidDesc *d, const char *name, bool bwd) {
    SCVAE_CHECK_ARG(d, "%s: NULL descriptor", name);
    SCVAE_CHECK_ARG(d->B > 0 && d->L > 0 && d->L <= kMidCols, "%s: bad B / L", name);
    SCVAE_CHECK_ARG(d->n_enc >= 1 && d->n_enc <= SCVAE_MID_MAX_LAYERS && d->n_dec >= 1 && d->n_dec <= SCVAE_MID_MAX_LAYERS,
                    "%s: 1..%d encoder and decoder layers", name, SCVAE_MID_MAX_LAYERS);
    SCVAE_CHECK_ARG(d->rows_per_cta > 0 && d->rows_per_cta <= kMidRows, "%s: rows_per_cta must be in 1..%d", name, kMidRows);
    SCVAE_CHECK_ARG(d->workspace && d->barrier && d->error, "%s: workspace / barrier / error are required", name);
    auto ok = [&](const MidLayer &l, bool first) {
        return l.w && l.ldw % 4 == 0 && l.ldw <= kMidCols && l.n_out > 0 && l.n_out < kMidCols && l.k_in <= l.ldw &&
               (first || l.k_in > l.n_in) && l.y && l.ldy % 4 == 0 && (!l.beta || (l.mean && l.rstd && l.moving_mean && l.moving_var));
    };
    for (int i = 0; i < d->n_enc; ++i) {
        // (the first encoder layer's weight is not touched here: only its outputs)
        MidLayer l = d->enc[i];
        if (i == 0) { l.ldw = 4; l.k_in = 1; l.n_in = 0; l.w = (const float *)d->workspace; }
        SCVAE_CHECK_ARG(ok(l, i == 0), "%s: encoder layer %d does not fit the fused middle (widths < 128)", name, i + 1);
    }
    for (int j = 0; j < d->n_dec; ++j)
        SCVAE_CHECK_ARG(ok(d->dec[j], false), "%s: decoder layer %d does not fit the fused middle", name, j + 1);
    SCVAE_CHECK_ARG(d->post.w && d->post.ldw % 4 == 0 && d->post.ldw <= kMidCols && d->post.n_out == 2 * d->L,
                    "%s: bad posterior layer", name);
    SCVAE_CHECK_ARG(d->ph && d->z && d->kl_row && d->d16 && d->ldd16 % 8 == 0 && d->ldz % 4 == 0, "%s: bad buffers", name);
    SCVAE_CHECK_ARG(d->deterministic || d->eps, "%s: eps is NULL", name);
    SCVAE_CHECK_ARG(((d->dec[0].k_in + 3) & ~3) <= kMidCols && d->ldz <= kMidCols, "%s: bad latent width", name);
    if (!bwd) {
        SCVAE_CHECK_ARG(d->y1_parts && d->y1_nsplit >= 1, "%s: first-layer product missing", name);
    } else {
        SCVAE_CHECK_ARG(d->dd_parts && d->dd_nsplit >= 1 && d->logp_parts && d->logp_nsplit >= 1 && d->logp && d->bound &&
                            d->dy1_16 && d->lddy1 % 8 == 0 && d->go_scalar != 0.f,
                        "%s: bad backward buffers", name);
        for (int i = 1; i < d->n_enc; ++i) SCVAE_CHECK_ARG(d->enc[i].dw, "%s: dw missing", name);
        for (int j = 0; j < d->n_dec; ++j) SCVAE_CHECK_ARG(d->dec[j].dw, "%s: dw missing", name);
        SCVAE_CHECK_ARG(d->post.dw, "%s: dw missing", name);
    }
    return 0;
}

static int mid_grid(const MidDesc *d) { return (d->B + d->rows_per_cta - 1) / d->rows_per_cta; }

static int64_t mid_workspace_floats_impl(const MidDesc *d) {
    const int64_t grid = mid_grid(d);
    int64_t n = (int64_t)kMidSlots * grid * 2 * kMidCols + grid * 4;
    for (int j = 0; j < d->n_dec; ++j) n += grid * d->dec[j].n_out * d->dec[j].ldw;
    n += grid * d->post.n_out * d->post.ldw;
    for (int i = 1; i < d->n_enc; ++i) n += grid * d->enc[i].n_out * d->enc[i].ldw;
    return n;
}

template <bool BWD, int TM>
static int mid_launch_t(const char *name, const MidDesc *d, int grid, cudaStream_t s) {
    const int smem = mid_smem_floats(d->rows_per_cta) * 4;
    auto kern = BWD ? vae_mid_bwd_kernel<TM> : vae_mid_fwd_kernel<TM>;
    static int have = 0;          // per kernel instantiation: largest size opted in so far
    if (smem > have) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        SCVAE_CHECK_ARG(e == cudaSuccess, "%s: cannot set smem attribute (%d bytes): %s", name, smem, cudaGetErrorString(e));
        have = smem;
    }
    kern<<<grid, kMidThreads, smem, s>>>(*d);
    SCVAE_CHECK_LAUNCH(name);
    return 0;
}

template <bool BWD>
static int mid_launch(const char *name, const MidDesc *d, cudaStream_t s) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = mid_grid(d);
    SCVAE_CHECK_ARG(grid <= sms, "%s: %d cells need %d CTAs of %d cells, the device has %d SMs (grid barrier)", name, d->B,
                    grid, d->rows_per_cta, sms);
    SCVAE_CHECK_ARG(d->workspace_floats >= mid_workspace_floats_impl(d), "%s: workspace too small", name);
    return d->rows_per_cta <= 32 ? mid_launch_t<BWD, 4>(name, d, grid, s) : mid_launch_t<BWD, 8>(name, d, grid, s);
}

}  // namespace scvae

using namespace scvae;

extern "C" int64_t scvae_vae_mid_workspace_floats(const scvae_mid_desc *d) {
    if (!d || d->B <= 0 || d->rows_per_cta <= 0) return 0;
    return mid_workspace_floats_impl(d);
}

extern "C" int scvae_vae_mid_fwd(const scvae_mid_desc *d, void *stream) {
    if (mid_check(d, "vae_mid_fwd", false)) return 1;
    return mid_launch<false>("vae_mid_fwd", d, (cudaStream_t)stream);
}

extern "C" int scvae_vae_mid_bwd(const scvae_mid_desc *d, void *stream) {
    if (mid_check(d, "vae_mid_bwd", true)) return 1;
    return mid_launch<true>("vae_mid_bwd", d, (cudaStream_t)stream);
}
