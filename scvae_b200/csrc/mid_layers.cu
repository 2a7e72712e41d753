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
constexpr int kMidSlots = 2 * SCVAE_MID_MAX_LAYERS;                // batch-normed layers (encoder + decoder)
constexpr int kMidFoldIt = 20;                                      // CTAs per warp in a fold: grid <= 160
// 240 registers per thread (256 threads -> 61 440 of the SM's 65 536): one 128-thread / 32-register CTA of
// the streamed feeder's pull kernel (gather.cu) fits beside a resident middle-kernel CTA, exactly as it
// does beside the fused heads kernel (640 x 96 registers) -- the pull of the next minibatch is hidden
// behind the step only while its CTAs can be resident
constexpr int kMidMaxRegs = 240;
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

// development aid: phase time stamps (ns) of every CTA, written only when d.timeline is given
__device__ __forceinline__ void mid_stamp(const MidDesc &d, int k) {
    if (d.timeline && threadIdx.x == 0 && k < 32) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        d.timeline[(int64_t)blockIdx.x * 32 + k] = (long long)t;
    }
}

__device__ __forceinline__ int mid_rows_of(int cta, int rows_per_cta, int B) {
    const int lo = cta * rows_per_cta;
    return max(0, min(rows_per_cta, B - lo));
}

// ---- compact code first ------------------------------------------------------------------------------
// These kernels run every phase ONCE: straight-line, fully unrolled code is fetched from memory at a
// few bytes per cycle (the first version spent 60 % of its cycles in `no instruction` stalls), so
// everything below is rolled loops over a division-free 2-D index space, shared helper functions
// (not inlined), and asynchronous copies (cp.async) for every global -> shared transfer: all of a
// phase's loads are in flight together without holding registers, and they can be issued ahead of a
// grid barrier or another phase's arithmetic.
// The activation / weight buffers are reached through pointers kept in a shared-memory context, so the
// compiler sees generic addresses: without this hint every operand load of the product loops is a
// 64-bit generic LD.E.128 with its descriptor moves instead of an LDS.128 with a 32-bit address.
#define MID_SHARED(p) __builtin_assume(__isShared(p))

template <int TM>
struct MidTraits {
    static constexpr int kSpan = TM * 8;          // rows covered by a product tile: 32 or 64
};

__device__ __forceinline__ uint32_t mid_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// 4-byte asynchronous copy; !valid writes zero without touching global memory
__device__ __forceinline__ void cp_async4(float *dst, const float *src, bool valid) {
    const int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(mid_smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async16(float *dst, const float *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(mid_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// dst[col][r] <- src[r0 + r][col] for col < ncol, r < span; zero for cells beyond this CTA's.
// Lanes run along the column index (coalesced global reads), warps along the cells.
__device__ __noinline__ void slab_copy_async(const MidCtx &c, float *dst, const float *__restrict__ src, int64_t ld,
                                             int ncol, int span) {
    MID_SHARED(&c); MID_SHARED(dst);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < span; r += kMidThreads / 32) {
        const bool valid = r < c.nr;
        const float *row = src + (int64_t)(c.r0 + (valid ? r : 0)) * ld;
        for (int col = lane; col < ncol; col += 32) cp_async4(dst + col * c.RP + r, row + col, valid);
    }
}
// dst[col][r] = alpha * sum_s src_s[r0 + r][col] (s < nsplit slices, `slice` floats apart), fixed order
__device__ __noinline__ void slab_sum(const MidCtx &c, float *dst, const float *__restrict__ src, int64_t ld, int ncol,
                                      int span, int nsplit, int64_t slice, float alpha) {
    MID_SHARED(&c); MID_SHARED(dst);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < span; r += kMidThreads / 32) {
        const bool valid = r < c.nr;
        const float *row = src + (int64_t)(c.r0 + (valid ? r : 0)) * ld;
        for (int col = lane; col < ncol; col += 32) {
            float v = 0.f;
            if (valid) {
#pragma unroll 4
                for (int s = 0; s < nsplit; ++s) v += __ldcg(row + (int64_t)s * slice + col);
            }
            dst[col * c.RP + r] = v * alpha;
        }
    }
}
// HBM[r0 + r][col] <- src[col][r] (row-major copy of a shared activation, valid cells only)
__device__ __noinline__ void slab_store(const MidCtx &c, const float *src, float *__restrict__ dst, int64_t ld, int ncol) {
    MID_SHARED(&c); MID_SHARED(src);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < c.nr; r += kMidThreads / 32)
        for (int col = lane; col < ncol; col += 32) dst[(int64_t)(c.r0 + r) * ld + col] = src[col * c.RP + r];
}

// ---- weight staging: sw[n][k] = W[row0 + n][k], n < N, k < round4(Kc) (natural layout, async) ---------
// (columns beyond the reduction length are zero in the stored weights -- forward -- or unused -- dgrad)
__device__ __noinline__ void stage_w_async(const MidCtx &c, const float *__restrict__ W, int64_t ldw, int row0, int N, int Kc) {
    MID_SHARED(&c);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int K4 = (Kc + 3) >> 2;
    for (int n = warp; n < N; n += kMidThreads / 32)
        for (int k4 = lane; k4 < K4; k4 += 32)
            cp_async16(c.sw + n * kMidWP + 4 * k4, W + (int64_t)(row0 + n) * ldw + 4 * k4);   // ldw % 4 == 0
}

// ---- forward product: C[n][r] = sum_k A[k][r] * sw[n][k], n < N; optional HBM copy Y[r0 + r][n] -------
// thread (ty = warp, tx = lane): rows ty TM + i, columns tx + 32 j.  The 8 lanes of a 128-bit phase
// read 8 consecutive rows of sw, 4 mod 32 floats apart: conflict free.  A must be finite (zero) on
// the rows [K, round4(K)).
template <int TM>
__device__ __noinline__ void mid_product(const MidCtx &c, const float *__restrict__ A, int K, float *__restrict__ C, int N,
                                         float *__restrict__ Y, int64_t ldy) {
    MID_SHARED(&c);
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
    const float *a = A + ty * TM;
    const float *sw = c.sw;
    MID_SHARED(a); MID_SHARED(sw); MID_SHARED(C);
    const int K4 = (K + 3) & ~3;
    const int RP = c.RP;           // (kept in a register: c lives in shared memory and would be re-read)
    // rows of sw this thread reads: clamped, so that the loads are unconditional (no branch per load);
    // the products of the clamped duplicates are simply not stored
    int wrow[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) wrow[j] = min(tx + 32 * j, N - 1) * kMidWP;
    float acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 1
    for (int k = 0; k < K4; k += 4) {
        float4 wv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) wv[j] = *reinterpret_cast<const float4 *>(sw + wrow[j] + k);
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
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = tx + 32 * j;
        if (n < N) {
#pragma unroll
            for (int i = 0; i < TM; i += 4)
                *reinterpret_cast<float4 *>(C + n * RP + ty * TM + i) =
                    make_float4(acc[i][j], acc[i + 1][j], acc[i + 2][j], acc[i + 3][j]);
            if (Y) {
#pragma unroll
                for (int i = 0; i < TM; ++i)
                    if (ty * TM + i < c.nr) Y[(int64_t)(c.r0 + ty * TM + i) * ldy + n] = acc[i][j];
            }
        }
    }
}

// ---- dgrad product: Out[k][r] (+)= sum_{n < Nred} G[n][r] * sw[n][k], k < Kout -------------------------
// thread (ty, tx): rows ty TM + i, columns tx 4 + j.  first: overwrite, else accumulate into Out.
template <int TM>
__device__ __noinline__ void mid_dgrad_tile(const MidCtx &c, const float *__restrict__ G, int Nred, float *__restrict__ Out,
                                            int Kout, bool first) {
    MID_SHARED(&c);
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
    const float *a = G + ty * TM;
    const float *w = c.sw + tx * 4;
    MID_SHARED(a); MID_SHARED(w); MID_SHARED(Out);
    const int RP = c.RP;
    float acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    if (tx * 4 < Kout) {
#pragma unroll 2
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
    __syncthreads();            // every thread is done reading G before Out (which may alias it) is written
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int k = tx * 4 + j;
        if (k < Kout) {
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                float4 *dst = reinterpret_cast<float4 *>(Out + k * RP + ty * TM + i);
                float4 v = make_float4(acc[i][j], acc[i + 1][j], acc[i + 2][j], acc[i + 3][j]);
                if (!first) {
                    const float4 o = *dst;
                    v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                }
                *dst = v;
            }
        }
    }
}

// ---- weight-gradient partial of this CTA: out[n][k] = sum_{r < nr} G[n][r] * I[k][r] ----------
// thread (ty, tx) of a 16 x 16 grid owns n = ty + 16 i, k = tx + 16 j.  out: (N, Kp) row-major
// in the workspace (Kp = padded input width = leading dimension of the weight).
template <int NJ>
__device__ __noinline__ void mid_wgrad_t(const MidCtx &c, const float *__restrict__ G, int N, const float *__restrict__ I,
                                         int Kp, float *__restrict__ out) {
    MID_SHARED(&c);
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const int ni = (N - ty + 15) >> 4, kj = (Kp - tx + 15) >> 4;    // valid i / j counts
    MID_SHARED(G); MID_SHARED(I);
    const int RP = c.RP, nr4 = c.nr4;
    // clamped rows: unconditional loads (the duplicates' sums are not stored)
    int grow[8], irow[NJ];
#pragma unroll
    for (int i = 0; i < 8; ++i) grow[i] = min(ty + 16 * i, N - 1) * RP;
#pragma unroll
    for (int j = 0; j < NJ; ++j) irow[j] = min(tx + 16 * j, Kp - 1) * RP;
    float acc[8][NJ];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;
#pragma unroll 1
    for (int r = 0; r < nr4; r += 4) {
        float4 g[8], x[NJ];
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = *reinterpret_cast<const float4 *>(G + grow[i] + r);
#pragma unroll
        for (int j = 0; j < NJ; ++j) x[j] = *reinterpret_cast<const float4 *>(I + irow[j] + r);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j)
                acc[i][j] = fmaf(g[i].x, x[j].x, fmaf(g[i].y, x[j].y, fmaf(g[i].z, x[j].z, fmaf(g[i].w, x[j].w, acc[i][j]))));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j)
            if (i < ni && j < kj) out[(int64_t)(ty + 16 * i) * Kp + tx + 16 * j] = acc[i][j];
}
__device__ __forceinline__ void mid_wgrad(const MidCtx &c, const float *G, int N, const float *I, int Kp, float *out) {
    if (Kp <= 64) mid_wgrad_t<4>(c, G, N, I, Kp, out);      // narrow inputs (the latent sample): half the tile
    else mid_wgrad_t<8>(c, G, N, I, Kp, out);
}

// augmented ones column (valid cells only) at `col`, zero columns behind it up to `col_end`
__device__ __noinline__ void set_aug_cols(const MidCtx &c, float *dst, int col, int col_end, int span) {
    MID_SHARED(&c); MID_SHARED(dst);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int cc = col + warp; cc < col_end; cc += kMidThreads / 32)
        for (int r = lane; r < span; r += 32) dst[cc * c.RP + r] = (cc == col && r < c.nr) ? 1.f : 0.f;
}

// Per-CTA partial statistics of Y[c][r] (columns < N over this CTA's cells): (mean, M2) -> workspace.
// One thread per (column, half of the cells); the two halves meet in shared memory in fixed order.
__device__ __noinline__ void mid_bn_partial(const MidCtx &c, const float *Y, int N, float *ws_stat) {
    MID_SHARED(&c); MID_SHARED(Y);
    // one pass: sums of (y - pivot) and (y - pivot)^2 with the column's first cell as the pivot (a
    // value within the spread of the data, so neither sum cancels): mean = pivot + s1 / n,
    // M2 = s2 - s1^2 / n
    const int col = threadIdx.x & (kMidCols - 1), half = threadIdx.x >> 7;
    const int h0 = half ? (c.nr >> 1) : 0, h1 = half ? c.nr : (c.nr >> 1);
    float *x = c.stat + 4 * kMidCols;        // [2 halves][2][128]
    float s1 = 0.f, s2 = 0.f;
    const float pivot = col < N ? Y[col * c.RP] : 0.f;
    if (col < N) {
#pragma unroll 4
        for (int r = h0; r < h1; ++r) {
            const float dv = Y[col * c.RP + r] - pivot;
            s1 += dv;
            s2 = fmaf(dv, dv, s2);
        }
    }
    x[(half * 2 + 0) * kMidCols + col] = s1;
    x[(half * 2 + 1) * kMidCols + col] = s2;
    __syncthreads();
    if (half == 0 && col < N) {
        const float t1 = x[col] + x[2 * kMidCols + col], t2 = x[kMidCols + col] + x[3 * kMidCols + col];
        const float inv_n = c.nr > 0 ? 1.f / (float)c.nr : 0.f;
        ws_stat[((int64_t)blockIdx.x * 2 + 0) * kMidCols + col] = pivot + t1 * inv_n;
        ws_stat[((int64_t)blockIdx.x * 2 + 1) * kMidCols + col] = fmaxf(t2 - t1 * t1 * inv_n, 0.f);
    }
}
// The fold for grids of at most 8 CTAs: thread t < 128 owns column t and adds the <= 8 partials in CTA
// order (the same arithmetic as the general fold, whose 8 warp lanes then hold one partial each).
__device__ __noinline__ void mid_bn_fold_small(const MidCtx &c, const MidDesc &d, const MidLayer &l, const float *ws_stat,
                                               float *scratch, float *s_mean, float *s_rstd) {
    MID_SHARED(&c); MID_SHARED(&d); MID_SHARED(s_mean); MID_SHARED(s_rstd);
    (void)scratch;
    const int N = l.n_out, G = gridDim.x, col = threadIdx.x;
    const float inv_b = 1.f / (float)d.B;
    if (col < N) {
        float mk[kMidThreads / 32], qk[kMidThreads / 32];
#pragma unroll
        for (int k = 0; k < kMidThreads / 32; ++k) {
            mk[k] = k < G ? __ldcg(ws_stat + ((int64_t)k * 2 + 0) * kMidCols + col) : 0.f;
            qk[k] = k < G ? __ldcg(ws_stat + ((int64_t)k * 2 + 1) * kMidCols + col) : 0.f;
        }
        float m = 0.f;
#pragma unroll
        for (int k = 0; k < kMidThreads / 32; ++k) m += fmaf((float)mid_rows_of(k, d.rows_per_cta, d.B), mk[k], 0.f);
        const float mean_all = m * inv_b;
        float q_all = 0.f;
#pragma unroll
        for (int k = 0; k < kMidThreads / 32; ++k) {
            const float nb = (float)mid_rows_of(k, d.rows_per_cta, d.B);
            const float dm = mk[k] - mean_all;
            q_all += qk[k] + nb * dm * dm;
        }
        const float var = q_all * inv_b;
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
// Every CTA folds all partials in the same fixed order.  With n_k cells, mean m_k and M2 q_k per CTA:
//   mean = sum n_k m_k / B,   M2 = sum q_k + sum n_k (m_k - mean)^2     (exact, no cancellation),
// each as plain sums: warp w takes the CTAs k = w mod 8 for 4 columns per lane, then the 8 per-warp
// sums are added in warp order.  `scratch`: [8 warps][128] floats of idle shared memory.  Result:
// s_mean / s_rstd [N]; CTA 0 writes the saved statistics and the moving averages.
__device__ __noinline__ void mid_bn_fold(const MidCtx &c, const MidDesc &d, const MidLayer &l, const float *ws_stat,
                                         float *scratch, float *s_mean, float *s_rstd) {
    MID_SHARED(&c); MID_SHARED(&d); MID_SHARED(scratch); MID_SHARED(s_mean); MID_SHARED(s_rstd);
    const int N = l.n_out;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x;
    const float inv_b = 1.f / (float)d.B;
    // (all loads of a pass are issued before the first use: the fold costs two L2 round trips)
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float mb[kMidFoldIt][4];
    if (G <= kMidThreads / 32) {
        // small grids (minibatches of a few hundred cells): one partial per warp, none of the
        // unrolled load sequence below is fetched or issued
        mid_bn_fold_small(c, d, l, ws_stat, scratch, s_mean, s_rstd);
        return;
    }
#pragma unroll
    for (int it = 0; it < kMidFoldIt; ++it) {
        const int k = warp + it * (kMidThreads / 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int col = lane + 32 * q;
            mb[it][q] = (k < G && col < N) ? __ldcg(ws_stat + ((int64_t)k * 2 + 0) * kMidCols + col) : 0.f;
        }
    }
#pragma unroll
    for (int it = 0; it < kMidFoldIt; ++it) {
        const int k = warp + it * (kMidThreads / 32);
        const float nb = (float)mid_rows_of(k, d.rows_per_cta, d.B);
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] = fmaf(nb, mb[it][q], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) scratch[warp * kMidCols + lane + 32 * q] = acc[q];
    __syncthreads();
    if (threadIdx.x < kMidCols) {
        float m = 0.f;
        for (int w = 0; w < kMidThreads / 32; ++w) m += scratch[w * kMidCols + threadIdx.x];
        s_mean[threadIdx.x] = m * inv_b;
    }
    __syncthreads();
    float qb[kMidFoldIt][4];
#pragma unroll
    for (int it = 0; it < kMidFoldIt; ++it) {
        const int k = warp + it * (kMidThreads / 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int col = lane + 32 * q;
            qb[it][q] = (k < G && col < N) ? __ldcg(ws_stat + ((int64_t)k * 2 + 1) * kMidCols + col) : 0.f;
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = 0.f;
#pragma unroll
    for (int it = 0; it < kMidFoldIt; ++it) {
        const int k = warp + it * (kMidThreads / 32);
        const float nb = (float)mid_rows_of(k, d.rows_per_cta, d.B);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float dm = mb[it][q] - s_mean[lane + 32 * q];
            acc[q] += qb[it][q] + nb * dm * dm;          // (nb == 0 beyond the grid)
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) scratch[warp * kMidCols + lane + 32 * q] = acc[q];
    __syncthreads();
    const int col = threadIdx.x;
    if (col < N) {
        float q_all = 0.f;
        for (int w = 0; w < kMidThreads / 32; ++w) q_all += scratch[w * kMidCols + col];
        const float var = q_all * inv_b;
        const float rstd = rsqrtf(var + kBnEps);
        s_rstd[col] = rstd;
        if (blockIdx.x == 0) {
            const float mean_all = s_mean[col];
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

// The descriptor (1.3 KB of kernel parameters) and the context are copied into SHARED memory once per
// CTA: the helper functions take them by reference, and a reference to a kernel parameter would make
// every thread spill the whole structure to local memory at kernel entry (53 MB of traffic per launch).
struct MidShared {
    MidDesc d;
    MidCtx c;
};
__device__ __forceinline__ void mid_setup(MidShared &sh, const MidDesc &param, float *smem) {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(&param);
    uint32_t *dst = reinterpret_cast<uint32_t *>(&sh.d);
    for (int i = threadIdx.x; i < (int)(sizeof(MidDesc) / 4); i += kMidThreads) dst[i] = src[i];
    if (threadIdx.x == 0) {
        MidCtx &c = sh.c;
        const int rows = param.rows_per_cta;
        c.RP = mid_pitch(rows);
        for (int b = 0; b < 4; ++b) c.act[b] = smem + b * kMidCols * c.RP;
        c.sw = c.act[3] + kMidCols * c.RP;
        c.stat = c.sw + kMidCols * kMidWP;
        c.bnv = c.stat + 8 * kMidCols;
        c.red = c.bnv + kMidSlots * 3 * kMidCols;
        c.r0 = blockIdx.x * rows;
        c.nr = mid_rows_of(blockIdx.x, rows, param.B);
        c.nr4 = (c.nr + 3) & ~3;
    }
    __syncthreads();
}

// workspace map (floats): [BN slots][grid][2][128] | [grid][4] bound partials | dW partials
__device__ __forceinline__ float *ws_stat_slot(const MidDesc &d, int slot) {
    return d.workspace + (int64_t)slot * gridDim.x * 2 * kMidCols;
}
__device__ __forceinline__ float *ws_bound(const MidDesc &d) {
    return d.workspace + (int64_t)kMidSlots * gridDim.x * 2 * kMidCols;
}
__device__ __forceinline__ float *ws_dw_base(const MidDesc &d) { return ws_bound(d) + (int64_t)gridDim.x * 4; }
// [grid][128] column sums of dY1: the last block of the workspace (behind the dW partials)
__device__ __forceinline__ float *ws_db1(const MidDesc &d) {
    return d.workspace + d.workspace_floats - (int64_t)gridDim.x * kMidCols;
}

// mean / rstd / beta of layer slot `s` (encoder layers first, then decoder layers) in shared memory
__device__ __forceinline__ float *bnv_mean(const MidCtx &c, int s) { return c.bnv + (s * 3 + 0) * kMidCols; }
__device__ __forceinline__ float *bnv_rstd(const MidCtx &c, int s) { return c.bnv + (s * 3 + 1) * kMidCols; }
__device__ __forceinline__ float *bnv_beta(const MidCtx &c, int s) { return c.bnv + (s * 3 + 2) * kMidCols; }

// Normalisation vectors of every layer -> shared memory.  mode 0: beta only (training forward: the
// batch statistics follow from the folds), 1: + moving statistics (evaluation), 2: + saved batch
// statistics (backward).  Layers without batch norm get (0, 1, 0): the identity.
__device__ __noinline__ void mid_load_bnv(const MidCtx &c, const MidDesc &d, int mode) {
    MID_SHARED(&c); MID_SHARED(&d);
    const int nl = d.n_enc + d.n_dec;
    for (int s = 0; s < nl; ++s) {
        const MidLayer &l = s < d.n_enc ? d.enc[s] : d.dec[s - d.n_enc];
        for (int col = threadIdx.x; col < kMidCols; col += kMidThreads) {
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
}

// H[c][r] = relu((Y[c][r] - mean) rstd + beta) for valid cells, 0 otherwise; X (nullable) = xhat.
// (lanes along the cells: conflict-free shared-memory access)
__device__ __noinline__ void mid_normalise(const MidCtx &c, int slot, int N, int span, const float *Y, float *X, float *H) {
    MID_SHARED(&c); MID_SHARED(Y); MID_SHARED(X); MID_SHARED(H);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float *m = bnv_mean(c, slot), *rs = bnv_rstd(c, slot), *bt = bnv_beta(c, slot);
#pragma unroll 4
    for (int col = warp; col < N; col += kMidThreads / 32) {
        const float mean = m[col], rstd = rs[col], beta = bt[col];
        for (int r = lane; r < span; r += 32) {
            const int i = col * c.RP + r;
            float xh = 0.f, h = 0.f;
            if (r < c.nr) {
                xh = (Y[i] - mean) * rstd;
                h = fmaxf(xh + beta, 0.f);
            }
            if (X) X[i] = xh;
            if (H) H[i] = h;
        }
    }
}

// =============================================================================================
// forward
// =============================================================================================
// One layer's normalisation + ReLU on Y (shared) -> H (shared), batch statistics through the grid
// barrier (H doubles as the scratch of the fold: it is written only afterwards).
__device__ __noinline__ void mid_fwd_bn_relu(const MidCtx &c, const MidDesc &d, const MidLayer &l, int slot, const float *Y,
                                             float *H, bool training, int span) {
    MID_SHARED(&c); MID_SHARED(&d); MID_SHARED(Y);
    const int N = l.n_out;
    if (l.beta && training) {
        mid_bn_partial(c, Y, N, ws_stat_slot(d, slot));
        mid_stamp(d, slot == 0 ? 2 : 10);
        mid_grid_sync(d.barrier, d.error);
        mid_stamp(d, slot == 0 ? 3 : 11);
        mid_bn_fold(c, d, l, ws_stat_slot(d, slot), H, bnv_mean(c, slot), bnv_rstd(c, slot));
        mid_stamp(d, slot == 0 ? 4 : 12);
    }
    mid_normalise(c, slot, N, span, Y, nullptr, H);
    set_aug_cols(c, H, N, (N + 1 + 3) & ~3, span);
}

template <int TM>
__global__ void __maxnreg__(kMidMaxRegs) vae_mid_fwd_kernel(const __grid_constant__ MidDesc param) {
    constexpr int SPAN = MidTraits<TM>::kSpan;
    pdl_wait();          // (programmatic dependent launch: common.cuh)
    pdl_trigger();
    extern __shared__ __align__(16) float mid_smem[];
    __shared__ MidShared sh;
    mid_setup(sh, param, mid_smem);
    const MidDesc &d = sh.d;
    const MidCtx &c = sh.c;
    const bool training = d.training != 0;
    const int L = d.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *E = c.act[3];       // noise slab [l][r]
    MID_SHARED(E);
    mid_stamp(d, 0);

    // ---- prologue: every load that depends on no other CTA, in ONE round trip ------------------------
    // first-layer pre-activations (the tensor-core product's split-K partials, folded in fixed order:
    // this is that GEMM's epilogue), the noise, the normalisation vectors of every layer, and the
    // weights of the first product of this kernel.
    const MidLayer &l0 = d.enc[0];
    const int N0 = l0.n_out;
    // posterior heads: one product for [mu | log_sigma] when both fit the 128-column tile
    const int post_parts = (2 * L <= kMidCols) ? 1 : 2;
    const int NP = post_parts == 1 ? 2 * L : L;
    const MidLayer &first = d.n_enc > 1 ? d.enc[1] : d.post;
    const int firstN = d.n_enc > 1 ? first.n_out : NP;
    stage_w_async(c, first.w, first.ldw, 0, firstN, first.k_in);
    const bool use_eps = !d.deterministic;
    if (use_eps && !d.generate_eps) slab_copy_async(c, E, d.eps, L, L, SPAN);
    if (d.y1_nsplit == 1 && d.y1_alpha == 1.f) slab_copy_async(c, c.act[1], d.y1_parts, d.y1_ld, N0, SPAN);
    else slab_sum(c, c.act[1], d.y1_parts, d.y1_ld, N0, SPAN, d.y1_nsplit, d.y1_slice, d.y1_alpha);
    mid_load_bnv(c, d, training ? 0 : 1);
    if (use_eps && d.generate_eps) {
        uint64_t offset = d.offset;
        if (d.offset_dev) offset += (uint64_t)(*d.offset_dev);
        const int64_t e0 = (int64_t)c.r0 * L;
        const int n_el = c.nr * L;                    // elements of this CTA's cells
        const int lead = (int)(e0 & 3);               // offset of the first one inside its Philox quad
        for (int q = threadIdx.x; q * 4 < n_el + lead; q += kMidThreads) {
            curandStatePhilox4_32_10_t st;            // same stream as scvae_fill_normal
            curand_init(d.seed, (unsigned long long)((e0 >> 2) + q), 4ull * offset, &st);
            const float4 v = curand_normal4(&st);
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int e = q * 4 + j - lead;       // element index relative to the CTA's first
                if (e >= 0 && e < n_el) {
                    const int r = e / L;
                    d.eps[e0 + e] = vv[j];
                    E[(e - r * L) * c.RP + r] = vv[j];
                }
            }
        }
    }
    cp_async_wait();
    __syncthreads();
    if (l0.y && d.y1_parts != l0.y) slab_store(c, c.act[1], l0.y, l0.ldy, N0);
    mid_stamp(d, 1);

    int cur = 0;          // buffer holding the current activation H
    mid_fwd_bn_relu(c, d, l0, 0, c.act[1], c.act[0], training, SPAN);
    __syncthreads();
    mid_stamp(d, 5);
    for (int i = 1; i < d.n_enc; ++i) {
        const MidLayer &l = d.enc[i];
        const int yb = (cur + 1) % 3, hb = (cur + 2) % 3;
        if (i > 1) {
            stage_w_async(c, l.w, l.ldw, 0, l.n_out, l.k_in);
            cp_async_wait();
            __syncthreads();
        }
        mid_product<TM>(c, c.act[cur], l.k_in, c.act[yb], l.n_out, l.y, l.ldy);
        __syncthreads();
        mid_fwd_bn_relu(c, d, l, i, c.act[yb], c.act[hb], training, SPAN);
        cur = hb;
        __syncthreads();
    }
    // ---- posterior heads: [mu | log_sigma] = h W^T (no batch norm, VAE:2268-2289) ----------------
    const int mb = (cur + 1) % 3, lb = (cur + 2) % 3;
    {
        const MidLayer &l = d.post;
        for (int part = 0; part < post_parts; ++part) {
            if (part > 0 || d.n_enc > 1) {
                stage_w_async(c, l.w, l.ldw, part * L, NP, l.k_in);
                cp_async_wait();
                __syncthreads();
            }
            mid_product<TM>(c, c.act[cur], l.k_in, c.act[part ? lb : mb], NP, nullptr, 0);
            __syncthreads();
        }
    }
    mid_stamp(d, 6);
    // the decoder's first weights travel while the sample is formed
    const MidLayer &dl0 = d.dec[0];
    stage_w_async(c, dl0.w, dl0.ldw, 0, dl0.n_out, dl0.k_in);
    // ---- sample and KL (VAE:2353-2369, :2624-2627) ----------------------------------------------------
    {
        float *sMu = c.act[mb], *sLs = post_parts == 1 ? c.act[mb] + L * c.RP : c.act[lb], *sZ = c.act[cur];
        MID_SHARED(sMu); MID_SHARED(sLs); MID_SHARED(sZ);
        const int Kzp = (dl0.k_in + 3) & ~3;
        // lanes along the latent dimension: coalesced HBM copies of [mu | raw log_sigma] and z
        for (int r = warp; r < SPAN; r += kMidThreads / 32) {
            const bool valid = r < c.nr;
            const int64_t row = c.r0 + r;
            float kl = 0.f;
            for (int l = lane; l < L; l += 32) {
                const int i = l * c.RP + r;
                float zv = 0.f;
                if (valid) {
                    const float mu = sMu[i];
                    const float raw = sLs[i];
                    const float ls = fminf(fmaxf(raw, -3.f), 3.f);
                    const float sigma = __expf(ls);
                    // tfp kl_divergence(Normal(mu, sigma), Normal(0, 1))
                    const float k = 0.5f * mu * mu + 0.5f * (sigma * sigma - 1.f) - ls;
                    kl += k;
                    zv = d.deterministic ? mu : mu + sigma * E[i];
                    if (d.kl_elem) d.kl_elem[row * L + l] = k;
                    d.ph[row * d.ldph + l] = mu;
                    d.ph[row * d.ldph + L + l] = raw;
                    d.z[row * d.ldz + l] = zv;
                }
                sZ[i] = zv;
            }
            // augmented column, decoder-input extras (VAE:2400-2441), zero padding
            for (int col = L + lane; col < Kzp; col += 32) {
                float v = 0.f;
                if (valid) {
                    if (col == L) v = 1.f;
                    else if (d.batch_index && col - (L + 1) < d.n_batches)
                        v = ((int)d.batch_index[row] == col - (L + 1)) ? 1.f : 0.f;
                    else if (d.count_sum && col == L + 1 + (d.batch_index ? d.n_batches : 0))
                        v = d.count_sum[row];
                    if (col < d.ldz) d.z[row * d.ldz + col] = v;
                }
                sZ[col * c.RP + r] = v;
            }
            kl = warp_sum(kl);
            if (valid && lane == 0 && d.kl_row) d.kl_row[row] = kl;
        }
    }
    cp_async_wait();
    __syncthreads();
    mid_stamp(d, 7);
    // ---- decoder layers ---------------------------------------------------------------------------
    for (int j = 0; j < d.n_dec; ++j) {
        const MidLayer &l = d.dec[j];
        const int yb = (cur + 1) % 3, hb = (cur + 2) % 3;
        if (j > 0) {
            stage_w_async(c, l.w, l.ldw, 0, l.n_out, l.k_in);
            cp_async_wait();
            __syncthreads();
        }
        mid_product<TM>(c, c.act[cur], l.k_in, c.act[yb], l.n_out, l.y, l.ldy);
        __syncthreads();
        mid_fwd_bn_relu(c, d, l, d.n_enc + j, c.act[yb], c.act[hb], training, SPAN);
        cur = hb;
        __syncthreads();
    }
    mid_stamp(d, 13);
    // ---- operand of the fused likelihood heads: fp16, augmented, zero padded to ldd16 columns ----
    {
        const int N = d.dec[d.n_dec - 1].n_out;
        const float *H = c.act[cur];
        MID_SHARED(H);
        const int groups = (int)(d.ldd16 >> 3);
        __half *out = reinterpret_cast<__half *>(d.d16);
        for (int r = warp; r < c.nr; r += kMidThreads / 32) {
            for (int g = lane; g < groups; g += 32) {
                const int c0 = g << 3;
                __align__(16) __half h[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) h[j] = __float2half_rn((c0 + j <= N) ? H[(c0 + j) * c.RP + r] : 0.f);   // column N = 1
                *reinterpret_cast<uint4 *>(out + (int64_t)(c.r0 + r) * d.ldd16 + c0) = *reinterpret_cast<const uint4 *>(h);
            }
        }
        if (d.h_last) slab_store(c, H, d.h_last, d.ldh_last, N + 1);
    }
    mid_stamp(d, 14);
}

// =============================================================================================
// backward
// =============================================================================================
// In place on Gd[c][r] (d loss / d activation): ReLU mask, batch-norm backward.  X holds xhat.
// dy = rstd (g - s1/B - xhat s2/B);  dbeta = s1  (s1 = sum g, s2 = sum g xhat over ALL cells).
// `scratch`: [8 warps][2][128] floats of idle shared memory.
__device__ __noinline__ void mid_bn_relu_bwd(const MidCtx &c, const MidDesc &d, const MidLayer &l, int s, int span, float *Gd,
                                             const float *X, float *ws_stat, float *scratch) {
    MID_SHARED(&c); MID_SHARED(&d); MID_SHARED(Gd); MID_SHARED(X); MID_SHARED(scratch);
    const int N = l.n_out;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float *rs = bnv_rstd(c, s), *bt = bnv_beta(c, s);
    // g = dH * [H > 0], H = relu(xhat + beta)
#pragma unroll 4
    for (int col = warp; col < N; col += kMidThreads / 32) {
        const float beta = bt[col];
        for (int r = lane; r < span; r += 32) {
            const int i = col * c.RP + r;
            Gd[i] = (r < c.nr && X[i] + beta > 0.f) ? Gd[i] : 0.f;
        }
    }
    __syncthreads();
    if (!l.beta) return;
    {
        // per-CTA column sums: one thread per (column, half of the cells)
        const int col = threadIdx.x & (kMidCols - 1), half = threadIdx.x >> 7;
        const int h0 = half ? (c.nr >> 1) : 0, h1 = half ? c.nr : (c.nr >> 1);
        float s1 = 0.f, s2 = 0.f;
        if (col < N) {
#pragma unroll 4
            for (int r = h0; r < h1; ++r) {
                const float g = Gd[col * c.RP + r];
                s1 += g;
                s2 = fmaf(g, X[col * c.RP + r], s2);
            }
        }
        float *x = c.stat + 4 * kMidCols;        // [2 halves][2][128]
        x[(half * 2 + 0) * kMidCols + col] = s1;
        x[(half * 2 + 1) * kMidCols + col] = s2;
        __syncthreads();
        if (half == 0 && col < N) {
            ws_stat[((int64_t)blockIdx.x * 2 + 0) * kMidCols + col] = x[col] + x[2 * kMidCols + col];
            ws_stat[((int64_t)blockIdx.x * 2 + 1) * kMidCols + col] = x[kMidCols + col] + x[3 * kMidCols + col];
        }
    }
    mid_grid_sync(d.barrier, d.error);
    float *t1 = c.stat, *t2 = c.stat + kMidCols;
    if (gridDim.x <= kMidThreads / 32) {
        // small grids: thread t < N adds the <= 8 partials of column t in CTA order
        const int col = threadIdx.x, G = gridDim.x;
        if (col < N) {
            float v1[kMidThreads / 32], v2[kMidThreads / 32];
#pragma unroll
            for (int k = 0; k < kMidThreads / 32; ++k) {
                v1[k] = k < G ? __ldcg(ws_stat + ((int64_t)k * 2 + 0) * kMidCols + col) : 0.f;
                v2[k] = k < G ? __ldcg(ws_stat + ((int64_t)k * 2 + 1) * kMidCols + col) : 0.f;
            }
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int k = 0; k < kMidThreads / 32; ++k) {
                s1 += v1[k];
                s2 += v2[k];
            }
            t1[col] = s1;
            t2[col] = s2;
            if (blockIdx.x == 0) l.dbeta[col] = s1;
        }
    } else {
        // fixed-order fold of the CTA partials: warp w takes the CTAs k = w mod 8
        const int G = gridDim.x;
        float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
        float v1[kMidFoldIt][4], v2[kMidFoldIt][4];
#pragma unroll
        for (int it = 0; it < kMidFoldIt; ++it) {          // every load in flight before the first add
            const int k = warp + it * (kMidThreads / 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int col = lane + 32 * q;
                const bool ok = k < G && col < N;
                v1[it][q] = ok ? __ldcg(ws_stat + ((int64_t)k * 2 + 0) * kMidCols + col) : 0.f;
                v2[it][q] = ok ? __ldcg(ws_stat + ((int64_t)k * 2 + 1) * kMidCols + col) : 0.f;
            }
        }
#pragma unroll
        for (int it = 0; it < kMidFoldIt; ++it)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                a1[q] += v1[it][q];
                a2[q] += v2[it][q];
            }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            scratch[(warp * 2 + 0) * kMidCols + lane + 32 * q] = a1[q];
            scratch[(warp * 2 + 1) * kMidCols + lane + 32 * q] = a2[q];
        }
        __syncthreads();
        const int col = threadIdx.x;
        if (col < N) {
            float s1 = 0.f, s2 = 0.f;
            for (int w = 0; w < kMidThreads / 32; ++w) {
                s1 += scratch[(w * 2 + 0) * kMidCols + col];
                s2 += scratch[(w * 2 + 1) * kMidCols + col];
            }
            t1[col] = s1;
            t2[col] = s2;
            if (blockIdx.x == 0) l.dbeta[col] = s1;
        }
    }
    __syncthreads();
    const float inv_b = 1.f / (float)d.B;
#pragma unroll 4
    for (int col = warp; col < N; col += kMidThreads / 32) {
        const float rstd = rs[col], m1 = t1[col] * inv_b, m2 = t2[col] * inv_b;
        for (int r = lane; r < span; r += 32) {
            const int i = col * c.RP + r;
            Gd[i] = r < c.nr ? rstd * (Gd[i] - m1 - X[i] * m2) : 0.f;
        }
    }
    __syncthreads();
}

template <int TM>
__global__ void __maxnreg__(kMidMaxRegs) vae_mid_bwd_kernel(const __grid_constant__ MidDesc param) {
    constexpr int SPAN = MidTraits<TM>::kSpan;
    pdl_wait();          // (programmatic dependent launch: common.cuh)
    pdl_trigger();
    extern __shared__ __align__(16) float mid_smem[];
    __shared__ MidShared sh;
    mid_setup(sh, param, mid_smem);
    const MidDesc &d = sh.d;
    const MidCtx &c = sh.c;
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
    MID_SHARED(Gd); MID_SHARED(X); MID_SHARED(T); MID_SHARED(F);
    mid_stamp(d, 0);
    {
        const int N = last.n_out;
        if (d.dd_nsplit == 1) slab_copy_async(c, Gd, d.dd_parts, d.dd_ld, N, SPAN);
        else slab_sum(c, Gd, d.dd_parts, d.dd_ld, N, SPAN, d.dd_nsplit, d.dd_slice, 1.f);
        slab_copy_async(c, X, last.y, last.ldy, N, SPAN);
        // the fp16 operand the heads saw, as 32-bit words (pairs of halves): T[w][r]
        slab_copy_async(c, T, reinterpret_cast<const float *>(d.d16), d.ldd16 >> 1, (N + 1) >> 1, SPAN);
        stage_w_async(c, last.w, last.ldw, 0, last.n_out, last.n_in);      // dgrad weights of the last decoder layer
        mid_load_bnv(c, d, 2);
        // per-cell scalars: threads 0..nr-1
        float lp = 0.f, klr = 0.f;
        if ((int)threadIdx.x < c.nr) {
            const int64_t row = c.r0 + threadIdx.x;
            for (int s = 0; s < d.logp_nsplit; ++s) lp += __ldcg(d.logp_parts + (int64_t)s * d.logp_slice + row);
            if (d.row_const) lp -= d.row_const[row];
            klr = d.kl_row[row];
        }
        cp_async_wait();
        __syncthreads();
        // X: y -> xhat; per-cell correction sum_h (h - h16) dd_h, lanes along the columns (coalesced d16)
        const float *m = bnv_mean(c, s_last), *rs = bnv_rstd(c, s_last), *bt = bnv_beta(c, s_last);
        const float inv_go = 1.f / d.go_scalar;
        for (int r = warp; r < SPAN; r += kMidThreads / 32) {
            const bool valid = r < c.nr;
            float corr = 0.f;
#pragma unroll 4
            for (int col = lane; col < N; col += 32) {
                const int i = col * c.RP + r;
                float xh = 0.f;
                if (valid) {
                    xh = (X[i] - m[col]) * rs[col];
                    const float h = fmaxf(xh + bt[col], 0.f);      // fp32 activation the heads saw in fp16
                    const float word = T[(col >> 1) * c.RP + r];
                    const __half2 h2 = *reinterpret_cast<const __half2 *>(&word);
                    corr = fmaf(h - ((col & 1) ? __high2float(h2) : __low2float(h2)), Gd[i], corr);
                }
                X[i] = xh;
            }
            corr = warp_sum(corr) * inv_go;
            if (lane == 0 && valid) c.stat[2 * kMidCols + r] = corr;
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
    }
    mid_stamp(d, 1);
    // ---- decoder layers, last to first -------------------------------------------------------------
    // invariant at the top: Gd = d loss / d activation of layer j, X = its xhat, sw = its dgrad weights
    // (in flight or landed); T, F free
    for (int j = d.n_dec - 1; j >= 0; --j) {
        const MidLayer &l = d.dec[j];
        const int s = d.n_enc + j;
        // issued ahead of the barrier inside the batch-norm backward: the layer's input -- the
        // previous decoder activation (its stored pre-activations) or the latent sample
        float *In = T, *Xprev = F;
        const int Kp = (int)l.ldw;
        if (j > 0) {
            slab_copy_async(c, Xprev, d.dec[j - 1].y, d.dec[j - 1].ldy, d.dec[j - 1].n_out, SPAN);
        } else {
            const int zc = min(Kp, (int)d.ldz);
            slab_copy_async(c, In, d.z, d.ldz, zc, SPAN);
            if (zc < Kp) set_aug_cols(c, In, Kp, Kp, SPAN);      // (never: ldz == Kp)
        }
        mid_stamp(d, 2);
        mid_bn_relu_bwd(c, d, l, s, SPAN, Gd, X, ws_stat_slot(d, slot++), j > 0 ? T : F);
        mid_stamp(d, 3);
        cp_async_wait();
        __syncthreads();
        if (j > 0) {
            mid_normalise(c, s - 1, d.dec[j - 1].n_out, SPAN, Xprev, Xprev, In);
            set_aug_cols(c, In, d.dec[j - 1].n_out, Kp, SPAN);
            __syncthreads();
        }
        mid_wgrad(c, Gd, l.n_out, In, Kp, dw_part(dw_ws, l));
        mid_stamp(d, 4);
        float *Out = X;      // xhat of this layer is dead: the dgrad output takes its buffer
        mid_dgrad_tile<TM>(c, Gd, l.n_out, Out, l.n_in, true);
        __syncthreads();
        mid_stamp(d, 5);
        // the next product's dgrad weights travel during the next phase
        if (j > 0) stage_w_async(c, d.dec[j - 1].w, d.dec[j - 1].ldw, 0, d.dec[j - 1].n_out, d.dec[j - 1].n_in);
        // rotate: Gd <- Out, X <- Xprev, free: old Gd (-> F), In (T stays T)
        float *oldG = Gd;
        Gd = Out;
        X = Xprev;
        F = oldG;
    }
    // ---- sample / KL backward (VAE:2353-2369 and the analytic KL): Gd = dZ[l][r] -----------------
    // dmu = dz + c mu;  dlog_sigma = (dz eps sigma + c (sigma^2 - 1)) [|raw| <= 3];  c = weight / B
    const int post_parts = (2 * L <= kMidCols) ? 1 : 2;
    const int NP = post_parts == 1 ? 2 * L : L;
    float *Gmu = Gd, *Gls = post_parts == 1 ? Gd + L * c.RP : X;     // (one buffer when both halves fit)
    MID_SHARED(Gd); MID_SHARED(Gmu); MID_SHARED(Gls);
    const MidLayer &pl = d.post;
    const MidLayer &prev = d.enc[d.n_enc - 1];
    stage_w_async(c, pl.w, pl.ldw, 0, NP, pl.n_in);
    slab_copy_async(c, F, prev.y, prev.ldy, prev.n_out, SPAN);       // pre-activations of the last encoder layer
#pragma unroll 4
    for (int r = warp; r < SPAN; r += kMidThreads / 32) {        // (unrolled: several cells' loads in flight)
        const bool valid = r < c.nr;
        const int64_t row = c.r0 + r;
        for (int l = lane; l < L; l += 32) {
            const int i = l * c.RP + r;
            float gm = 0.f, gl = 0.f;
            if (valid) {
                const float mu = d.ph[row * d.ldph + l];
                const float raw = d.ph[row * d.ldph + L + l];
                const float e = d.deterministic ? 0.f : d.eps[row * L + l];
                const float ls = fminf(fmaxf(raw, -3.f), 3.f);
                const float sigma = __expf(ls);
                const float dz = Gd[i];
                gm = dz + kl_coef * mu;
                const float mask = (raw < -3.f || raw > 3.f) ? 0.f : 1.f;
                gl = (dz * e * sigma + kl_coef * (sigma * sigma - 1.f)) * mask;
            }
            Gmu[i] = gm;      // (the element this thread just read)
            Gls[i] = gl;
        }
    }
    mid_stamp(d, 6);
    // ---- posterior heads: wgrad of both halves, dgrad summed over them ----------------------------
    {
        float *In = T;
        const int Kp = (int)pl.ldw;
        cp_async_wait();
        __syncthreads();
        mid_normalise(c, d.n_enc - 1, prev.n_out, SPAN, F, F, In);      // F = xhat of the last encoder layer (kept)
        set_aug_cols(c, In, prev.n_out, Kp, SPAN);
        __syncthreads();
        mid_stamp(d, 7);
        float *mine = dw_part(dw_ws, pl);
        mid_wgrad(c, Gmu, NP, In, Kp, mine);
        if (post_parts == 2) mid_wgrad(c, Gls, L, In, Kp, mine + (int64_t)L * Kp);
        __syncthreads();
        mid_stamp(d, 8);
        mid_dgrad_tile<TM>(c, Gmu, NP, In, pl.n_in, true);          // d loss / d H of the last encoder layer
        __syncthreads();
        if (post_parts == 2) {
            stage_w_async(c, pl.w, pl.ldw, L, L, pl.n_in);
            cp_async_wait();
            __syncthreads();
            mid_dgrad_tile<TM>(c, Gls, L, In, pl.n_in, false);
            __syncthreads();
        }
        mid_stamp(d, 9);
        // Gd <- In (T's buffer); X <- F (xhat); free: Gmu's buffer and X's old one
        float *f0 = Gd, *f1 = X;
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
            stage_w_async(c, l.w, l.ldw, 0, l.n_out, l.n_in);
            slab_copy_async(c, Xprev, d.enc[i - 1].y, d.enc[i - 1].ldy, d.enc[i - 1].n_out, SPAN);
        }
        mid_bn_relu_bwd(c, d, l, i, SPAN, Gd, X, ws_stat_slot(d, slot++), i > 0 ? In : T);
        if (i == 0) break;
        cp_async_wait();
        __syncthreads();
        mid_normalise(c, i - 1, d.enc[i - 1].n_out, SPAN, Xprev, Xprev, In);
        set_aug_cols(c, In, d.enc[i - 1].n_out, Kp, SPAN);
        __syncthreads();
        mid_wgrad(c, Gd, l.n_out, In, Kp, dw_part(dw_ws, l));
        float *Out = X;
        mid_dgrad_tile<TM>(c, Gd, l.n_out, Out, l.n_in, true);
        __syncthreads();
        float *oldG = Gd;
        Gd = Out;
        X = Xprev;
        F = oldG;
    }
    mid_stamp(d, 10);
    // dY1 (B, H1) -> fp16 + rounding remainder, scaled into range, zero padded: operand of dW1 = dY1^T X
    {
        const int N = d.enc[0].n_out;
        const int groups = (int)(d.lddy1 >> 3);
        __half *out = reinterpret_cast<__half *>(d.dy1_16);
        __half *out_lo = reinterpret_cast<__half *>(d.dy1_16_lo);
        for (int r = warp; r < c.nr; r += kMidThreads / 32) {
            for (int g = lane; g < groups; g += 32) {
                const int c0 = g << 3;
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
        }
        if (d.dy1) slab_store(c, Gd, d.dy1, d.lddy1_f32, N);
        if (d.enc[0].dw) {
            // bias gradient of the first layer = column sums of dY1, in fp32 (behind a batch norm it is
            // zero up to rounding, and the fp16 copy above would turn it into rounding noise that Adam
            // normalises to +-lr): per-CTA sums here, folded in CTA order after the barrier
            const int col = threadIdx.x & (kMidCols - 1), half = threadIdx.x >> 7;
            const int h0 = half ? (c.nr >> 1) : 0, h1 = half ? c.nr : (c.nr >> 1);
            float s1 = 0.f;
            if (col < N) {
#pragma unroll 4
                for (int r = h0; r < h1; ++r) s1 += Gd[col * c.RP + r];
            }
            float *x = c.stat + 4 * kMidCols;
            __syncthreads();
            x[half * kMidCols + col] = s1;
            __syncthreads();
            if (half == 0) ws_db1(d)[(int64_t)blockIdx.x * kMidCols + col] = x[col] + x[kMidCols + col];
        }
    }
    // ---- all partials are in the workspace: fold them in fixed order --------------------------------
    mid_stamp(d, 11);
    mid_grid_sync(d.barrier, d.error);
    mid_stamp(d, 12);
    {
        float *cursor = ws_dw_base(d);
        const int G = gridDim.x;
        auto reduce = [&](const MidLayer &l) {
            const int64_t n = (int64_t)l.n_out * l.ldw;
            const int64_t stride = (int64_t)gridDim.x * kMidThreads;
            // four elements per thread and trip: with a small grid a thread owns many elements, and
            // one element's few partials alone would leave a single load in flight
            for (int64_t e = (int64_t)blockIdx.x * kMidThreads + threadIdx.x; e < n; e += 4 * stride) {
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                const bool in1 = e + stride < n, in2 = e + 2 * stride < n, in3 = e + 3 * stride < n;
                if (!in1) {          // large grids: one element per thread, eight partials in flight
#pragma unroll 8
                    for (int k = 0; k < G; ++k) s0 += __ldcg(cursor + (int64_t)k * n + e);
                    l.dw[e] = s0;
                    continue;
                }
#pragma unroll 4
                for (int k = 0; k < G; ++k) {                                              // CTA order
                    const float *src = cursor + (int64_t)k * n + e;
                    s0 += __ldcg(src);
                    if (in1) s1 += __ldcg(src + stride);
                    if (in2) s2 += __ldcg(src + 2 * stride);
                    if (in3) s3 += __ldcg(src + 3 * stride);
                }
                l.dw[e] = s0;
                if (in1) l.dw[e + stride] = s1;
                if (in2) l.dw[e + 2 * stride] = s2;
                if (in3) l.dw[e + 3 * stride] = s3;
            }
            cursor += (int64_t)gridDim.x * n;
        };
        for (int j = d.n_dec - 1; j >= 0; --j) reduce(d.dec[j]);
        reduce(d.post);
        for (int i = d.n_enc - 1; i >= 1; --i) reduce(d.enc[i]);
        if (d.enc[0].dw && blockIdx.x == (gridDim.x > 1 ? gridDim.x - 2 : 0)) {
            // first-layer bias gradient: thread (column, half) adds its half of the CTA partials in
            // CTA order, the two halves are added in order (a CTA near the end of the grid: it
            // owns few or none of the elements folded above)
            const int col = threadIdx.x & (kMidCols - 1), half = threadIdx.x >> 7;
            const int k0 = half ? (G >> 1) : 0, k1 = half ? G : (G >> 1);
            const float *src = ws_db1(d) + col;
            float s0 = 0.f;
#pragma unroll 8
            for (int k = k0; k < k1; ++k) s0 += __ldcg(src + (int64_t)k * kMidCols);
            float *x = c.stat + 4 * kMidCols;
            __syncthreads();
            x[half * kMidCols + col] = s0;
            __syncthreads();
            const MidLayer &l0 = d.enc[0];
            if (half == 0 && col < l0.n_out) l0.dw[(int64_t)col * l0.ldw + l0.n_in] = x[col] + x[kMidCols + col];
            __syncthreads();
        }
        if (blockIdx.x == gridDim.x - 1) {       // (the last CTA has the fewest cells: least other work)
            for (int k = threadIdx.x; k < G; k += kMidThreads) {
                c.stat[k] = __ldcg(ws_bound(d) + k * 4 + 0);
                c.stat[kMidThreads + k] = __ldcg(ws_bound(d) + k * 4 + 1);
            }
            __syncthreads();
            // fixed order: lane l sums the CTAs l, l + 32, ...; lane 0 then adds the lanes in order
            float lp = 0.f, kl = 0.f;
            if (warp == 0) {
                for (int k = lane; k < G; k += 32) {
                    lp += c.stat[k];
                    kl += c.stat[kMidThreads + k];
                }
                float lps = 0.f, kls = 0.f;
                for (int src = 0; src < 32; ++src) {
                    lps += __shfl_sync(0xffffffffu, lp, src);
                    kls += __shfl_sync(0xffffffffu, kl, src);
                }
                lp = lps;
                kl = kls;
            }
            if (threadIdx.x == 0) {
                // lower bound (R = 1): mean_b(log p - KL); weighted; ENRE; KL (VAE:2715-2734)
                const float inv_b = 1.f / (float)d.B;
                d.bound[0] = (lp - kl) * inv_b;
                d.bound[1] = (lp - weight * kl) * inv_b;
                d.bound[2] = lp * inv_b;
                d.bound[3] = kl * inv_b;
            }
        }
    }
    mid_stamp(d, 13);
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
    n += grid * kMidCols;          // per-CTA column sums of dY1 (bias gradient of the first layer)
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
    launch_pdl(BWD ? kPdlMidBwd : kPdlMidFwd, kern, dim3(grid), dim3(kMidThreads), smem, s, *d);
    SCVAE_CHECK_LAUNCH(name);
    return 0;
}

template <bool BWD>
static int mid_launch(const char *name, const MidDesc *d, cudaStream_t s) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = mid_grid(d);
    SCVAE_CHECK_ARG(grid <= sms && grid <= kMidFoldIt * (kMidThreads / 32),
                    "%s: %d cells need %d CTAs of %d cells, the device has %d SMs (grid barrier)", name, d->B, grid,
                    d->rows_per_cta, sms);
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
