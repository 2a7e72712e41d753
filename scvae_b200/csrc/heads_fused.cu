// Fused likelihood heads: GEMM -> count likelihood -> (log p, d log p/d a) -> dgrad GEMM in ONE
// kernel, so the (cells x P*genes) head pre-activations never exist in HBM.
//
// Replaces, for one training step, the chain
//     X_TILDE/<param> fully_connected  (VAE:2466-2489)   a   = d W^T          (cells x P*genes)
//     p_x_given_z.log_prob + reduce_sum (VAE:2583-2590)  logp, da             (cells x P*genes)
//     autodiff of the heads w.r.t. the decoder output     dd  = da W           (cells x H)
// and leaves only `da` (in fp16) for the weight-gradient product dW = da^T d, which is a plain
// tcgen05 GEMM (scvae_gemm_f16).  Per step at C2 this removes ~2.6 GB of HBM traffic.
//
// Structure (one CTA = 128 cells x a range of 64-gene tiles, 12 warps, 1 CTA/SM):
//   warp 0     TMA producer: decoder-output tile d (128 x 128 fp16, once), then per gene tile the
//              head-weight tiles W_p (64 x 128 fp16) and the target tile t (128 x 64 u16);
//   warp 1     one thread issues tcgen05.mma kind::f16:
//                MMA1  S_p[128 x 64]  = d . W_p^T            (TMEM, double buffered)
//                MMA2  dd[128 x 128] += da_p . W_p            (TMEM, accumulated over all tiles;
//                      W_p is the SAME smem tile read as an MN-major operand)
//   warps 4-11 epilogue: tcgen05.ld S -> likelihood math (likelihood_math.cuh) -> row-sum of
//              log p in registers, da_p -> fp16 -> swizzled smem (operand of MMA2 AND source of
//              the TMA store of da to HBM);
//   MMA1 of tile n+1 overlaps the epilogue of tile n (issue order: MMA1(n+1), MMA2(n)).
// Gene-range partial sums (log p per cell, dd) are combined deterministically / by TMA
// reduce-add.  Gradients are scaled by `scale` before the fp16 conversion and un-scaled in the
// consumers (loss-scaling against fp16 underflow).
#include "fused_math.cuh"
#include "tc_common.cuh"

namespace scvae {

constexpr int FM = 128;        // cells per CTA
constexpr int FG = 64;         // genes per tile
constexpr int FK = 128;        // padded hidden width (fp16 elements)
// 4 control warps + EW epilogue warps.  A tile has 16 (32-row x 16-gene) chunks and all epilogue
// warps meet at the da tile once per tile, so EW must divide 16: 16 warps x 1 chunk (4 per
// scheduler, 96 registers, 1 KB of fp16 total_count scratch each) for every likelihood
__host__ __device__ constexpr int fused_epi_warps(int P) { return 16; }
__host__ __device__ constexpr int fused_threads(int P) { return 128 + 32 * fused_epi_warps(P); }
// head-weight ring: a stage is held from its TMA load until MMA2 of its tile has finished, i.e.
// for two tiles; a third stage takes the reload off the critical path (smem allows it for P <= 2)
__host__ __device__ constexpr int fused_w_stages(int P) { return P <= 2 ? 3 : 2; }
// da tile buffers: with two, an epilogue warp may run one tile ahead of MMA2 / the TMA store, so
// the warps drift apart and MUFU-bound math overlaps the latency-bound x >= 2 fix-ups of others
__host__ __device__ constexpr int fused_a_bufs(int P) { return 1; }
constexpr int FDBytes = FM * FK * 2;         // 32 KB
constexpr int FWBytes = FG * FK * 2;         // 16 KB per head
constexpr int FABytes = FM * FG * 2;         // 16 KB per head (da tile)
constexpr int FTBytes = FM * FG * 2;         // 16 KB (u16 targets)
constexpr int FRBytes = 32 * 16 * 2;         // per epilogue warp: fp16 total_count of a 16-gene chunk

__host__ __device__ constexpr int fused_smem_bytes(int P) {
    return FDBytes + fused_w_stages(P) * P * FWBytes + fused_a_bufs(P) * P * FABytes + 2 * FTBytes +
           fused_epi_warps(P) * FRBytes +
           1024 /*align*/ + 256 /*barriers*/;
}

__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// explicit shared-space accesses (generic LD/ST to shared memory take the long L1TEX path)
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_v4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint16_t lds_u16(uint32_t a) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u16(uint32_t a, uint16_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(v) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void bar_sync_n(int id, int n) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&h);
}

struct FusedParams {
    int M, G, t_rows;
    int tiles_per_cta, n_tiles, gsplit;
    int64_t head_stride;      // columns between heads in W16 rows / da16 columns (multiple of 64)
    int64_t part_stride;      // rows per split slice of logp_part
    const float *go;          // nullable per-row upstream gradient
    float go_scalar, scale, inv_scale;
    int has_const;            // sum_g lgamma(1+t) is subtracted by the finish kernel
    float *logp_part;
    long long *dbg;           // development aid: per-tile clock64 timeline of CTA 0 (NULL in production)
};
#ifdef SCVAE_FUSED_TIMELINE      // tools/fused_bench.py --timeline (build with -DSCVAE_FUSED_TIMELINE)
#define FUSED_DBG(n, ev)                                                        \
    do {                                                                        \
        if (p.dbg && blockIdx.x == 0 && (n) < 40) p.dbg[(n) * 16 + (ev)] = clock64(); \
    } while (0)
#else
#define FUSED_DBG(n, ev) do { } while (0)
#endif
// A 16-gene chunk in which a clip of the reference is active (fused_math.cuh): exact masked math
// of likelihood_math.cuh, out of line (rare; keeps the register budget of the fast path small).
struct FusedSlowIn {
    uint32_t tw[8];
    uint32_t sv[3][16];
};
struct FusedSlowOut {
    uint32_t packed[3][8];
    float acc;
};
template <int KIND, bool T_HALF, bool BWD>
__device__ __noinline__ FusedSlowOut fused_slow_chunk(FusedSlowIn in, float gs_row, int g_first, int G,
                                                      bool has_const) {
    constexpr int P = Lik<KIND>::P;
    FusedSlowOut out;
    out.acc = 0.f;
#pragma unroll 1
    for (int blk = 0; blk < 2; ++blk) {
        float x[8], av[3][8], gv[3][8];
#pragma unroll
        for (int i = 0; i < 4; ++i) fused_cvt2<T_HALF>(in.tw[blk * 4 + i], x[2 * i], x[2 * i + 1]);
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int h = 0; h < P; ++h) av[h][j] = __uint_as_float(in.sv[h][blk * 8 + j]);
        float acc8 = 0.f;
        lik_group<KIND, BWD, 8>(x, av, has_const, acc8, gv);
        const bool valid = (g_first + blk * 8) < G;
        out.acc += valid ? acc8 : 0.f;
        if (!BWD) continue;
        const float gsv = valid ? gs_row : 0.f;
#pragma unroll
        for (int h = 0; h < P; ++h)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                out.packed[h][blk * 4 + j] = pack_half2(gv[h][2 * j] * gsv, gv[h][2 * j + 1] * gsv);
    }
    return out;
}

// T_HALF: targets are fp16 (exact for counts <= 2048) instead of uint16.  BWD = false: forward
// only (evaluation passes): MMA1 + log p; no gradient, no da tile, no MMA2, no dd.
template <int KIND, bool T_HALF, bool BWD>
__global__ void __launch_bounds__(fused_threads(Lik<KIND>::P), 1)
heads_fused_kernel(const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmW,
                   const __grid_constant__ CUtensorMap tmT, const __grid_constant__ CUtensorMap tmDA,
                   const __grid_constant__ CUtensorMap tmDD, const FusedParams p) {
    pdl_wait();          // (programmatic dependent launch: common.cuh)
    pdl_trigger();
    constexpr int P = Lik<KIND>::P;
    constexpr int EW = fused_epi_warps(P);       // epilogue warps
    constexpr int EJ = EW / 4;                   // epilogue warps per TMEM lane quadrant
    static_assert(EJ == 1 || EJ == 2 || EJ == 4, "epilogue warps per quadrant must divide the 4 chunks of a tile");
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sD = smem;
    constexpr int NW = fused_w_stages(P);        // head-weight ring depth
    uint8_t *sW = sD + FDBytes;                 // [NW stages][P][2 k-halves][64 genes][128 B]
    constexpr int NA = fused_a_bufs(P);          // da tile buffers
    uint8_t *sA = sW + NW * P * FWBytes;        // [NA][P][128 rows][128 B]
    uint8_t *sT = sA + NA * P * FABytes;        // [2 stages][128 rows][128 B]
    uint8_t *sR = sT + 2 * FTBytes;             // [EW warps][4 chunks][32 lanes][16 B]
    uint64_t *bars = (uint64_t *)(sR + EW * FRBytes);
    enum { D_FULL = 0, W_FULL = 1, W_EMPTY = W_FULL + NW, T_FULL = W_EMPTY + NW, T_EMPTY = T_FULL + 2,
           S_FULL = T_EMPTY + 2, S_EMPTY = S_FULL + 2, A_FULL = S_EMPTY + 2, A_EMPTY = A_FULL + NA,
           A_STORED = A_EMPTY + NA, DD_FULL = A_STORED + NA, NBARS = DD_FULL + 1 };
    const uint32_t bar0 = smem_u32(bars);
    auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    uint32_t *tmem_slot = (uint32_t *)(bars + NBARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rt = blockIdx.x / p.gsplit, gs = blockIdx.x % p.gsplit;
    const int row0 = rt * FM;
    const int tile0 = gs * p.tiles_per_cta;
    const int tile1 = min(tile0 + p.tiles_per_cta, p.n_tiles);
    const int ntile = tile1 - tile0;             // >= 1 by construction

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmD) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmT) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDD) : "memory");
    }
    if (warp == 1 && lane == 0) {
        mbar_init(bar(D_FULL), 1);
        for (int i = 0; i < NW; ++i) {
            mbar_init(bar(W_FULL + i), 1);
            mbar_init(bar(W_EMPTY + i), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar(T_FULL + i), 1);
            mbar_init(bar(T_EMPTY + i), EW);
            mbar_init(bar(S_FULL + i), 1);
            mbar_init(bar(S_EMPTY + i), EW);
        }
        for (int i = 0; i < NA; ++i) {
            mbar_init(bar(A_FULL + i), EW);
            mbar_init(bar(A_EMPTY + i), 1);
            mbar_init(bar(A_STORED + i), 1);
        }
        mbar_init(bar(DD_FULL), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_S = tmem_base;                     // [2][P][64 columns]
    const uint32_t tmem_DD = tmem_base + 2 * P * FG;       // 128 columns

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            mbar_expect_tx(bar(D_FULL), FDBytes);
            tma_load_2d(smem_u32(sD), &tmD, 0, row0, bar(D_FULL));
            tma_load_2d(smem_u32(sD) + FDBytes / 2, &tmD, 64, row0, bar(D_FULL));
            const int trow0 = row0 % p.t_rows;
            for (int n = 0; n < ntile; ++n) {
                const int st = n & 1;
                const uint32_t ph = (n >> 1) & 1;
                const int g0 = (tile0 + n) * FG;
                const int ws = n % NW;
                mbar_wait(bar(W_EMPTY + ws), (uint32_t)((n / NW) & 1) ^ 1u);
                FUSED_DBG(n, 9);
                mbar_expect_tx(bar(W_FULL + ws), P * FWBytes);
#pragma unroll
                for (int h = 0; h < P; ++h) {
                    const uint32_t dst = smem_u32(sW + (ws * P + h) * FWBytes);
                    const int wrow = (int)(h * p.head_stride) + g0;
                    tma_load_2d(dst, &tmW, 0, wrow, bar(W_FULL + ws));
                    tma_load_2d(dst + FWBytes / 2, &tmW, 64, wrow, bar(W_FULL + ws));
                }
                mbar_wait(bar(T_EMPTY + st), ph ^ 1);
                mbar_expect_tx(bar(T_FULL + st), FTBytes);
                tma_load_2d(smem_u32(sT + st * FTBytes), &tmT, g0, trow0, bar(T_FULL + st));
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // D = f32, A = B = f16; MMA1: 128 x 64, both K-major; MMA2: 128 x 128, B MN-major
            const uint32_t idesc1 = (1u << 4) | ((uint32_t)(FG >> 3) << 17) | ((uint32_t)(FM >> 4) << 24);
            const uint32_t idesc2 = (1u << 4) | (1u << 16) | ((uint32_t)(FK >> 3) << 17) |
                                    ((uint32_t)(FM >> 4) << 24);
            mbar_wait(bar(D_FULL), 0);
            tc_fence_after();
            const uint32_t aD = smem_u32(sD);
            auto mma2 = [&](int n) {   // dd += da(n) . W(n)
                const int st = n % NW;
                const int ab = n % NA;
                mbar_wait(bar(A_FULL + ab), (uint32_t)((n / NA) & 1));
                FUSED_DBG(n, 1);
                tc_fence_after();
#pragma unroll
                for (int h = 0; h < P; ++h) {
                    const uint32_t aA = smem_u32(sA + (ab * P + h) * FABytes);
                    const uint32_t aW = smem_u32(sW + (st * P + h) * FWBytes);
#pragma unroll
                    for (int j = 0; j < FG / 16; ++j) {
                        const uint64_t da = make_desc(aA + 32 * j, 16, 1024, 2);
                        const uint64_t db = make_desc(aW + 2048 * j, FWBytes / 2, 1024, 2);
                        tc_mma_f16(tmem_DD, da, db, idesc2, (n > 0 || h > 0 || j > 0) ? 1u : 0u);
                    }
                }
                tc_commit(bar(A_EMPTY + ab));
                tc_commit(bar(W_EMPTY + st));
            };
            for (int n = 0; n < ntile; ++n) {
                const int st = n & 1;
                const uint32_t ph = (n >> 1) & 1;
                const int ws = n % NW;
                mbar_wait(bar(W_FULL + ws), (uint32_t)((n / NW) & 1));
                mbar_wait(bar(S_EMPTY + st), ph ^ 1);
                FUSED_DBG(n, 0);
                tc_fence_after();
#pragma unroll
                for (int h = 0; h < P; ++h) {
                    const uint32_t aW = smem_u32(sW + (ws * P + h) * FWBytes);
                    const uint32_t dS = tmem_S + (st * P + h) * FG;
#pragma unroll
                    for (int kk = 0; kk < FK / 16; ++kk) {
                        const uint32_t off = (kk >> 2) * (FDBytes / 2) + (kk & 3) * 32;
                        const uint32_t offw = (kk >> 2) * (FWBytes / 2) + (kk & 3) * 32;
                        tc_mma_f16(dS, make_desc(aD + off, 16, 1024, 2), make_desc(aW + offw, 16, 1024, 2),
                                   idesc1, kk > 0 ? 1u : 0u);
                    }
                }
                tc_commit(bar(S_FULL + st));
                if (BWD) {
                    if (n > 0) mma2(n - 1);
                } else {
                    tc_commit(bar(W_EMPTY + ws));     // forward only: the weights are done with MMA1
                }
            }
            if (BWD) {
                mma2(ntile - 1);
                tc_commit(bar(DD_FULL));
            }
        }
    } else if (warp == 3) {
        // ===== da store warp: smem da tile -> HBM (fp16), releases the tile for the next epilogue
        if (BWD && lane == 0) {
            for (int n = 0; n < ntile; ++n) {
                const int g0 = (tile0 + n) * FG;
                const int ab = n % NA;
                mbar_wait(bar(A_FULL + ab), (uint32_t)((n / NA) & 1));
                FUSED_DBG(n, 2);
#pragma unroll
                for (int h = 0; h < P; ++h)
                    tma_store_2d(&tmDA, (int)(h * p.head_stride) + g0, row0,
                                 smem_u32(sA + (ab * P + h) * FABytes));
                tma_commit();
                tma_wait_read<0>();
                FUSED_DBG(n, 3);
                mbar_arrive(bar(A_STORED + ab));
            }
            tma_wait_all();
        }
    } else if (warp >= 4) {
        // ===== epilogue =====
        const int e = warp - 4;
        const int q = e & 3;                     // TMEM lane quadrant (== warp % 4)
        const int ej = e >> 2;                   // index among the EJ warps sharing this quadrant
        const int half = ej;                     // (dd epilogue: warps ej < 2 take 64 columns each)
        const int row = q * 32 + lane;
        const int grow = row0 + row;
        const float gs_row = (grow < p.M ? (p.go ? p.go[grow] : p.go_scalar) : 0.f) * p.scale;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const uint32_t rscr = smem_u32(sR + e * FRBytes);            // this warp's total_count chunk
        const uint32_t a_row0 = smem_u32(sA) + (uint32_t)row * 128u;  // this row of da buffer 0, head 0
        const bool fixups = Lik<KIND>::NB || p.has_const == 0;       // any x >= 2 special function left?
        float accA = 0.f, accB = 0.f;      // log p of this row = accA - ln2 * accB
        for (int n = 0; n < ntile; ++n) {
            const int st = n & 1;
            const uint32_t ph = (n >> 1) & 1;
            const int g0 = (tile0 + n) * FG;
            mbar_wait(bar(S_FULL + st), ph);
            mbar_wait(bar(T_FULL + st), ph);
            if (lane == 0 && (e == 0 || e == EW - 1)) FUSED_DBG(n, e == 0 ? 4 : 10);
            tc_fence_after();
            const uint32_t trow = smem_u32(sT + st * FTBytes) + (uint32_t)row * 128u;
            const int ab = n % NA;
            const uint32_t a_row = a_row0 + (uint32_t)(ab * P * FABytes);
            bool first_write = true;
#pragma unroll 1
            for (int sub = ej; sub < 4; sub += EJ) {           // EJ divides 4: same chunks every tile
                const int gc = sub * 16;                       // first gene of this 16-gene chunk
                uint32_t sv[3][16];
#pragma unroll
                for (int h = 0; h < P; ++h) tc_ld16(tmem_S + lane_addr + (st * P + h) * FG + gc, sv[h]);
                const int c0 = gc >> 3;                         // 16-byte chunk index in the row
                const uint32_t sw0 = (uint32_t)(((c0) ^ (row & 7)) << 4);
                const uint32_t sw1 = (uint32_t)(((c0 + 1) ^ (row & 7)) << 4);
                const uint4 t0 = lds_v4(trow + sw0);
                const uint4 t1 = lds_v4(trow + sw1);
                const uint32_t tw[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
                tc_wait_ld();
                // is any clip of the reference active in this chunk? (fused_math.cuh)
                float mlogit = 0.f, mlog = 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
#pragma unroll
                    for (int h = 0; h < P - 1; ++h) mlogit = fmaxf(mlogit, fabsf(__uint_as_float(sv[h][j])));
                    mlog = fmaxf(mlog, fabsf(__uint_as_float(sv[P - 1][j])));
                }
                const bool slow = (mlogit > kFastLogitMax) || (mlog > kFastLogMax);
                uint32_t packed[3][8];
                uint32_t flags = 0;      // element 2i -> bit 14 - i, element 2i + 1 -> bit 30 - i: target >= 2
                if (!slow) {
#pragma unroll
                    for (int blk = 0; blk < 2; ++blk) {      // two blocks of 8 genes
                        // genes >= G only exist in the last tile; G % 8 == 0 keeps blocks uniform
                        if ((g0 + gc + blk * 8) >= p.G) {
#pragma unroll
                            for (int h = 0; h < P; ++h)
#pragma unroll
                                for (int j = 0; j < 4; ++j) packed[h][blk * 4 + j] = 0u;
                            continue;
                        }
                        float x[8], av[3][8], gv[3][8], rr[8];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint32_t w = tw[blk * 4 + i];
                            fused_cvt2<T_HALF>(w, x[2 * i], x[2 * i + 1]);
                            flags |= fused_flags2<T_HALF>(w) >> (blk * 4 + i);
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j)
#pragma unroll
                            for (int h = 0; h < P; ++h) av[h][j] = __uint_as_float(sv[h][blk * 8 + j]);
                        if (BWD) fused_fast8<KIND>(x, av, gs_row, accA, accB, rr, gv);
                        else fused_fwd8<KIND>(x, av, accA, accB, rr);
                        if (Lik<KIND>::NB) {
                            // fp16 is ample for the x >= 2 corrections: their error is ~1e-3 x per
                            // term with random sign, against a row sum of thousands
                            sts_v4(rscr + (uint32_t)((blk * 32 + lane) * 16), pack_half2(rr[0], rr[1]),
                                   pack_half2(rr[2], rr[3]), pack_half2(rr[4], rr[5]), pack_half2(rr[6], rr[7]));
                        }
                        if (BWD) {
#pragma unroll
                            for (int h = 0; h < P; ++h)
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    packed[h][blk * 4 + j] = pack_half2(gv[h][2 * j], gv[h][2 * j + 1]);
                        }
                    }
                } else {
                    FusedSlowIn in;
#pragma unroll
                    for (int i = 0; i < 8; ++i) in.tw[i] = tw[i];
#pragma unroll
                    for (int h = 0; h < P; ++h)
#pragma unroll
                        for (int j = 0; j < 16; ++j) in.sv[h][j] = sv[h][j];
                    const FusedSlowOut out =
                        fused_slow_chunk<KIND, T_HALF, BWD>(in, gs_row, g0 + gc, p.G, p.has_const != 0);
                    accA += out.acc;
#pragma unroll
                    for (int h = 0; h < P; ++h)
#pragma unroll
                        for (int j = 0; j < 8; ++j) packed[h][j] = out.packed[h][j];
                }
                if (BWD && first_write && n >= NA) {
                    // the previous tile in this da buffer must have been consumed by MMA2 and read
                    // by its TMA store before it is overwritten
                    if (lane == 0 && e == 0) FUSED_DBG(n, 5);
                    mbar_wait(bar(A_EMPTY + ab), (uint32_t)(((n - NA) / NA) & 1));
                    if (lane == 0 && e == 0) FUSED_DBG(n, 6);
                    mbar_wait(bar(A_STORED + ab), (uint32_t)(((n - NA) / NA) & 1));
                    if (lane == 0 && e == 0) FUSED_DBG(n, 7);
                }
                first_write = false;
                if (BWD) {
#pragma unroll
                    for (int h = 0; h < P; ++h) {
                        const uint32_t arow = a_row + (uint32_t)(h * FABytes);
                        sts_v4(arow + sw0, packed[h][0], packed[h][1], packed[h][2], packed[h][3]);
                        sts_v4(arow + sw1, packed[h][4], packed[h][5], packed[h][6], packed[h][7]);
                    }
                }
                // targets >= 2 (about 3 % of a single-cell matrix): lgamma / digamma differences,
                // one trip per flagged element of this row; the fp16 gradient is patched in place
                if (fixups) {
                    // two flagged elements per trip, straight-line: the two latency chains overlap
                    while (flags) {
                        int b[2];
                        b[0] = 31 - __clz(flags);
                        flags &= ~(1u << b[0]);
                        const bool two = flags != 0;
                        b[1] = two ? 31 - __clz(flags) : b[0];
                        flags &= ~(1u << b[1]);
                        uint32_t off[2];
                        float xj[2], rj[2], g_old[2], D[2], Pd[2];
                        int jj[2];
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const int hi = b[k] >> 4;
                            jj[k] = 2 * ((hi ? 30 : 14) - b[k]) + hi;      // element of the chunk
                            off[k] = ((jj[k] & 8) ? sw1 : sw0) + (uint32_t)((jj[k] & 7) << 1);
                        }
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const uint16_t bits = lds_u16(trow + off[k]);
                            xj[k] = T_HALF ? __half2float(__ushort_as_half(bits)) : (float)bits;
                            if (Lik<KIND>::NB) {
                                rj[k] = __half2float(__ushort_as_half(lds_u16(
                                    rscr + (uint32_t)((((jj[k] >> 3) * 32 + lane) * 16) + ((jj[k] & 7) << 1)))));
                                g_old[k] = BWD ? __half2float(__ushort_as_half(lds_u16(
                                                     a_row + (uint32_t)((P - 1) * FABytes) + off[k])))
                                               : 0.f;
                            }
                        }
                        float extra = 0.f;
                        if (Lik<KIND>::NB) {
#pragma unroll
                            for (int k = 0; k < 2; ++k) lgamma_diff_prod(rj[k], xj[k], D[k], Pd[k]);
                            if (fmaxf(xj[0], xj[1]) > (float)kProdMax) {      // rare: large counts
#pragma unroll
                                for (int k = 0; k < 2; ++k) lgamma_diff_ge2(rj[k], xj[k], D[k], Pd[k]);
                            }
                            extra = D[0] + (two ? D[1] : 0.f);
                            if (BWD)
                            sts_u16(a_row + (uint32_t)((P - 1) * FABytes) + off[0],
                                    __half_as_ushort(__float2half_rn(fmaf(rj[0] * Pd[0], gs_row, g_old[0]))));
                            if (BWD && two)
                                sts_u16(a_row + (uint32_t)((P - 1) * FABytes) + off[1],
                                        __half_as_ushort(__float2half_rn(fmaf(rj[1] * Pd[1], gs_row, g_old[1]))));
                        }
                        if (!p.has_const) extra -= lgamma1p_ge2(xj[0]) + (two ? lgamma1p_ge2(xj[1]) : 0.f);
                        accA += extra;
                    }
                }
            }
            // S(n) and t(n) are in registers / consumed: release them
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar(S_EMPTY + st));
                mbar_arrive(bar(T_EMPTY + st));
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0 && (e == 0 || e == EW - 1)) FUSED_DBG(n, e == 0 ? 8 : 11);
            if (BWD && lane == 0) mbar_arrive(bar(A_FULL + ab));
        }
        // ---- log p partial of this gene range: combine the warps of a quadrant, fixed order ----
        // (the partials reuse the first 128 bytes of each warp's own total_count chunk)
        __syncwarp();
        sts_f32(rscr + (uint32_t)lane * 4u, fmaf(-kLn2, accB, accA));
        bar_sync_n(1, 32 * EW);
        if (ej == 0 && grow < p.M) {
            float tot = 0.f;
#pragma unroll
            for (int j = 0; j < EJ; ++j) tot += reinterpret_cast<const float *>(sR + (j * 4 + q) * FRBytes)[lane];
            p.logp_part[(int64_t)gs * p.part_stride + grow] = tot;
        }
        // ---- dd partial: TMEM -> staging smem (the t stages) -> TMA reduce-add ----
        if (BWD) mbar_wait(bar(DD_FULL), 0);
        tc_fence_after();
        uint8_t *stage = sT + half * FTBytes;       // 128 rows x 32 fp32 columns
        const bool half_issuer = (lane == 0 && q == 0);
#pragma unroll 1
        for (int c = 0; BWD && c < 2 && ej < 2; ++c) {
            uint32_t v[32];
            tc_ld32(tmem_DD + lane_addr + half * 64 + c * 32, v);
            tc_wait_ld();
            if (c > 0) {
                if (half_issuer) tma_wait_read<0>();
                bar_sync_n(2 + half, 128);
            }
            uint8_t *dst = stage + row * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint4 val = make_uint4(
                    __float_as_uint(__uint_as_float(v[4 * j]) * p.inv_scale),
                    __float_as_uint(__uint_as_float(v[4 * j + 1]) * p.inv_scale),
                    __float_as_uint(__uint_as_float(v[4 * j + 2]) * p.inv_scale),
                    __float_as_uint(__uint_as_float(v[4 * j + 3]) * p.inv_scale));
                *reinterpret_cast<uint4 *>(dst + ((j ^ (row & 7)) << 4)) = val;
            }
            fence_async_smem();
            bar_sync_n(2 + half, 128);
            if (half_issuer) {
                // partial of this gene range -> its own workspace slice; the finish kernel sums
                // the slices in fixed order (deterministic, unlike a reduce-add into dd)
                tma_store_2d(&tmDD, half * 64 + c * 32, gs * (int)p.part_stride + row0, smem_u32(stage));
                tma_commit();
            }
        }
        if (BWD && half_issuer && ej < 2) tma_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

// One CTA per cell m: logp[m] = sum_s part[s][m] - row_const[m % t_rows], and (backward)
// dd[m][c] = sum_s dd_part[s][m][c] for c < dd_cols, 0 for the padding columns [dd_cols, lddd).
// The gene-range partials s are summed in a FIXED order (8 interleaved lanes of partial sums, then a
// fixed tree), so the result does not depend on the launch.  256 threads = 32 column quads x 8 lanes
// of s: with the ~150 gene ranges of a small minibatch a thread adds ~20 partials, not 150.
__global__ void __launch_bounds__(256)
fused_finish_kernel(const float *__restrict__ part, int64_t part_stride, int gsplit, int M,
                    const float *__restrict__ row_const, int t_rows, float *__restrict__ logp,
                    const float *__restrict__ dd_part, float *__restrict__ dd, int64_t lddd, int dd_cols) {
    __shared__ float4 red[8][32];
    __shared__ float red_lp[8];
    const int m = blockIdx.x;
    const int cq = threadIdx.x & 31, sl = threadIdx.x >> 5;
    // log p: thread t adds s = t, t + 256, ...
    float lp = 0.f;
    for (int s = threadIdx.x; s < gsplit; s += 256) lp += part[(int64_t)s * part_stride + m];
    lp = warp_sum(lp);
    if (cq == 0) red_lp[sl] = lp;
    const int c = cq << 2;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (dd_part) {
        const float *src = dd_part + (int64_t)m * FK + c;
        const int64_t step = part_stride * FK;
        int s = sl;
        for (; s + 24 < gsplit; s += 32) {          // four loads in flight
            const float4 v0 = *reinterpret_cast<const float4 *>(src + (int64_t)s * step);
            const float4 v1 = *reinterpret_cast<const float4 *>(src + (int64_t)(s + 8) * step);
            const float4 v2 = *reinterpret_cast<const float4 *>(src + (int64_t)(s + 16) * step);
            const float4 v3 = *reinterpret_cast<const float4 *>(src + (int64_t)(s + 24) * step);
            acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
            acc.x += v1.x; acc.y += v1.y; acc.z += v1.z; acc.w += v1.w;
            acc.x += v2.x; acc.y += v2.y; acc.z += v2.z; acc.w += v2.w;
            acc.x += v3.x; acc.y += v3.y; acc.z += v3.z; acc.w += v3.w;
        }
        for (; s < gsplit; s += 8) {
            const float4 v = *reinterpret_cast<const float4 *>(src + (int64_t)s * step);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        red[sl][cq] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) t += red_lp[j];
        logp[m] = t - (row_const ? row_const[m % t_rows] : 0.f);
    }
    if (dd_part && sl == 0 && c < lddd) {
        float4 t = red[0][cq];
#pragma unroll
        for (int j = 1; j < 8; ++j) {
            const float4 v = red[j][cq];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        if (c + 0 >= dd_cols) t.x = 0.f;
        if (c + 1 >= dd_cols) t.y = 0.f;
        if (c + 2 >= dd_cols) t.z = 0.f;
        if (c + 3 >= dd_cols) t.w = 0.f;
        *reinterpret_cast<float4 *>(dd + (int64_t)m * lddd + c) = t;     // lddd % 4 == 0
    }
}

// The same reduction for few gene ranges and many cells (the training minibatch: gsplit ~ 9, M = 4096):
// one thread per (cell, column quad), the partials added in s order with all loads of a thread in
// flight; the lane of column quad 0 also adds the cell's log p partials.  (The CTA-per-cell form above
// spends its time in block scheduling when there is next to nothing to add per cell.)
__global__ void __launch_bounds__(256)
fused_finish_wide_kernel(const float *__restrict__ part, int64_t part_stride, int gsplit, int M,
                         const float *__restrict__ row_const, int t_rows, float *__restrict__ logp,
                         const float *__restrict__ dd_part, float *__restrict__ dd, int64_t lddd, int dd_cols) {
    pdl_wait();          // (programmatic dependent launch: common.cuh)
    pdl_trigger();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = (int)(i & 31) << 2;
    const int64_t m = i >> 5;
    if (m >= M) return;
    if (c == 0) {
        float lp = 0.f;
#pragma unroll 4
        for (int s = 0; s < gsplit; ++s) lp += part[(int64_t)s * part_stride + m];
        logp[m] = lp - (row_const ? row_const[m % t_rows] : 0.f);
    }
    if (!dd_part || c >= lddd) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int s = 0; s < gsplit; ++s) {
        const float4 v = *reinterpret_cast<const float4 *>(dd_part + ((int64_t)s * part_stride + m) * FK + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (c + 0 >= dd_cols) acc.x = 0.f;
    if (c + 1 >= dd_cols) acc.y = 0.f;
    if (c + 2 >= dd_cols) acc.z = 0.f;
    if (c + 3 >= dd_cols) acc.w = 0.f;
    *reinterpret_cast<float4 *>(dd + m * lddd + c) = acc;     // lddd % 4 == 0
}

static inline int make_map_u16(CUtensorMap *map, const void *base, int64_t rows, int64_t cols, int64_t ld,
                               int box_cols, int box_rows) {
    EncodeTiledFn enc = get_encode();
    SCVAE_CHECK_ARG(enc, "heads_fused: cuTensorMapEncodeTiled unavailable");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, (void *)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SCVAE_CHECK_ARG(r == CUDA_SUCCESS, "heads_fused: tensor map (u16) failed (%d)", (int)r);
    return 0;
}

static long long *g_fused_dbg = nullptr;

struct FusedPlan {
    int row_tiles, n_tiles, gsplit, tiles_per_cta;
};
static FusedPlan fused_plan(int M, int G) {
    FusedPlan f;
    f.row_tiles = (M + FM - 1) / FM;
    f.n_tiles = (G + FG - 1) / FG;
    static int waves = 0;                                      // CTA waves over the 148 SMs (default 2)
    if (!waves) {
        const char *e = getenv("SCVAE_FUSED_WAVES");
        waves = (e && atoi(e) > 0) ? atoi(e) : 2;
    }
    int target = (waves * 148 + f.row_tiles / 2) / f.row_tiles;
    if (target < 1) target = 1;
    if (target > f.n_tiles) target = f.n_tiles;
    f.tiles_per_cta = (f.n_tiles + target - 1) / target;
    f.gsplit = (f.n_tiles + f.tiles_per_cta - 1) / f.tiles_per_cta;
    return f;
}

template <int KIND, bool T_HALF, bool BWD>
static int launch_fused_t(const void *d16, const void *w16, const void *t16, int64_t ldt, int t_rows,
                        int M, int G,
                        int64_t head_stride, const float *go, float go_scalar, float scale, void *da16,
                        float *dd, int64_t lddd, int dd_cols, float *logp_part, const float *row_const,
                        float *logp, cudaStream_t s) {
    constexpr int P = Lik<KIND>::P;
    const FusedPlan f = fused_plan(M, G);
    CUtensorMap tmD, tmW, tmT, tmDA, tmDD;
    if (make_map(&tmD, d16, M, FK, FK, 64, FM, CU_TENSOR_MAP_SWIZZLE_128B, 2)) return 1;
    if (make_map(&tmW, w16, (int64_t)P * head_stride, FK, FK, 64, FG, CU_TENSOR_MAP_SWIZZLE_128B, 2)) return 1;
    if (make_map_u16(&tmT, t16, t_rows, G, ldt, FG, FM)) return 1;
    if (BWD) {
        if (make_map(&tmDA, da16, M, (int64_t)P * head_stride, (int64_t)P * head_stride, FG, FM,
                     CU_TENSOR_MAP_SWIZZLE_128B, 2))
            return 1;
        // dd partials: one (rows_pad x 128) fp32 slice per gene range, behind the log p partials
        if (make_map(&tmDD, logp_part + (int64_t)f.gsplit * f.row_tiles * FM, (int64_t)f.gsplit * f.row_tiles * FM,
                     FK, FK, 32, FM))
            return 1;
    } else {
        tmDA = tmD;      // unused by the forward-only kernel
        tmDD = tmD;
    }
    FusedParams p;
    p.M = M; p.G = G; p.t_rows = t_rows;
    p.tiles_per_cta = f.tiles_per_cta; p.n_tiles = f.n_tiles; p.gsplit = f.gsplit;
    p.head_stride = head_stride;
    p.part_stride = (int64_t)f.row_tiles * FM;
    p.go = go; p.go_scalar = go_scalar; p.scale = scale; p.inv_scale = 1.f / scale;
    p.logp_part = logp_part;
    p.has_const = row_const != nullptr;
    p.dbg = g_fused_dbg;
    constexpr int smem = fused_smem_bytes(P);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(heads_fused_kernel<KIND, T_HALF, BWD>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        SCVAE_CHECK_ARG(e == cudaSuccess, "heads_fused: cannot set smem attribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    launch_pdl(kPdlHeads, heads_fused_kernel<KIND, T_HALF, BWD>, dim3(f.row_tiles * f.gsplit), dim3(fused_threads(P)), smem, s,
               tmD, tmW, tmT, tmDA, tmDD, p);
    SCVAE_CHECK_LAUNCH("heads_fused");
    const float *dd_part = BWD ? logp_part + (int64_t)f.gsplit * f.row_tiles * FM : nullptr;
    if (f.gsplit <= 16)
        launch_pdl(kPdlFinish, fused_finish_wide_kernel, dim3((unsigned)(((int64_t)M * 32 + 255) / 256)), dim3(256), 0, s,
                   (const float *)logp_part, p.part_stride, f.gsplit, M, row_const, t_rows, logp, dd_part, dd, lddd, dd_cols);
    else
        fused_finish_kernel<<<M, 256, 0, s>>>(logp_part, p.part_stride, f.gsplit, M, row_const, t_rows, logp, dd_part,
                                              dd, lddd, dd_cols);
    SCVAE_CHECK_LAUNCH("heads_fused_finish");
    return 0;
}

template <int KIND>
static int launch_fused(const void *d16, const void *w16, const void *t16, int64_t ldt, int t_is_half, int t_rows,
                        int M, int G, int64_t head_stride, const float *go, float go_scalar, float scale,
                        void *da16, float *dd, int64_t lddd, int dd_cols, float *logp_part,
                        const float *row_const, float *logp, cudaStream_t s) {
    const bool bwd = da16 != nullptr;     // forward only when no gradient buffers are given
#define FUSED_GO(TH, BW)                                                                                          \
    return launch_fused_t<KIND, TH, BW>(d16, w16, t16, ldt, t_rows, M, G, head_stride, go, go_scalar, scale, da16, \
                                        dd, lddd, dd_cols, logp_part, row_const, logp, s)
    if (t_is_half) {
        if (bwd) FUSED_GO(true, true);
        FUSED_GO(true, false);
    }
    if (bwd) FUSED_GO(false, true);
    FUSED_GO(false, false);
#undef FUSED_GO
}

}  // namespace scvae

using namespace scvae;

// development aid (not part of the public header): device buffer of 40 x 16 int64 that receives
// the clock64 timeline of CTA 0 of subsequent launches; NULL switches it off
extern "C" void scvae_heads_fused_debug(void *buf) { g_fused_dbg = (long long *)buf; }

extern "C" int64_t scvae_heads_fused_workspace_floats(int M, int G) {
    if (M <= 0 || G <= 0) return 0;
    const FusedPlan f = fused_plan(M, G);
    // per gene range: log p partials (rows_pad) + decoder-gradient partials (rows_pad x 128)
    return (int64_t)f.gsplit * f.row_tiles * FM * (1 + FK);
}

extern "C" int scvae_heads_fused_fwd(int kind, const void *d16, const void *w16, int64_t head_stride,
                                     const void *t16, int64_t ldt, int t_is_half, int t_rows, int M, int G,
                                     const float *row_const, float *logp, float *workspace, void *stream) {
    SCVAE_CHECK_ARG(d16 && w16 && t16 && logp && workspace, "heads_fused_fwd: NULL pointer");
    SCVAE_CHECK_ARG(M > 0 && G > 0 && G % 8 == 0 && t_rows > 0, "heads_fused_fwd: bad shape (G must be a multiple of 8)");
    SCVAE_CHECK_ARG(head_stride % 64 == 0 && head_stride >= G, "heads_fused_fwd: head_stride must be a multiple of 64");
    SCVAE_CHECK_ARG(ldt % 8 == 0, "heads_fused_fwd: bad leading dimension");
    SCVAE_CHECK_ARG(M == t_rows || t_rows % FM == 0,
                    "heads_fused_fwd: targets must tile in multiples of 128 rows (t_rows=%d, M=%d)", t_rows, M);
    cudaStream_t s = (cudaStream_t)stream;
    switch (kind) {
#define CASE(KK)                                                                                              \
    case KK:                                                                                                  \
        return launch_fused<KK>(d16, w16, t16, ldt, t_is_half, t_rows, M, G, head_stride, nullptr, 0.f, 1.f, nullptr, \
                                nullptr, 0, 0, workspace, row_const, logp, s);
        CASE(SCVAE_LIK_POISSON)
        CASE(SCVAE_LIK_NB)
        CASE(SCVAE_LIK_ZIP)
        CASE(SCVAE_LIK_ZINB)
#undef CASE
    }
    set_error("heads_fused_fwd: unknown kind %d", kind);
    return 1;
}

extern "C" int scvae_heads_fused_bwd(int kind, const void *d16, const void *w16, int64_t head_stride,
                                     const void *t16, int64_t ldt, int t_is_half, int t_rows, int M, int G,
                                     const float *row_const, const float *go, float go_scalar, float scale,
                                     void *da16, float *dd, int64_t lddd, int dd_cols, float *logp,
                                     float *workspace, void *stream) {
    SCVAE_CHECK_ARG(d16 && w16 && t16 && da16 && dd && logp && workspace, "heads_fused_bwd: NULL pointer");
    SCVAE_CHECK_ARG(M > 0 && G > 0 && G % 8 == 0 && t_rows > 0, "heads_fused_bwd: bad shape (G must be a multiple of 8)");
    SCVAE_CHECK_ARG(head_stride % 64 == 0 && head_stride >= G, "heads_fused_bwd: head_stride must be a multiple of 64");
    SCVAE_CHECK_ARG(ldt % 8 == 0 && lddd % 4 == 0 && dd_cols <= FK, "heads_fused_bwd: bad leading dimensions");
    SCVAE_CHECK_ARG(M == t_rows || t_rows % FM == 0,
                    "heads_fused_bwd: targets must tile in multiples of 128 rows (t_rows=%d, M=%d)", t_rows, M);
    SCVAE_CHECK_ARG(scale > 0.f, "heads_fused_bwd: scale must be positive");
    cudaStream_t s = (cudaStream_t)stream;
    switch (kind) {
#define CASE(KK)                                                                                              \
    case KK:                                                                                                  \
        return launch_fused<KK>(d16, w16, t16, ldt, t_is_half, t_rows, M, G, head_stride, go, go_scalar, scale, da16, dd, \
                                lddd, dd_cols, workspace, row_const, logp, s);
        CASE(SCVAE_LIK_POISSON)
        CASE(SCVAE_LIK_NB)
        CASE(SCVAE_LIK_ZIP)
        CASE(SCVAE_LIK_ZINB)
#undef CASE
    }
    set_error("heads_fused_bwd: unknown kind %d", kind);
    return 1;
}
