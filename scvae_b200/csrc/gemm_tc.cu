// tcgen05 (5th-gen tensor core) GEMM for the dense layers that touch the gene axis:
// encoder layer 1 (cells x genes)(genes x H), the likelihood-parameter heads (cells x H)(H x P*genes)
// and their dgrad / wgrad products (MU:53-59 fully_connected and its autodiff gradients).
//
//   * operands are fp32 in HBM and are consumed as kind::tf32 (no conversion pass, no extra
//     copy): TMA (cp.async.bulk.tensor, SWIZZLE_128B) stages 128 x 32 tiles into a 5-deep
//     shared-memory ring, one elected thread issues tcgen05.mma 128x128x8 into TMEM, four
//     epilogue warps drain the double-buffered accumulator with tcgen05.ld and write it back
//     through swizzled shared memory with TMA stores (or TMA reduce-add when accumulating);
//   * all three products of a dense layer use the SAME row-major buffers: forward is
//     K-major x K-major, dgrad K-major x MN-major, wgrad MN-major x MN-major (UMMA descriptors
//     with the transpose bits), so no transposed copies of activations, gradients or weights
//     ever exist in HBM;
//   * persistent CTAs (<= one per SM) walk (m-tile, n-tile, k-split) work items; skinny
//     products (N ~ 100, K ~ 20000) are split along K into a workspace and reduced in a fixed
//     order by a second tiny kernel, so results are deterministic.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace scvae {

constexpr int BM = 128, BN = 128, BK = 32;       // BK fp32 = 128 bytes = one swizzle row
constexpr int kStages = 5;
constexpr int kTileBytes = BM * BK * 4;          // 16 KB per operand per stage
constexpr int kStageBytes = 2 * kTileBytes;
constexpr int kEpiBytes = BM * 32 * 4;           // 128 rows x 32 fp32 columns
constexpr int kSmemBytes = kStages * kStageBytes + 2 * kEpiBytes + 1024 /*align*/ + 256 /*barriers*/;
constexpr int kTmemCols = 2 * BN;                // double-buffered fp32 accumulator
constexpr int kThreads = 256;

// ---- PTX wrappers -----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int c0, int c1, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map),
                 "r"(c0), "r"(c1), "r"(src)
                 : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap *map, int c0, int c1, uint32_t src) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map),
                 "r"(c0), "r"(c1), "r"(src)
                 : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor), version 1.  layout_type:
// 2 = SWIZZLE_128B (16-byte chunks, K-major operands), 1 = SWIZZLE_128B_BASE32B (32-byte
// chunks: the only layout the tensor core accepts for MN-major 32-bit operands).
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    d |= (uint64_t)layout_type << 61;
    return d;
}

struct GemmParams {
    int M, N, K;
    int tiles_m, tiles_n, nsplit, kb_per_split, nkb;
    int accumulate;     // TMA reduce-add into C (only when nsplit == 1)
    int ws_rows;        // rows per split slice of the workspace (multiple of BM)
    uint32_t mn_lbo, mn_sbo;  // descriptor strides of MN-major operand tiles
};

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *epi = smem + kStages * kStageBytes;
    uint64_t *bars = (uint64_t *)(epi + 2 * kEpiBytes);
    // bars: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], then tmem base address
    const uint32_t bar_full = smem_u32(bars);
    const uint32_t bar_empty = smem_u32(bars + kStages);
    const uint32_t bar_tfull = smem_u32(bars + 2 * kStages);
    const uint32_t bar_tempty = smem_u32(bars + 2 * kStages + 2);
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int total = p.tiles_m * p.tiles_n * p.nsplit;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < total; item += gridDim.x) {
                const int tm = item % p.tiles_m;
                const int tn = (item / p.tiles_m) % p.tiles_n;
                const int sp = item / (p.tiles_m * p.tiles_n);
                const int kb0 = sp * p.kb_per_split;
                const int kb1 = min(kb0 + p.kb_per_split, p.nkb);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t sa = smem_u32(smem + stage * kStageBytes);
                    const uint32_t sb = sa + kTileBytes;
                    const uint32_t full = bar_full + 8 * stage;
                    mbar_expect_tx(full, kStageBytes);
                    if (A_MN) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) tma_load_2d(sa + j * 4096, &tmA, tm * BM + 32 * j, kb * BK, full);
                    } else {
                        tma_load_2d(sa, &tmA, kb * BK, tm * BM, full);
                    }
                    if (B_MN) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) tma_load_2d(sb + j * 4096, &tmB, tn * BN + 32 * j, kb * BK, full);
                    } else {
                        tma_load_2d(sb, &tmB, kb * BK, tn * BN, full);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, M=128, N=128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) |
                                   ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)(BM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int t = 0;
            for (int item = blockIdx.x; item < total; item += gridDim.x, ++t) {
                const int sp = item / (p.tiles_m * p.tiles_n);
                const int kb0 = sp * p.kb_per_split;
                const int kb1 = min(kb0 + p.kb_per_split, p.nkb);
                const int acc = t & 1;
                const uint32_t acc_phase = (t >> 1) & 1;
                mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * kStageBytes);
                    const uint32_t sb = sa + kTileBytes;
                    const uint64_t da = A_MN ? make_desc(sa, p.mn_lbo, p.mn_sbo, 1) : make_desc(sa, 16, 1024, 2);
                    const uint64_t db = B_MN ? make_desc(sb, p.mn_lbo, p.mn_sbo, 1) : make_desc(sb, 16, 1024, 2);
                    // per UMMA_K = 8 fp32 step: K-major +32 B inside the swizzle row; MN-major +8 K-rows
                    const uint64_t sta = A_MN ? (1024 >> 4) : (32 >> 4);
                    const uint64_t stb = B_MN ? (1024 >> 4) : (32 >> 4);
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k)
                        tc_mma_tf32(tmem_d, da + k * sta, db + k * stb, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    tc_commit(bar_empty + 8 * stage);  // frees the smem slot once the MMAs retire
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                tc_commit(bar_tfull + 8 * acc);        // accumulator complete
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> swizzled smem -> TMA store =====
        const int q = warp - 4;                 // TMEM lane quadrant of this warp
        const int row = q * 32 + lane;          // row inside the 128-row tile
        const bool issuer = (threadIdx.x == 128);
        int t = 0;
        int buf = 0;
        for (int item = blockIdx.x; item < total; item += gridDim.x, ++t) {
            const int tm = item % p.tiles_m;
            const int tn = (item / p.tiles_m) % p.tiles_n;
            const int sp = item / (p.tiles_m * p.tiles_n);
            const int acc = t & 1;
            const uint32_t acc_phase = (t >> 1) & 1;
            mbar_wait(bar_tfull + 8 * acc, acc_phase);
            tc_fence_after();
            const int out_row = (p.nsplit > 1 ? sp * p.ws_rows : 0) + tm * BM;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                if (tn * BN + c * 32 >= p.N) break;   // uniform: nothing to store
                uint32_t v[32];
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + c * 32, v);
                tc_wait_ld();
                if (issuer) tma_wait_read<1>();       // staging buffer `buf` is free again
                epi_bar_sync();
                uint8_t *dst = epi + buf * kEpiBytes + row * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint4 val = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    *reinterpret_cast<uint4 *>(dst + ((j ^ (row & 7)) << 4)) = val;
                }
                fence_async_smem();
                epi_bar_sync();
                if (issuer) {
                    const uint32_t src = smem_u32(epi + buf * kEpiBytes);
                    if (p.accumulate && p.nsplit == 1)
                        tma_reduce_add_2d(&tmC, tn * BN + c * 32, out_row, src);
                    else
                        tma_store_2d(&tmC, tn * BN + c * 32, out_row, src);
                    tma_commit();
                }
                buf ^= 1;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
        }
        if (issuer) tma_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols)
                     : "memory");
    }
}

// C[m, n] (+)= sum_s ws[s][m][n], fixed order.
__global__ void splitk_reduce_kernel(const float *__restrict__ ws, int64_t ldw, int ws_rows, int nsplit, int M,
                                     int N, float *__restrict__ C, int64_t ldc, int accumulate) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y;
    if (n >= N) return;
    float acc = 0.f;
    for (int s = 0; s < nsplit; ++s) acc += ws[((int64_t)s * ws_rows + m) * ldw + n];
    float *c = C + (int64_t)m * ldc + n;
    *c = accumulate ? *c + acc : acc;
}

// ---- host side ----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

// 2-D fp32 tensor map over a row-major (rows, cols) matrix with leading dimension ld.
static int make_map(CUtensorMap *map, const float *base, int64_t rows, int64_t cols, int64_t ld, int box_cols,
                    int box_rows, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn enc = get_encode();
    SCVAE_CHECK_ARG(enc, "gemm_tf32: cuTensorMapEncodeTiled unavailable");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SCVAE_CHECK_ARG(r == CUDA_SUCCESS, "gemm_tf32: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld",
                    (int)r, (long long)rows, (long long)cols, (long long)ld);
    return 0;
}

struct SplitPlan {
    int tiles_m, tiles_n, nkb, nsplit, kb_per_split;
};

static SplitPlan plan_split(int M, int N, int K) {
    SplitPlan s;
    s.tiles_m = (M + BM - 1) / BM;
    s.tiles_n = (N + BN - 1) / BN;
    s.nkb = (K + BK - 1) / BK;
    const int tiles = s.tiles_m * s.tiles_n;
    int nsplit = 1;
    if (tiles < 111 && s.nkb >= 16) {
        nsplit = (148 + tiles - 1) / tiles;
        const int cap = s.nkb / 8;  // at least 8 k-blocks (32 KB x 8) per work item
        if (nsplit > cap) nsplit = cap;
        if (nsplit < 1) nsplit = 1;
    }
    const char *force = getenv("SCVAE_TC_NSPLIT");
    if (force && atoi(force) > 0) nsplit = atoi(force);
    s.kb_per_split = (s.nkb + nsplit - 1) / nsplit;
    s.nsplit = (s.nkb + s.kb_per_split - 1) / s.kb_per_split;  // no empty splits
    return s;
}

}  // namespace scvae

using namespace scvae;

extern "C" int64_t scvae_gemm_tf32_workspace_bytes(int layout, int M, int N, int K) {
    (void)layout;
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    const SplitPlan s = plan_split(M, N, K);
    if (s.nsplit == 1) return 0;
    const int64_t ldw = (N + 3) & ~3;
    return (int64_t)s.nsplit * s.tiles_m * BM * ldw * 4;
}

extern "C" int scvae_gemm_tf32(int layout, int M, int N, int K, const float *A, int64_t lda, const float *B,
                               int64_t ldb, float *C, int64_t ldc, int accumulate, void *workspace,
                               int64_t workspace_bytes, void *stream) {
    SCVAE_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0, "gemm_tf32: bad arguments");
    SCVAE_CHECK_ARG(aligned16(A) && aligned16(B) && aligned16(C) && lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0,
                    "gemm_tf32: operands must be 16-byte aligned with leading dimensions multiple of 4");
    cudaStream_t s = (cudaStream_t)stream;
    SplitPlan sp = plan_split(M, N, K);
    const int64_t ldw = (N + 3) & ~3;
    const int ws_rows = sp.tiles_m * BM;
    if (sp.nsplit > 1) {
        const int64_t need = (int64_t)sp.nsplit * ws_rows * ldw * 4;
        if (!workspace || workspace_bytes < need) {  // no workspace: run unsplit
            sp.nsplit = 1;
            sp.kb_per_split = sp.nkb;
        }
    }
    CUtensorMap tmA, tmB, tmC;
    const bool a_mn = (layout == SCVAE_GEMM_TN), b_mn = (layout != SCVAE_GEMM_NT);
    // K-major operand (rows = M or N, cols = K): box 32 (K) x 128 (rows).
    // MN-major operand (rows = K, cols = M or N): box 32 (MN) x 32 (K rows), four per tile,
    // swizzled in 32-byte chunks (Swizzle<2,5,2>, atom = 128 B of MN x 4 K rows).
    if (a_mn) { if (make_map(&tmA, A, K, M, lda, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return 1; }
    else      { if (make_map(&tmA, A, M, K, lda, 32, 128)) return 1; }
    if (b_mn) { if (make_map(&tmB, B, K, N, ldb, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return 1; }
    else      { if (make_map(&tmB, B, N, K, ldb, 32, 128)) return 1; }
    if (sp.nsplit > 1) {
        if (make_map(&tmC, (const float *)workspace, (int64_t)sp.nsplit * ws_rows, N, ldw, 32, 128)) return 1;
    } else {
        if (make_map(&tmC, C, M, N, ldc, 32, 128)) return 1;
    }
    GemmParams p;
    p.M = M; p.N = N; p.K = K;
    p.tiles_m = sp.tiles_m; p.tiles_n = sp.tiles_n; p.nsplit = sp.nsplit;
    p.kb_per_split = sp.kb_per_split; p.nkb = sp.nkb;
    p.accumulate = accumulate; p.ws_rows = ws_rows;
    p.mn_lbo = 4096; p.mn_sbo = 512;  // MN-group (TMA box) stride; 4-row K-group stride
    if (const char *e = getenv("SCVAE_TC_MN_LBO")) p.mn_lbo = (uint32_t)atoi(e);
    if (const char *e = getenv("SCVAE_TC_MN_SBO")) p.mn_sbo = (uint32_t)atoi(e);

    const int total = sp.tiles_m * sp.tiles_n * sp.nsplit;
    int sms = 148;
    {
        static int cached = 0;
        if (!cached) {
            int dev = 0, n = 0;
            if (cudaGetDevice(&dev) == cudaSuccess &&
                cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
                cached = n;
            else
                cached = 148;
        }
        sms = cached;
    }
    const int grid = total < sms ? total : sms;
#define LAUNCH(AM, BMN)                                                                                     \
    do {                                                                                                    \
        static bool attr_set = false;                                                                       \
        if (!attr_set) {                                                                                    \
            cudaError_t e = cudaFuncSetAttribute(gemm_tf32_kernel<AM, BMN>,                                 \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);  \
            SCVAE_CHECK_ARG(e == cudaSuccess, "gemm_tf32: cannot set smem attribute: %s",                   \
                            cudaGetErrorString(e));                                                         \
            attr_set = true;                                                                                \
        }                                                                                                   \
        gemm_tf32_kernel<AM, BMN><<<grid, kThreads, kSmemBytes, s>>>(tmA, tmB, tmC, p);                     \
    } while (0)
    switch (layout) {
        case SCVAE_GEMM_NT: LAUNCH(false, false); break;
        case SCVAE_GEMM_NN: LAUNCH(false, true); break;
        case SCVAE_GEMM_TN: LAUNCH(true, true); break;
        default: set_error("gemm_tf32: unknown layout %d", layout); return 1;
    }
#undef LAUNCH
    SCVAE_CHECK_LAUNCH("gemm_tf32");
    if (sp.nsplit > 1) {
        const dim3 grid2((N + 127) / 128, M);
        SCVAE_CHECK_ARG(M <= 65535, "gemm_tf32: split-K reduce supports M <= 65535");
        splitk_reduce_kernel<<<grid2, 128, 0, s>>>((const float *)workspace, ldw, ws_rows, sp.nsplit, M, N, C, ldc,
                                                   accumulate);
        SCVAE_CHECK_LAUNCH("splitk_reduce");
    }
    return 0;
}
