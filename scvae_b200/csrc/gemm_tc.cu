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
#include "tc_common.cuh"

namespace scvae {

constexpr int BM = 128, BN = 128, BK = 32;       // BK fp32 = 128 bytes = one swizzle row
constexpr int kStages = 5;
constexpr int kTileBytes = BM * BK * 4;          // 16 KB per operand per stage
constexpr int kStageBytes = 2 * kTileBytes;
constexpr int kEpiBytes = BM * 32 * 4;           // 128 rows x 32 fp32 columns
// dual mode (a second, low-order copy of one operand accumulated into the same tile): three tiles
// per stage, four stages
constexpr int kDualStages = 4;
constexpr int kDualStageBytes = 3 * kTileBytes;
constexpr int kSmemBytes = kStages * kStageBytes + 2 * kEpiBytes + 1024 /*align*/ + 256 /*barriers*/;
constexpr int kSmemBytesDual = kDualStages * kDualStageBytes + 2 * kEpiBytes + 1024 + 256;
// pair mode (dual forward only): one work item covers TWO m-tiles that share the (high, low) weight
// tiles -- four tiles per stage for two output tiles instead of six, three stages
constexpr int kPairStages = 3;
constexpr int kPairStageBytes = 4 * kTileBytes;
constexpr int kSmemBytesPair = kPairStages * kPairStageBytes + 2 * kEpiBytes + 1024 + 256;
constexpr int kTmemCols = 2 * BN;                // double-buffered fp32 accumulator
constexpr int kThreads = 256;

struct GemmParams {
    int M, N, K;
    int tiles_m, tiles_n, nsplit, kb_per_split, nkb;
    int accumulate;     // TMA reduce-add into C (only when nsplit == 1)
    int ws_rows;        // rows per split slice of the workspace (multiple of BM)
    uint32_t mn_lbo, mn_sbo;  // descriptor strides of MN-major operand tiles
    float alpha;              // accumulator scale applied in the epilogue
    int streamk;              // 1: CTAs own equal contiguous ranges of (tile, k-block) units
    int units_per_cta, total_units;
    int dual;                 // 0: C = A B;  1: C = (A + X) B;  2: C = A (B + X)   (X: the low-order half
                              // of an fp32 operand split into two fp16 matrices, same shape and major)
    int m_mult;               // m-tiles per work item (2 in pair mode: tiles_m counts pairs)
    int stream_hint;          // bit 0 / 1: operand A / B streams through once -> L2 evict-first loads
};

// One unit of work of a CTA: k-blocks [kb0, kb1) of output tile (tm, tn).
struct Segment {
    int tm, tn, kb0, kb1, out_row;
    bool partial;             // tile shared with another CTA: combine with TMA reduce-add
};

// Work sequence of this CTA; the producer, MMA and epilogue roles all walk the same sequence.
//   tile/split mode: items blockIdx.x, blockIdx.x + gridDim.x, ... of the (m-tile, n-tile, k-split)
//     list; a split writes its partial tile into its own workspace slice (reduced afterwards in
//     fixed order);
//   stream-K mode (tile count just above the SM count): the CTA owns units_per_cta consecutive
//     (tile, k-block) units; units_per_cta >= k-blocks per tile, so a tile is shared by at most
//     two CTAs and the reduce-add of two partials into a zeroed C is order independent.
struct WorkIter {
    int it = 0, u = -1;
    __device__ __forceinline__ bool next(const GemmParams &p, Segment &s) {
        if (p.streamk) {
            const int u1 = min(((int)blockIdx.x + 1) * p.units_per_cta, p.total_units);
            if (u < 0) u = (int)blockIdx.x * p.units_per_cta;
            if (u >= u1) return false;
            const int tile = u / p.nkb;
            s.kb0 = u % p.nkb;
            s.kb1 = min(p.nkb, s.kb0 + (u1 - u));
            u += s.kb1 - s.kb0;
            s.tm = tile % p.tiles_m;
            s.tn = tile / p.tiles_m;
            s.out_row = s.tm * BM * p.m_mult;
            s.partial = !(s.kb0 == 0 && s.kb1 == p.nkb);
            return true;
        }
        const int item = (int)blockIdx.x + it * (int)gridDim.x;
        if (item >= p.tiles_m * p.tiles_n * p.nsplit) return false;
        ++it;
        s.tm = item % p.tiles_m;
        s.tn = (item / p.tiles_m) % p.tiles_n;
        const int sp = item / (p.tiles_m * p.tiles_n);
        s.kb0 = sp * p.kb_per_split;
        s.kb1 = min(s.kb0 + p.kb_per_split, p.nkb);
        s.out_row = (p.nsplit > 1 ? sp * p.ws_rows : 0) + s.tm * BM * p.m_mult;
        s.partial = false;
        return true;
    }
};

// F16: operands are fp16 (kind::f16, 64-element k-blocks, plain 128B swizzle for both majors);
// otherwise fp32 consumed as tf32 (32-element k-blocks, 32-byte-atom swizzle for MN-major).
// DUAL (compile time, so that no tcgen05.mma sits under a run-time predicate): 0: C = A B;
// 1: C = (A + X) B;  2: C = A (B + X).
template <bool F16, bool A_MN, bool B_MN, int DUAL = 0, bool PAIR = false>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmX,
                 const GemmParams p) {
    pdl_wait();          // (programmatic dependent launch: common.cuh)
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    static_assert(!PAIR || (DUAL == 2 && !A_MN), "pair mode: forward with split weights only");
    constexpr int nstages = PAIR ? kPairStages : (DUAL ? kDualStages : kStages);
    constexpr int stage_bytes = PAIR ? kPairStageBytes : (DUAL ? kDualStageBytes : kStageBytes);
    constexpr int ACC_COLS = PAIR ? 2 * BN : BN;     // TMEM columns of one work item's accumulators
    constexpr int TMEM_COLS = 2 * ACC_COLS;          // double buffered
    uint8_t *epi = smem + nstages * stage_bytes;
    uint64_t *bars = (uint64_t *)(epi + 2 * kEpiBytes);
    // bars: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], then tmem base address
    const uint32_t bar_full = smem_u32(bars);
    const uint32_t bar_empty = smem_u32(bars + kStages);
    const uint32_t bar_tfull = smem_u32(bars + 2 * kStages);
    const uint32_t bar_tempty = smem_u32(bars + 2 * kStages + 2);
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int EB = F16 ? 2 : 4;             // operand element bytes
    constexpr int BKE = 128 / EB;               // elements per k-block (one 128-byte swizzle row)
    constexpr int UK = 32 / EB;                 // elements per tcgen05.mma along K
    constexpr int MN_BOX = 128 / EB;            // MN elements per TMA box of an MN-major tile
    constexpr int MN_BOXES = BM / MN_BOX;       // boxes per 128-wide tile
    constexpr int MN_BOX_BYTES = BKE * 128;     // BKE k-rows of 128 bytes
    constexpr uint32_t MN_LAYOUT = F16 ? 2u : 1u;   // SWIZZLE_128B vs SWIZZLE_128B_BASE32B
    constexpr uint32_t MN_KSTEP = UK * 128;     // bytes per UMMA K step in an MN-major tile

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
        if (DUAL) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
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
                     "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            WorkIter work;
            Segment sg;
            const bool hint_a = (p.stream_hint & 1) != 0, hint_b = (p.stream_hint & 2) != 0;
            const uint64_t pol = l2_policy_evict_first();
            while (work.next(p, sg)) {
                const int tm = PAIR ? 2 * sg.tm : sg.tm, tn = sg.tn, kb0 = sg.kb0, kb1 = sg.kb1;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t sa = smem_u32(smem + stage * stage_bytes);
                    const uint32_t sb = sa + kTileBytes;
                    const uint32_t sx = sb + kTileBytes;
                    const uint32_t full = bar_full + 8 * stage;
                    mbar_expect_tx(full, stage_bytes);
                    // pair mode: the second m-tile (rows beyond M arrive as zeros)
                    if (PAIR) {
                        if (hint_a) tma_load_2d_hint(sx + kTileBytes, &tmA, kb * BKE, (tm + 1) * BM, full, pol);
                        else tma_load_2d(sx + kTileBytes, &tmA, kb * BKE, (tm + 1) * BM, full);
                    }
                    if (DUAL == 1) {
                        if (A_MN) {
#pragma unroll
                            for (int j = 0; j < MN_BOXES; ++j)
                                tma_load_2d(sx + j * MN_BOX_BYTES, &tmX, tm * BM + MN_BOX * j, kb * BKE, full);
                        } else {
                            tma_load_2d(sx, &tmX, kb * BKE, tm * BM, full);
                        }
                    } else if (DUAL == 2) {
                        if (B_MN) {
#pragma unroll
                            for (int j = 0; j < MN_BOXES; ++j)
                                tma_load_2d(sx + j * MN_BOX_BYTES, &tmX, tn * BN + MN_BOX * j, kb * BKE, full);
                        } else {
                            tma_load_2d(sx, &tmX, kb * BKE, tn * BN, full);
                        }
                    }
                    if (A_MN) {
#pragma unroll
                        for (int j = 0; j < MN_BOXES; ++j) {
                            if (hint_a) tma_load_2d_hint(sa + j * MN_BOX_BYTES, &tmA, tm * BM + MN_BOX * j, kb * BKE, full, pol);
                            else tma_load_2d(sa + j * MN_BOX_BYTES, &tmA, tm * BM + MN_BOX * j, kb * BKE, full);
                        }
                    } else {
                        if (hint_a) tma_load_2d_hint(sa, &tmA, kb * BKE, tm * BM, full, pol);
                        else tma_load_2d(sa, &tmA, kb * BKE, tm * BM, full);
                    }
                    if (B_MN) {
#pragma unroll
                        for (int j = 0; j < MN_BOXES; ++j) {
                            if (hint_b) tma_load_2d_hint(sb + j * MN_BOX_BYTES, &tmB, tn * BN + MN_BOX * j, kb * BKE, full, pol);
                            else tma_load_2d(sb + j * MN_BOX_BYTES, &tmB, tn * BN + MN_BOX * j, kb * BKE, full);
                        }
                    } else {
                        if (hint_b) tma_load_2d_hint(sb, &tmB, kb * BKE, tn * BN, full, pol);
                        else tma_load_2d(sb, &tmB, kb * BKE, tn * BN, full);
                    }
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32|f16, M=128, N=128
            constexpr uint32_t FMT = F16 ? 0u : 2u;
            const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((A_MN ? 1u : 0u) << 15) |
                                   ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)(BM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int t = 0;
            WorkIter work;
            Segment sg;
            for (; work.next(p, sg); ++t) {
                const int kb0 = sg.kb0, kb1 = sg.kb1;
                const int acc = t & 1;
                const uint32_t acc_phase = (t >> 1) & 1;
                mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * ACC_COLS;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * stage_bytes);
                    const uint32_t sb = sa + kTileBytes;
                    const uint32_t sx = sb + kTileBytes;
                    const uint64_t da = A_MN ? make_desc(sa, p.mn_lbo, p.mn_sbo, MN_LAYOUT) : make_desc(sa, 16, 1024, 2);
                    const uint64_t db = B_MN ? make_desc(sb, p.mn_lbo, p.mn_sbo, MN_LAYOUT) : make_desc(sb, 16, 1024, 2);
                    // per UMMA K step: K-major +32 B inside the swizzle row; MN-major +UK k-rows
                    const uint64_t sta = A_MN ? (MN_KSTEP >> 4) : (32 >> 4);
                    const uint64_t stb = B_MN ? (MN_KSTEP >> 4) : (32 >> 4);
#pragma unroll
                    for (int k = 0; k < BKE / UK; ++k) {
                        if (F16)
                            tc_mma_f16(tmem_d, da + k * sta, db + k * stb, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        else
                            tc_mma_tf32(tmem_d, da + k * sta, db + k * stb, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    if constexpr (F16 && DUAL != 0) {
                        // low-order half of the split operand into the same accumulator
                        const uint64_t dx = (DUAL == 1 ? A_MN : B_MN) ? make_desc(sx, p.mn_lbo, p.mn_sbo, MN_LAYOUT)
                                                                      : make_desc(sx, 16, 1024, 2);
                        const uint64_t a2 = DUAL == 1 ? dx : da, b2 = DUAL == 1 ? db : dx;
#pragma unroll
                        for (int k = 0; k < BKE / UK; ++k) tc_mma_f16(tmem_d, a2 + k * sta, b2 + k * stb, idesc, 1u);
                        if constexpr (PAIR) {
                            // second m-tile against the same (high, low) weight tiles
                            const uint64_t da1 = make_desc(sx + kTileBytes, 16, 1024, 2);
#pragma unroll
                            for (int k = 0; k < BKE / UK; ++k)
                                tc_mma_f16(tmem_d + BN, da1 + k * sta, db + k * stb, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
#pragma unroll
                            for (int k = 0; k < BKE / UK; ++k) tc_mma_f16(tmem_d + BN, da1 + k * sta, dx + k * stb, idesc, 1u);
                        }
                    }
                    tc_commit(bar_empty + 8 * stage);  // frees the smem slot once the MMAs retire
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
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
        WorkIter work;
        Segment sg;
        for (; work.next(p, sg); ++t) {
            const int tn = sg.tn;
            const int acc = t & 1;
            const uint32_t acc_phase = (t >> 1) & 1;
            mbar_wait(bar_tfull + 8 * acc, acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int half = 0; half < (PAIR ? 2 : 1); ++half) {
            const int out_row = sg.out_row + half * BM;
            if (half && (2 * sg.tm + 1) * BM >= p.M) break;      // (uniform) no second m-tile
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                if (tn * BN + c * 32 >= p.N) break;   // uniform: nothing to store
                uint32_t v[32];
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS + half * BN + c * 32, v);
                tc_wait_ld();
                if (issuer) tma_wait_read<1>();       // staging buffer `buf` is free again
                epi_bar_sync();
                uint8_t *dst = epi + buf * kEpiBytes + row * 128;
                if (p.alpha != 1.f) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * p.alpha);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint4 val = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    *reinterpret_cast<uint4 *>(dst + ((j ^ (row & 7)) << 4)) = val;
                }
                fence_async_smem();
                epi_bar_sync();
                if (issuer) {
                    const uint32_t src = smem_u32(epi + buf * kEpiBytes);
                    if ((p.accumulate && p.nsplit == 1) || sg.partial)
                        tma_reduce_add_2d(&tmC, tn * BN + c * 32, out_row, src);
                    else
                        tma_store_2d(&tmC, tn * BN + c * 32, out_row, src);
                    tma_commit();
                }
                buf ^= 1;
            }
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
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS)
                     : "memory");
    }
}

// C[m, n] (+)= alpha * sum_s ws[s][m][n], fixed order.  One thread per 4 columns (ldw, ldc % 4 == 0).
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float *__restrict__ ws, int64_t ldw, int ws_rows, int nsplit, int M, int N,
                     float *__restrict__ C, int64_t ldc, int accumulate, float alpha) {
    pdl_wait();          // (programmatic dependent launch: common.cuh)
    pdl_trigger();
    const int n4 = (N + 3) >> 2;
    const int64_t total = (int64_t)M * n4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(i / n4), n = (int)(i % n4) << 2;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < nsplit; ++s) {
            const float4 v = *reinterpret_cast<const float4 *>(ws + ((int64_t)s * ws_rows + m) * ldw + n);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        float *c = C + (int64_t)m * ldc + n;
        const float a[4] = {acc.x * alpha, acc.y * alpha, acc.z * alpha, acc.w * alpha};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (n + j < N) c[j] = accumulate ? c[j] + a[j] : a[j];
    }
}

struct SplitPlan {
    int tiles_m, tiles_n, nkb, nsplit, kb_per_split;
    int streamk, units_per_cta, grid;
};

// Upper bound on the persistent CTAs of subsequent tensor-core GEMM launches (0 = all SMs): lets
// the caller leave SMs free for kernels it runs concurrently on another stream.
static thread_local int g_sm_limit = 0;

static int device_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            sms = n;
        else
            sms = 148;
    }
    return sms;
}
static int num_sms() {
    const int n = device_sms();
    return (g_sm_limit > 0 && g_sm_limit < n) ? g_sm_limit : n;
}

// Wave efficiency of `items` equal work items on `sms` persistent CTAs.
static double wave_efficiency(int items, int sms) {
    const int rounds = (items + sms - 1) / sms;
    return (double)items / ((double)rounds * sms);
}

static SplitPlan plan_split(int M, int N, int K, int bke = BK, bool allow_streamk = false, bool pair_m = false) {
    SplitPlan s;
    const int sms = num_sms();
    s.tiles_m = (M + BM - 1) / BM;
    if (pair_m) s.tiles_m = (s.tiles_m + 1) / 2;      // work items cover two m-tiles
    s.tiles_n = (N + BN - 1) / BN;
    s.nkb = (K + bke - 1) / bke;
    s.streamk = 0;
    s.units_per_cta = 0;
    const int tiles = s.tiles_m * s.tiles_n;
    int nsplit = 1;
    if (tiles < sms && s.nkb >= 16) {
        // split K so that the work items fill whole waves of SMs; among near-best choices take
        // the smallest split (least workspace traffic); >= 8 k-blocks per item
        const int cap = s.nkb / 8 < 64 ? s.nkb / 8 : 64;
        double best = wave_efficiency(tiles, sms);
        for (int ns = 2; ns <= cap; ++ns) {
            const int kbp = (s.nkb + ns - 1) / ns;
            const int eff_ns = (s.nkb + kbp - 1) / kbp;
            const double e = wave_efficiency(tiles * eff_ns, sms);
            if (e > best + 0.03) {
                best = e;
                nsplit = ns;
            }
        }
    } else if (allow_streamk && tiles > sms && tiles < 4 * sms && s.nkb >= 8 &&
               wave_efficiency(tiles, sms) < 0.85) {
        s.streamk = 1;
        s.units_per_cta = (int)(((int64_t)tiles * s.nkb + sms - 1) / sms);   // >= nkb since tiles > sms
    }
    const char *force = getenv("SCVAE_TC_NSPLIT");
    if (force && atoi(force) > 0) { nsplit = atoi(force); s.streamk = 0; }
    s.kb_per_split = (s.nkb + nsplit - 1) / nsplit;
    s.nsplit = (s.nkb + s.kb_per_split - 1) / s.kb_per_split;  // no empty splits
    const int total = tiles * s.nsplit;
    s.grid = s.streamk ? sms : (total < sms ? total : sms);
    return s;
}

// Pair mode (two m-tiles per work item) of the dual forward product: off with SCVAE_TC_PAIR=0.
static bool pair_enabled(int M) {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("SCVAE_TC_PAIR");
        on = (e && atoi(e) == 0) ? 0 : 1;
    }
    return on && M > BM;
}

static int64_t workspace_bytes(int M, int N, int K, int bke) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    const int64_t ldw = (N + 3) & ~3;
    const SplitPlan s = plan_split(M, N, K, bke);
    int64_t need = s.nsplit == 1 ? 0 : (int64_t)s.nsplit * s.tiles_m * BM * ldw * 4;
    if (bke == 64 && pair_enabled(M)) {       // the same product may run in pair mode (split operand)
        const SplitPlan q = plan_split(M, N, K, bke, false, true);
        const int64_t alt = q.nsplit == 1 ? 0 : (int64_t)q.nsplit * q.tiles_m * 2 * BM * ldw * 4;
        if (alt > need) need = alt;
    }
    return need;
}

template <bool F16>
static int launch_gemm(const char *name, int layout, int M, int N, int K, const void *A, int64_t lda, const void *B,
                       int64_t ldb, float *C, int64_t ldc, int accumulate, float alpha, void *workspace,
                       int64_t workspace_bytes_given, cudaStream_t s, const void *X = nullptr, int64_t ldx = 0,
                       int dual = 0) {
    constexpr int EB = F16 ? 2 : 4;
    constexpr int BKE = 128 / EB;
    constexpr int LDM = 16 / EB;   // leading dimensions must be multiples of 16 bytes
    SCVAE_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0, "%s: bad arguments", name);
    SCVAE_CHECK_ARG(aligned16(A) && aligned16(B) && aligned16(C) && lda % LDM == 0 && ldb % LDM == 0 && ldc % 4 == 0,
                    "%s: operands must be 16-byte aligned with 16-byte-multiple leading dimensions", name);
    const bool pair = F16 && dual == 2 && layout == SCVAE_GEMM_NT && pair_enabled(M);
    SplitPlan sp = plan_split(M, N, K, BKE, /*allow_streamk=*/!accumulate && !pair, pair);
    const int64_t ldw = (N + 3) & ~3;
    const int ws_rows = sp.tiles_m * BM * (pair ? 2 : 1);
    if (sp.nsplit > 1) {
        const int64_t need = (int64_t)sp.nsplit * ws_rows * ldw * 4;
        if (!workspace || workspace_bytes_given < need) {  // no workspace: run unsplit
            sp.nsplit = 1;
            sp.kb_per_split = sp.nkb;
        }
    }
    CUtensorMap tmA, tmB, tmC, tmX;
    const bool a_mn = (layout == SCVAE_GEMM_TN), b_mn = (layout != SCVAE_GEMM_NT);
    SCVAE_CHECK_ARG(dual == 0 || (F16 && X && (dual == 1 || dual == 2) && aligned16(X) && ldx % LDM == 0),
                    "%s: bad split operand", name);
    // K-major operand (rows = M or N, cols = K): box BKE (K) x 128 (rows), SWIZZLE_128B.
    // MN-major operand (rows = K, cols = M or N): boxes of 128 bytes (MN) x BKE (K rows); fp32
    // needs the 32-byte-atom swizzle (Swizzle<2,5,2>, atom = 128 B of MN x 4 K rows).
    const CUtensorMapSwizzle mn_sw = F16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    if (a_mn) { if (make_map(&tmA, A, K, M, lda, 128 / EB, BKE, mn_sw, EB)) return 1; }
    else      { if (make_map(&tmA, A, M, K, lda, BKE, 128, CU_TENSOR_MAP_SWIZZLE_128B, EB)) return 1; }
    if (b_mn) { if (make_map(&tmB, B, K, N, ldb, 128 / EB, BKE, mn_sw, EB)) return 1; }
    else      { if (make_map(&tmB, B, N, K, ldb, BKE, 128, CU_TENSOR_MAP_SWIZZLE_128B, EB)) return 1; }
    if (dual == 1) {
        if (a_mn) { if (make_map(&tmX, X, K, M, ldx, 128 / EB, BKE, mn_sw, EB)) return 1; }
        else      { if (make_map(&tmX, X, M, K, ldx, BKE, 128, CU_TENSOR_MAP_SWIZZLE_128B, EB)) return 1; }
    } else if (dual == 2) {
        if (b_mn) { if (make_map(&tmX, X, K, N, ldx, 128 / EB, BKE, mn_sw, EB)) return 1; }
        else      { if (make_map(&tmX, X, N, K, ldx, BKE, 128, CU_TENSOR_MAP_SWIZZLE_128B, EB)) return 1; }
    } else {
        tmX = tmA;
    }
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
    p.alpha = sp.nsplit > 1 ? 1.f : alpha;
    // MN-group (TMA box) stride; K-group stride (4 rows for the 32-byte-atom swizzle, else 8)
    p.mn_lbo = BKE * 128; p.mn_sbo = F16 ? 1024 : 512;
    if (const char *e = getenv("SCVAE_TC_MN_LBO")) p.mn_lbo = (uint32_t)atoi(e);
    if (const char *e = getenv("SCVAE_TC_MN_SBO")) p.mn_sbo = (uint32_t)atoi(e);

    p.dual = dual;
    p.m_mult = pair ? 2 : 1;
    // which operand streams through once: the (cells x genes)-sized one.  NT: A (the minibatch);
    // TN: A (the head pre-activation gradient).  Only for operands far larger than the other one.
    // (TN with the output gradient split, B = the minibatch: marking it costs 6 us per step -- it is
    // the last reader of X16 in a step and the next densify rewrites the buffer -- so it is off.)
    p.stream_hint = 0;
    {
        static int enabled = -1;       // bit 0: NT forward, bit 1: TN with split output gradient, bit 2: TN
        if (enabled < 0) {
            const char *e = getenv("SCVAE_TC_L2_HINT");
            enabled = e ? atoi(e) : 5;      // (measured at C2: none 0.575, 4: 0.565, 5: 0.560, 7: 0.566 ms per step)
        }
        if (enabled && F16) {
            if ((enabled & 1) && layout == SCVAE_GEMM_NT && (int64_t)M >= 8 * (int64_t)N) p.stream_hint = 1;
            if ((enabled & 2) && layout == SCVAE_GEMM_TN && dual == 1 && (int64_t)N >= 8 * (int64_t)M) p.stream_hint = 2;
            if ((enabled & 4) && layout == SCVAE_GEMM_TN && dual == 0 && (int64_t)M >= 8 * (int64_t)N) p.stream_hint = 1;
        }
    }
    p.streamk = sp.streamk;
    p.units_per_cta = sp.units_per_cta;
    p.total_units = sp.tiles_m * sp.tiles_n * sp.nkb;
    if (sp.nsplit > 1) {   // (the no-workspace fallback above may have changed the item count)
        sp.grid = sp.tiles_m * sp.tiles_n * sp.nsplit < num_sms() ? sp.tiles_m * sp.tiles_n * sp.nsplit : num_sms();
    } else if (!sp.streamk) {
        sp.grid = sp.tiles_m * sp.tiles_n < num_sms() ? sp.tiles_m * sp.tiles_n : num_sms();
    }
    const int grid = sp.grid;
    if (sp.streamk) {   // partial tiles are reduce-added into C
        // (the N columns of the product only: the caller may keep other columns of a wider C)
        cudaError_t e = (int64_t)N == ldc ? cudaMemsetAsync(C, 0, (size_t)M * ldc * sizeof(float), s)
                                          : cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0,
                                                              (size_t)N * sizeof(float), (size_t)M, s);
        SCVAE_CHECK_ARG(e == cudaSuccess, "%s: memset failed: %s", name, cudaGetErrorString(e));
    }
#define LAUNCH(AM, BMN, DU, PR)                                                                             \
    do {                                                                                                    \
        constexpr int smem_bytes = (PR) ? kSmemBytesPair : ((DU) ? kSmemBytesDual : kSmemBytes);            \
        static bool attr_set = false;                                                                       \
        if (!attr_set) {                                                                                    \
            cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<F16, AM, BMN, DU, PR>,                      \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);  \
            SCVAE_CHECK_ARG(e == cudaSuccess, "%s: cannot set smem attribute: %s", name,                    \
                            cudaGetErrorString(e));                                                         \
            attr_set = true;                                                                                \
        }                                                                                                   \
        /* (behind the stream-K memset the predecessor is not a kernel: plain launch) */                   \
        launch_pdl(sp.streamk ? 0 : kPdlGemm, gemm_tc_kernel<F16, AM, BMN, DU, PR>, dim3(grid), dim3(kThreads), \
                   smem_bytes, s, tmA, tmB, tmC, tmX, p);                                                   \
    } while (0)
    if (dual == 0) {
        switch (layout) {
            case SCVAE_GEMM_NT: LAUNCH(false, false, 0, false); break;
            case SCVAE_GEMM_NN: LAUNCH(false, true, 0, false); break;
            case SCVAE_GEMM_TN: LAUNCH(true, true, 0, false); break;
            default: set_error("%s: unknown layout %d", name, layout); return 1;
        }
    } else if (F16 && dual == 2 && layout == SCVAE_GEMM_NT) {
        if (pair) LAUNCH(false, false, 2, true);     // forward with split weights, two m-tiles per item
        else LAUNCH(false, false, 2, false);
    } else if (F16 && dual == 1 && layout == SCVAE_GEMM_TN) {
        LAUNCH(true, true, 1, false);     // weight gradient with split output gradient
    } else {
        set_error("%s: split operands are built for NT (which = 2) and TN (which = 1)", name);
        return 1;
    }
#undef LAUNCH
    SCVAE_CHECK_LAUNCH(name);
    if (sp.nsplit > 1) {
        int64_t blocks = ((int64_t)M * ((N + 3) >> 2) + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        launch_pdl(kPdlReduce, splitk_reduce_kernel, dim3((unsigned)blocks), dim3(256), 0, s, (const float *)workspace, ldw,
                   ws_rows, sp.nsplit, M, N, C, ldc, accumulate, alpha);
        SCVAE_CHECK_LAUNCH("splitk_reduce");
    }
    return 0;
}

}  // namespace scvae

using namespace scvae;

extern "C" int scvae_gemm_sm_limit(int max_ctas) {
    const int prev = g_sm_limit;
    g_sm_limit = max_ctas > 0 ? max_ctas : 0;
    return prev;
}

extern "C" int scvae_gemm_f16_split(int layout, int M, int N, int K, const void *A, int64_t lda, const void *B,
                                    int64_t ldb, const void *X, int64_t ldx, int which, float *C, int64_t ldc,
                                    int accumulate, float alpha, void *workspace, int64_t workspace_bytes,
                                    void *stream) {
    return launch_gemm<true>("gemm_f16_split", layout, M, N, K, A, lda, B, ldb, C, ldc, accumulate, alpha, workspace,
                             workspace_bytes, (cudaStream_t)stream, X, ldx, which);
}

extern "C" int64_t scvae_gemm_tf32_workspace_bytes(int layout, int M, int N, int K) {
    (void)layout;
    return workspace_bytes(M, N, K, 32);
}
extern "C" int64_t scvae_gemm_f16_workspace_bytes(int layout, int M, int N, int K) {
    (void)layout;
    return workspace_bytes(M, N, K, 64);
}

extern "C" int scvae_gemm_tf32(int layout, int M, int N, int K, const float *A, int64_t lda, const float *B,
                               int64_t ldb, float *C, int64_t ldc, int accumulate, void *workspace,
                               int64_t workspace_bytes, void *stream) {
    return launch_gemm<false>("gemm_tf32", layout, M, N, K, A, lda, B, ldb, C, ldc, accumulate, 1.f, workspace,
                              workspace_bytes, (cudaStream_t)stream);
}

extern "C" int scvae_gemm_f16(int layout, int M, int N, int K, const void *A, int64_t lda, const void *B,
                              int64_t ldb, float *C, int64_t ldc, int accumulate, float alpha, void *workspace,
                              int64_t workspace_bytes, void *stream) {
    return launch_gemm<true>("gemm_f16", layout, M, N, K, A, lda, B, ldb, C, ldc, accumulate, alpha, workspace,
                             workspace_bytes, (cudaStream_t)stream);
}
