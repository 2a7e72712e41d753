// a1: minibatch gather.  Replaces the host-side x_train[idx].toarray() (VAE:994-998,
// GMVAE:1078-1082): the CSR count matrix lives in HBM and each step densifies B rows into the
// augmented (B, ldx) fp32 layout the GEMMs and the likelihood kernels read.  HBM-write-bound.
#include <cuda_fp16.h>

#include "common.cuh"

namespace scvae {

constexpr int kChunk = 8192;   // genes staged in shared memory per pass (32 KB)

// One CTA per output row.  The row is assembled chunk by chunk in shared memory (zero, scatter
// the row's non-zeros that fall into the chunk, write out), so every output byte is written
// exactly once with coalesced 128-bit stores -- to up to three destinations:
//   x   fp32 augmented (column G = 1)          -> exact-fp32 / evaluation paths
//   x16 fp16 augmented (column G = 1)          -> fp16 tensor-core first layer
//   t16 uint16 (clamped to 65535), zero padded -> targets of the fused likelihood heads
template <typename IdxT, typename ValT>
__global__ void __launch_bounds__(256)
csr_densify_kernel(const int64_t *__restrict__ indptr, const IdxT *__restrict__ indices,
                   const ValT *__restrict__ values, const int64_t *__restrict__ rows, int G,
                   float *__restrict__ x, int64_t ldx, float *__restrict__ row_const, int rebase,
                   uint16_t *__restrict__ t16, int64_t ldt16, __half *__restrict__ x16, int64_t ldx16) {
    __shared__ float red[32];
    __shared__ __align__(16) float buf[kChunk];
    const int b = blockIdx.x;
    const int64_t row = rows ? rows[b] : b;
    const int64_t base = rebase ? indptr[0] : 0;
    const int64_t s = indptr[row] - base, e = indptr[row + 1] - base;
    int64_t width = 0;
    if (x) width = ldx;
    if (t16 && ldt16 > width) width = ldt16;
    if (x16 && ldx16 > width) width = ldx16;
    float acc = 0.f;
    for (int c0 = 0; c0 < width; c0 += kChunk) {
        for (int i = threadIdx.x * 4; i < kChunk; i += blockDim.x * 4)
            *reinterpret_cast<float4 *>(buf + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        for (int64_t i = s + threadIdx.x; i < e; i += blockDim.x) {
            const int c = (int)indices[i];
            if (c >= c0 && c < c0 + kChunk && c < G) {
                const float v = (float)values[i];
                buf[c - c0] = v;
                if (row_const && v > 0.f) acc += lgammaf(1.f + v);
            }
        }
        if (G >= c0 && G < c0 + kChunk && threadIdx.x == 0) buf[G - c0] = 1.f;   // augmented ones column
        __syncthreads();
        // write-out, 8 columns per thread per trip
        for (int i = threadIdx.x * 8; i < kChunk; i += blockDim.x * 8) {
            const int c = c0 + i;
            if (c >= width) break;
            const float4 lo = *reinterpret_cast<const float4 *>(buf + i);
            const float4 hi = *reinterpret_cast<const float4 *>(buf + i + 4);
            const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            if (x) {
                float *xr = x + (int64_t)b * ldx + c;
                if (c + 8 <= ldx && (ldx & 3) == 0) {
                    *reinterpret_cast<float4 *>(xr) = lo;
                    *reinterpret_cast<float4 *>(xr + 4) = hi;
                } else {
                    for (int j = 0; j < 8 && c + j < ldx; ++j) xr[j] = v[j];
                }
            }
            if (x16) {
                __half *hr = x16 + (int64_t)b * ldx16 + c;
                if (c + 8 <= ldx16) {
                    uint4 pk;
                    __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
                    __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
                    pk.x = *reinterpret_cast<uint32_t *>(&h0);
                    pk.y = *reinterpret_cast<uint32_t *>(&h1);
                    pk.z = *reinterpret_cast<uint32_t *>(&h2);
                    pk.w = *reinterpret_cast<uint32_t *>(&h3);
                    *reinterpret_cast<uint4 *>(hr) = pk;
                } else {
                    for (int j = 0; j < 8 && c + j < ldx16; ++j) hr[j] = __float2half_rn(v[j]);
                }
            }
            if (t16) {
                uint16_t *tr = t16 + (int64_t)b * ldt16 + c;
                uint16_t u[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    u[j] = (c + j < G) ? (uint16_t)fminf(fmaxf(v[j], 0.f), 65535.f) : (uint16_t)0;
                if (c + 8 <= ldt16) {
                    uint4 pk;
                    pk.x = u[0] | ((uint32_t)u[1] << 16);
                    pk.y = u[2] | ((uint32_t)u[3] << 16);
                    pk.z = u[4] | ((uint32_t)u[5] << 16);
                    pk.w = u[6] | ((uint32_t)u[7] << 16);
                    *reinterpret_cast<uint4 *>(tr) = pk;
                } else {
                    for (int j = 0; j < 8 && c + j < ldt16; ++j) tr[j] = u[j];
                }
            }
        }
        __syncthreads();
    }
    if (row_const) {
        const float tot = block_sum(acc, red);
        if (threadIdx.x == 0) row_const[b] = tot;
    }
}

// 16-bit outputs only (the fused training path): the whole row is assembled in shared memory as
// uint16 counts in ONE pass (zero, scatter, write), then written as fp16 (x16, augmented) and/or
// uint16 (t16).  Dynamic shared memory: 2 * row_width bytes.
// X16_DIRECT (x16 only, the common case: counts <= 2048 serve as fp16 input AND targets): the row
// is assembled as fp16 bit patterns, so the write-out is a plain 128-bit copy.
template <typename IdxT, typename ValT, bool X16_DIRECT>
__global__ void __launch_bounds__(256)
csr_densify16_kernel(const int64_t *__restrict__ indptr, const IdxT *__restrict__ indices,
                     const ValT *__restrict__ values, const int64_t *__restrict__ rows, int G,
                     float *__restrict__ row_const, int rebase, uint16_t *__restrict__ t16, int64_t ldt16,
                     __half *__restrict__ x16, int64_t ldx16, int width8, int part8) {
    // blockIdx.y selects a column part of the row (part8 groups of 8 columns each): more CTAs in
    // flight hide the dependent indptr -> indices -> values load chain
    pdl_trigger();       // (the first-layer product behind this kernel may be scheduled early: common.cuh)
    extern __shared__ __align__(16) uint16_t row16[];
    __shared__ float red[32];
    const int b = blockIdx.x;
    const int g8_lo = blockIdx.y * part8;
    const int g8_hi = min(g8_lo + part8, width8);
    const int c_lo = g8_lo << 3, c_hi = g8_hi << 3;
    const int64_t row = rows ? rows[b] : b;
    const int64_t base = rebase ? indptr[0] : 0;
    const int64_t s = indptr[row] - base, e = indptr[row + 1] - base;
    uint4 *row4 = reinterpret_cast<uint4 *>(row16);
    for (int i = threadIdx.x; i < g8_hi - g8_lo; i += blockDim.x) row4[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    float acc = 0.f;
    // four (index, value) pairs per thread per trip are loaded before any is consumed, so the
    // row's non-zeros stream in with 8 independent loads in flight per thread
    for (int64_t i0 = s + threadIdx.x; i0 < e; i0 += 4 * blockDim.x) {
        int cc[4];
        float vv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t i = i0 + (int64_t)k * blockDim.x;
            const bool in = i < e;
            cc[k] = in ? (int)indices[i] : -1;
            vv[k] = in ? (float)values[i] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = cc[k];
            const float v = vv[k];
            if (c >= c_lo && c < c_hi && c < G)
                row16[c - c_lo] = X16_DIRECT ? __half_as_ushort(__float2half_rn(v))
                                             : (uint16_t)fminf(fmaxf(v, 0.f), 65535.f);
            // the per-cell constant sum_g lgamma(1 + x) is accumulated by part 0 over the whole row
            if (row_const && blockIdx.y == 0 && c >= 0 && c < G && v > 0.f) acc += lgammaf(1.f + v);
        }
    }
    if (X16_DIRECT && threadIdx.x == 0 && G >= c_lo && G < c_hi) row16[G - c_lo] = 0x3C00;   // ones column
    __syncthreads();
    for (int i = threadIdx.x; i < g8_hi - g8_lo; i += blockDim.x) {
        const int c = (g8_lo + i) << 3;
        const uint4 pk = row4[i];
        if (X16_DIRECT) {
            if (c < ldx16) *reinterpret_cast<uint4 *>(x16 + (int64_t)b * ldx16 + c) = pk;
            continue;
        }
        if (t16 && c < ldt16) *reinterpret_cast<uint4 *>(t16 + (int64_t)b * ldt16 + c) = pk;
        if (x16 && c < ldx16) {
            const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float lo = (float)(w[j] & 0xffffu), hi = (float)(w[j] >> 16);
                if (c + 2 * j == G) lo = 1.f;          // augmented ones column
                if (c + 2 * j + 1 == G) hi = 1.f;
                const __half2 h = __floats2half2_rn(lo, hi);
                o[j] = *reinterpret_cast<const uint32_t *>(&h);
            }
            *reinterpret_cast<uint4 *>(x16 + (int64_t)b * ldx16 + c) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
    if (row_const && blockIdx.y == 0) {
        const float tot = block_sum(acc, red);
        if (threadIdx.x == 0) row_const[b] = tot;
    }
}

template <typename IdxT, typename ValT>
static int launch_densify(const int64_t *indptr, const IdxT *indices, const ValT *values, const int64_t *rows,
                          int B, int G, float *x, int64_t ldx, float *row_const, int rebase, uint16_t *t16,
                          int64_t ldt16, __half *x16, int64_t ldx16, cudaStream_t s) {
    if (!x) {   // 16-bit outputs only: single pass through shared memory
        int64_t width = 0;
        if (t16) width = ldt16;
        if (x16 && ldx16 > width) width = ldx16;
        const int width8 = (int)((width + 7) >> 3);
        // one CTA per row while the row fits 64 KB of shared memory (re-scanning the row's
        // non-zeros per column part costs more than the lost occupancy); parts beyond that
        int parts = (width8 + 4095) / 4096;
        if (parts < 1) parts = 1;
        const int part8 = (width8 + parts - 1) / parts;
        const int smem = part8 * 16;
        if (smem <= 200 * 1024) {
            static bool attr_set = false;
            if (!attr_set) {
                cudaFuncSetAttribute(csr_densify16_kernel<IdxT, ValT, false>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
                cudaFuncSetAttribute(csr_densify16_kernel<IdxT, ValT, true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
                attr_set = true;
            }
            if (x16 && !t16)
                csr_densify16_kernel<IdxT, ValT, true><<<dim3(B, parts), 256, smem, s>>>(
                    indptr, indices, values, rows, G, row_const, rebase, t16, ldt16, x16, ldx16, width8, part8);
            else
                csr_densify16_kernel<IdxT, ValT, false><<<dim3(B, parts), 256, smem, s>>>(
                    indptr, indices, values, rows, G, row_const, rebase, t16, ldt16, x16, ldx16, width8, part8);
            SCVAE_CHECK_LAUNCH("csr_densify16");
            return 0;
        }
    }
    csr_densify_kernel<IdxT, ValT><<<B, 256, 0, s>>>(indptr, indices, values, rows, G, x, ldx, row_const, rebase, t16,
                                                     ldt16, x16, ldx16);
    SCVAE_CHECK_LAUNCH("csr_densify");
    return 0;
}

// ---- packed row slabs (hotloop.PackedStream) -----------------------------------------------------
// slab = int32 row_offset[B + 1] | float row_const[B] | (pad to 16) | row strings; a row string is
//   u16 nesc | u16 nnz | u8 blocks[nblk] | (u8 index in block, u8 count)[nnz] | (u16 entry, u16 count)[nesc]
// (little endian), padded to a multiple of 16 bytes: blocks[k] = non-zeros of the row among the genes
// [255 k, 255 k + 255) (so a count fits one byte), a count byte of 255 is an escape whose value is
// looked up by entry position in the row's (sorted) escape list.  ~2 bytes per non-zero.
// One CTA per row: block counts -> prefix sums in shared memory; every entry finds its block by
// binary search of its position in the row; the row is assembled in shared memory as for the CSR
// form and written as fp16 (x16, augmented) and / or uint16 (t16).
constexpr int kPackedBlock = 255;      // genes per block: a block's non-zero count fits one byte
__host__ __device__ inline int64_t packed_rows_offset(int B) { return (((int64_t)8 * B + 4) + 15) & ~(int64_t)15; }
template <bool X16_DIRECT>
__global__ void __launch_bounds__(256)
csr_densify_packed_kernel(const uint8_t *__restrict__ slab, int B, int G, int nblk, float *__restrict__ row_const,
                          uint16_t *__restrict__ t16, int64_t ldt16, __half *__restrict__ x16, int64_t ldx16,
                          int width8) {
    pdl_trigger();       // (as in csr_densify16_kernel)
    extern __shared__ __align__(16) uint16_t row16[];
    __shared__ int prefix[260];
    __shared__ int warp_tot[8];
    const int b = blockIdx.x;
    const int32_t *ro = reinterpret_cast<const int32_t *>(slab);
    const uint8_t *str = slab + packed_rows_offset(B) + ro[b];
    const int nesc = (int)str[0] | ((int)str[1] << 8);
    const int nnz = (int)str[2] | ((int)str[3] << 8);
    const uint8_t *blocks = str + 4;
    const uint8_t *entries = blocks + nblk;
    const uint8_t *esc = entries + 2 * nnz;
    uint4 *row4 = reinterpret_cast<uint4 *>(row16);
    for (int i = threadIdx.x; i < width8; i += blockDim.x) row4[i] = make_uint4(0u, 0u, 0u, 0u);
    // inclusive scan of the block counts (nblk <= 256: one element per thread)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        int v = (int)threadIdx.x < nblk ? (int)blocks[threadIdx.x] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += n;
        }
        if (lane == 31) warp_tot[warp] = v;
        __syncthreads();
        int base = 0;
        for (int w = 0; w < warp; ++w) base += warp_tot[w];
        if (threadIdx.x == 0) prefix[0] = 0;
        prefix[threadIdx.x + 1] = v + base;       // prefix[k] = entries of the blocks < k
    }
    __syncthreads();
    for (int i0 = threadIdx.x; i0 < nnz; i0 += 4 * blockDim.x) {
        int lo8[4], val[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = i0 + k * (int)blockDim.x;
            const bool in = i < nnz;
            const uint8_t *p = entries + 2 * i;
            lo8[k] = in ? (int)p[0] : -1;
            val[k] = in ? (int)p[1] : 0;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (lo8[k] < 0) continue;
            const int pos = i0 + k * (int)blockDim.x;
            if (val[k] == 255) {                         // escape: counts >= 255, by entry position
                int lo = 0, hi = nesc - 1;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    const int at = (int)esc[4 * mid] | ((int)esc[4 * mid + 1] << 8);
                    if (at < pos) lo = mid + 1;
                    else hi = mid;
                }
                val[k] = (int)esc[4 * lo + 2] | ((int)esc[4 * lo + 3] << 8);
            }
            int lo = 0, hi = nblk;                       // largest blk with prefix[blk] <= pos
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (prefix[mid] <= pos) lo = mid;
                else hi = mid;
            }
            const int c = lo * kPackedBlock + lo8[k];
            if (c < G)
                row16[c] = X16_DIRECT ? __half_as_ushort(__float2half_rn((float)val[k])) : (uint16_t)val[k];
        }
    }
    if (X16_DIRECT && threadIdx.x == 0) row16[G] = 0x3C00;   // ones column
    if (row_const && threadIdx.x == 0) row_const[b] = reinterpret_cast<const float *>(slab + 4 * (B + 1))[b];
    __syncthreads();
    for (int i = threadIdx.x; i < width8; i += blockDim.x) {
        const int c = i << 3;
        const uint4 pk = row4[i];
        if (X16_DIRECT) {
            if (c < ldx16) *reinterpret_cast<uint4 *>(x16 + (int64_t)b * ldx16 + c) = pk;
            continue;
        }
        if (t16 && c < ldt16) *reinterpret_cast<uint4 *>(t16 + (int64_t)b * ldt16 + c) = pk;
        if (x16 && c < ldx16) {
            const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float lo = (float)(w[j] & 0xffffu), hi = (float)(w[j] >> 16);
                if (c + 2 * j == G) lo = 1.f;          // augmented ones column
                if (c + 2 * j + 1 == G) hi = 1.f;
                const __half2 h = __floats2half2_rn(lo, hi);
                o[j] = *reinterpret_cast<const uint32_t *>(&h);
            }
            *reinterpret_cast<uint4 *>(x16 + (int64_t)b * ldx16 + c) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// ---- device-side assembly of a packed slab: the GPU pulls the minibatch's row strings itself -------
// store: all row strings (16-byte aligned, 16-byte multiples) in PINNED HOST memory, read over PCIe
// through its unified address; row_off / row_const_all / order live in device memory.  Kernel 1 (one
// CTA) turns the B string lengths into the slab's offset table, kernel 2 (one CTA per row) copies the
// strings with 128-bit loads.  No host work per step beyond the two launches.
__global__ void __launch_bounds__(1024)
packed_offsets_kernel(const int64_t *__restrict__ row_off, const float *__restrict__ row_const_all,
                      const int64_t *__restrict__ order, int B, uint8_t *__restrict__ slab) {
    __shared__ int warp_sum_s[32];
    __shared__ int carry_s;
    int32_t *ro = reinterpret_cast<int32_t *>(slab);
    float *rc = reinterpret_cast<float *>(slab + 4 * (int64_t)(B + 1));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int r0 = 0; r0 < B; r0 += 1024) {
        const int r = r0 + threadIdx.x;
        int len = 0;
        if (r < B) {
            const int64_t i = order[r];
            len = (int)(row_off[i + 1] - row_off[i]);
            rc[r] = row_const_all[i];
        }
        int v = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += n;
        }
        if (lane == 31) warp_sum_s[warp] = v;
        __syncthreads();
        int base = carry_s;
        for (int w = 0; w < warp; ++w) base += warp_sum_s[w];
        if (r < B) ro[r] = base + v - len;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = base + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) ro[B] = carry_s;
}

// <= 32 registers and no shared memory, so that a CTA fits beside ANY kernel of the training step it
// runs next to (the fused heads kernel leaves 4096 registers per SM): the pull is hidden behind the
// previous step only if its CTAs are resident while that step's kernels are.
__global__ void __launch_bounds__(128, 16)
packed_pull_kernel(const uint8_t *__restrict__ store, const int64_t *__restrict__ row_off,
                   const int64_t *__restrict__ order, int B, uint8_t *__restrict__ slab, int64_t capacity) {
    const int32_t *ro = reinterpret_cast<const int32_t *>(slab);
    const int64_t base = packed_rows_offset(B);
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const int64_t i = order[b];
        const int64_t off = row_off[i];
        const int n16 = (int)((row_off[i + 1] - off) >> 4);
        if (base + ro[b] + ((int64_t)n16 << 4) > capacity) continue;     // (never: the buffer holds the worst case)
        const uint4 *src = reinterpret_cast<const uint4 *>(store + off);
        uint4 *dst = reinterpret_cast<uint4 *>(slab + base + ro[b]);
        // several 128-bit loads per thread in flight: the latency is a PCIe round trip
        int k = threadIdx.x;
        for (; k + 3 * 128 < n16; k += 4 * 128) {
            const uint4 v0 = src[k], v1 = src[k + 128], v2 = src[k + 256], v3 = src[k + 384];
            dst[k] = v0; dst[k + 128] = v1; dst[k + 256] = v2; dst[k + 384] = v3;
        }
        for (; k < n16; k += 128) dst[k] = src[k];
    }
}

// sum_g lgamma(1 + x) of every CSR row (one warp per row): a per-cell constant of the data set,
// computed once when the matrix is loaded instead of in every minibatch assembly
template <typename ValT>
__global__ void __launch_bounds__(256)
csr_row_constants_kernel(const int64_t *__restrict__ indptr, const ValT *__restrict__ values, int64_t n_rows,
                         float *__restrict__ out) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int lane = threadIdx.x & 31;
    float acc = 0.f;
    for (int64_t i = indptr[row] + lane; i < indptr[row + 1]; i += 32) {
        const float v = (float)values[i];
        if (v > 0.f) acc += lgammaf(1.f + v);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[row] = acc;
}

__global__ void gather_f32_kernel(const float *__restrict__ src, const int64_t *__restrict__ rows, int B,
                                  float *__restrict__ dst) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) dst[b] = src[rows ? rows[b] : b];
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
                                 float *row_const, int rebase, void *t16, int64_t ldt16, void *x16,
                                 int64_t ldx16, void *stream) {
    using namespace scvae;
    SCVAE_CHECK_ARG(indptr && indices && values && (x || x16 || t16), "csr_densify: NULL pointer");
    SCVAE_CHECK_ARG(B >= 0 && G > 0 && (!x || ldx >= G), "csr_densify: bad shape (B=%d G=%d ldx=%lld)", B, G,
                    (long long)ldx);
    SCVAE_CHECK_ARG(!t16 || (ldt16 % 8 == 0 && ldt16 >= G && aligned16(t16)), "csr_densify: bad t16 layout");
    SCVAE_CHECK_ARG(!x16 || (ldx16 % 8 == 0 && ldx16 > G && aligned16(x16)), "csr_densify: bad x16 layout");
    if (B == 0) return 0;
    return launch_densify<int32_t, float>(indptr, indices, values, rows, B, G, x, ldx, row_const, rebase,
                                          (uint16_t *)t16, ldt16, (__half *)x16, ldx16, (cudaStream_t)stream);
}

extern "C" int scvae_csr_densify_u16(const int64_t *indptr, const void *indices_u16, const void *values_u16,
                                     const int64_t *rows, int B, int G, float *x, int64_t ldx, float *row_const,
                                     int rebase, void *t16, int64_t ldt16, void *x16, int64_t ldx16,
                                     void *stream) {
    using namespace scvae;
    SCVAE_CHECK_ARG(indptr && indices_u16 && values_u16 && (x || x16 || t16), "csr_densify_u16: NULL pointer");
    SCVAE_CHECK_ARG(B >= 0 && G > 0 && G <= 65536 && (!x || ldx >= G), "csr_densify_u16: bad shape");
    SCVAE_CHECK_ARG(!t16 || (ldt16 % 8 == 0 && ldt16 >= G && aligned16(t16)), "csr_densify_u16: bad t16 layout");
    SCVAE_CHECK_ARG(!x16 || (ldx16 % 8 == 0 && ldx16 > G && aligned16(x16)), "csr_densify_u16: bad x16 layout");
    if (B == 0) return 0;
    return launch_densify<uint16_t, uint16_t>(indptr, (const uint16_t *)indices_u16, (const uint16_t *)values_u16,
                                              rows, B, G, x, ldx, row_const, rebase, (uint16_t *)t16, ldt16,
                                              (__half *)x16, ldx16, (cudaStream_t)stream);
}

extern "C" int64_t scvae_packed_rows_offset(int B) { return scvae::packed_rows_offset(B); }

extern "C" int scvae_csr_densify_packed(const void *slab, int B, int G, float *row_const, void *t16, int64_t ldt16,
                                        void *x16, int64_t ldx16, void *stream) {
    using namespace scvae;
    SCVAE_CHECK_ARG(slab && (x16 || t16) && B >= 0 && G > 0 && G <= 65280, "csr_densify_packed: bad arguments");
    SCVAE_CHECK_ARG((reinterpret_cast<uintptr_t>(slab) & 3u) == 0, "csr_densify_packed: the slab must be 4-byte aligned");
    SCVAE_CHECK_ARG(!t16 || (ldt16 % 8 == 0 && ldt16 >= G && aligned16(t16)), "csr_densify_packed: bad t16 layout");
    SCVAE_CHECK_ARG(!x16 || (ldx16 % 8 == 0 && ldx16 > G && aligned16(x16)), "csr_densify_packed: bad x16 layout");
    if (B == 0) return 0;
    int64_t width = 0;
    if (t16) width = ldt16;
    if (x16 && ldx16 > width) width = ldx16;
    const int width8 = (int)((width + 7) >> 3);
    const int smem = width8 * 16;
    SCVAE_CHECK_ARG(smem <= 200 * 1024, "csr_densify_packed: row too wide");
    const int nblk = (G + kPackedBlock - 1) / kPackedBlock;
    const uint8_t *sl = (const uint8_t *)slab;
    cudaStream_t s = (cudaStream_t)stream;
#define PACKED(DIRECT)                                                                                            \
    do {                                                                                                          \
        static bool attr_set = false;                                                                             \
        if (!attr_set) {                                                                                          \
            cudaFuncSetAttribute(csr_densify_packed_kernel<DIRECT>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                 200 * 1024);                                                                     \
            attr_set = true;                                                                                      \
        }                                                                                                         \
        csr_densify_packed_kernel<DIRECT><<<B, 256, smem, s>>>(sl, B, G, nblk, row_const, (uint16_t *)t16, ldt16, \
                                                               (__half *)x16, ldx16, width8);                     \
    } while (0)
    if (x16 && !t16) PACKED(true);
    else PACKED(false);
#undef PACKED
    SCVAE_CHECK_LAUNCH("csr_densify_packed");
    return 0;
}

extern "C" int scvae_packed_pull(const void *store, const int64_t *row_off, const float *row_const_all,
                                 const int64_t *order, int B, void *slab, int64_t slab_capacity, void *stream) {
    using namespace scvae;
    SCVAE_CHECK_ARG(store && row_off && row_const_all && order && slab && B >= 0, "packed_pull: bad arguments");
    SCVAE_CHECK_ARG(aligned16(store) && aligned16(slab) && slab_capacity >= packed_rows_offset(B),
                    "packed_pull: store and slab must be 16-byte aligned, the slab large enough");
    if (B == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    packed_offsets_kernel<<<1, 1024, 0, s>>>(row_off, row_const_all, order, B, (uint8_t *)slab);
    SCVAE_CHECK_LAUNCH("packed_offsets");
    // persistent CTAs walk the rows
    const int pull_ctas = B < 148 ? B : 148;      // one per SM: the step's big kernels leave room for exactly one
    packed_pull_kernel<<<pull_ctas, 128, 0, s>>>((const uint8_t *)store, row_off, order, B, (uint8_t *)slab,
                                                  slab_capacity);
    SCVAE_CHECK_LAUNCH("packed_pull");
    return 0;
}

extern "C" int scvae_csr_row_constants(const int64_t *indptr, const void *values, int values_are_u16,
                                       int64_t n_rows, float *out, void *stream) {
    using namespace scvae;
    SCVAE_CHECK_ARG(indptr && values && out && n_rows >= 0, "csr_row_constants: bad arguments");
    if (n_rows == 0) return 0;
    const unsigned blocks = (unsigned)((n_rows + 7) / 8);
    if (values_are_u16)
        csr_row_constants_kernel<uint16_t><<<blocks, 256, 0, (cudaStream_t)stream>>>(
            indptr, (const uint16_t *)values, n_rows, out);
    else
        csr_row_constants_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(indptr, (const float *)values,
                                                                                  n_rows, out);
    SCVAE_CHECK_LAUNCH("csr_row_constants");
    return 0;
}

extern "C" int scvae_gather_f32(const float *src, const int64_t *rows, int B, float *dst, void *stream) {
    using namespace scvae;
    SCVAE_CHECK_ARG(src && dst && B >= 0, "gather_f32: bad arguments");
    if (B == 0) return 0;
    gather_f32_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(src, rows, B, dst);
    SCVAE_CHECK_LAUNCH("gather_f32");
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
