// Data-parallel exchange fused with the optimiser: ONE kernel per rank does
//     reduce-scatter of the flat gradient   (P2P loads of every peer's slice over NVLink)
//  -> elementwise clip + TensorFlow Adam     (VAE:2736-2770) on this rank's 1/W slice
//  -> all-gather of the updated parameters   (P2P stores into every peer's parameter buffer)
// instead of  ncclAllReduce(grad) ; adam(all parameters on every rank).  Every element of the
// gradient crosses NVLink once in each direction (as in a ring all-reduce), the Adam pass and
// its 7 x 4 B / parameter of HBM traffic shrink by 1/W, and the reduction order is fixed (rank
// 0, 1, ..., W-1), so replicas stay bit-identical.  The reference has no distributed code; this
// replaces the AdamOptimizer apply of a single process for W processes.
//
// Cross-GPU synchronisation (no host involvement, CUDA-graph capturable): two flag barriers per
// launch on a symmetric flag buffer, sequence-numbered so that flags never need resetting:
//   entry  "my gradients are complete"            - then peers' gradient slices are read;
//   exit   "my parameter stores to you are done"  - then the stream may go on to read them.
// Waiting is bounded: after 30 s without progress a rank sets an error word and carries on
// (results are then invalid, but the GPU is not left hanging).
#include "common.cuh"

namespace scvae {

constexpr int kDpMaxWorld = 16;
struct DpPeers {
    const float *grad[kDpMaxWorld];
    float *param[kDpMaxWorld];
    uint32_t *flags[kDpMaxWorld];      // per rank: [2 barriers][kDpMaxWorld] sequence numbers
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// peer gradient: written by another GPU, must not be served from a stale local cache line
__device__ __forceinline__ float4 ld_peer4(const float *p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
// returns false on timeout.  The bound is wall time (30 s): ranks legitimately drift apart by seconds
// between steps (one of them writes a checkpoint, encodes a data set, captures a graph); only a rank
// that is gone should trip it.  The host checks the error word (PeerExchange.timed_out) every epoch.
constexpr unsigned long long kFlagWaitNs = 30ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ bool wait_flag(const uint32_t *f, uint32_t seq) {
    unsigned long long t0 = 0;
    for (unsigned it = 0;; ++it) {
        // sequence numbers only grow; signed distance tolerates wrap-around
        if ((int32_t)(ld_acquire_sys(f) - seq) >= 0) return true;
        __nanosleep(100);
        if ((it & 1023u) == 1023u) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > kFlagWaitNs) return false;
        }
    }
}

// ctl (local, zero-initialised): [0] last completed sequence number, [1] CTAs done, [2] error
// WORLD > 0: compile-time rank count, U elements (float4) per thread and trip with all W x U peer
// loads in flight before the first use (NVLink latency is ~2 us: memory-level parallelism is
// everything here); WORLD == 0: any rank count, one element per trip.
template <int WORLD, int U>
__global__ void __launch_bounds__(256)
dp_reduce_adam_kernel(DpPeers peers, int world_rt, int rank, float *__restrict__ m, float *__restrict__ v, int64_t n,
                      const int64_t *__restrict__ step, float lr, float beta1, float beta2, float eps, float clip,
                      float gscale, const float *__restrict__ scalars, uint32_t *ctl) {
    const int world = WORLD > 0 ? WORLD : world_rt;
    __shared__ float s_lr_t;
    __shared__ uint32_t s_seq;
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        const double t = (double)(*step + 1);
        const double lr_eff = (double)lr * (scalars ? (double)scalars[0] : 1.0);
        s_lr_t = adam_step_size(lr_eff, beta1, beta2, t);
        s_seq = ctl[0] + 1u;
    }
    __syncthreads();
    const float lr_t = s_lr_t;
    const uint32_t seq = s_seq;

    // ---- entry barrier: every rank's gradient buffer is complete -------------------------
    if (blockIdx.x == 0 && threadIdx.x < world) st_release_sys(peers.flags[threadIdx.x] + rank, seq);
    if (threadIdx.x < world) {
        if (!wait_flag(peers.flags[rank] + threadIdx.x, seq)) ctl[2] = 1u;
    }
    __syncthreads();

    // ---- this rank's slice: sum over ranks in fixed order, clip, Adam, broadcast ---------
    const int64_t n4 = n >> 2;                                   // n % 4 == 0 (checked on the host)
    const int64_t per = (n4 + world - 1) / world;
    const int64_t lo = (int64_t)rank * per, hi = (lo + per < n4) ? lo + per : n4;
    const float ob1 = 1.f - beta1, ob2 = 1.f - beta2;
    auto update = [&](float &pi, float gi, float &mi, float &vi) {
        gi = fminf(fmaxf(gi * gscale, -clip), clip);
        mi = beta1 * mi + ob1 * gi;
        vi = beta2 * vi + ob2 * gi * gi;
        pi = pi - lr_t * mi / (sqrtf(vi) + eps);
    };
    auto finish = [&](int64_t i, float4 g) {
        float4 p4 = reinterpret_cast<const float4 *>(peers.param[rank])[i];
        float4 m4 = reinterpret_cast<float4 *>(m)[i];
        float4 v4 = reinterpret_cast<float4 *>(v)[i];
        update(p4.x, g.x, m4.x, v4.x);
        update(p4.y, g.y, m4.y, v4.y);
        update(p4.z, g.z, m4.z, v4.z);
        update(p4.w, g.w, m4.w, v4.w);
        reinterpret_cast<float4 *>(m)[i] = m4;
        reinterpret_cast<float4 *>(v)[i] = v4;
        if (WORLD > 0) {
#pragma unroll
            for (int r = 0; r < (WORLD > 0 ? WORLD : 1); ++r) reinterpret_cast<float4 *>(peers.param[r])[i] = p4;
        } else {
            for (int r = 0; r < world; ++r) reinterpret_cast<float4 *>(peers.param[r])[i] = p4;
        }
    };
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t first = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (WORLD > 0) {
        constexpr int W = WORLD > 0 ? WORLD : 1;
        for (int64_t i0 = first; i0 < hi; i0 += (int64_t)U * stride) {
            float4 x[U][W];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = i0 + (int64_t)u * stride;
                if (i < hi) {
#pragma unroll
                    for (int r = 0; r < W; ++r) x[u][r] = ld_peer4(peers.grad[r] + 4 * i);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = i0 + (int64_t)u * stride;
                if (i < hi) {
                    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int r = 0; r < W; ++r) {     // fixed order: rank 0, 1, ...
                        g.x += x[u][r].x; g.y += x[u][r].y; g.z += x[u][r].z; g.w += x[u][r].w;
                    }
                    finish(i, g);
                }
            }
        }
    } else {
        for (int64_t i = first; i < hi; i += stride) {
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < world; ++r) {
                const float4 x = ld_peer4(peers.grad[r] + 4 * i);
                g.x += x.x; g.y += x.y; g.z += x.z; g.w += x.w;
            }
            finish(i, g);
        }
    }

    // ---- exit barrier: my stores have reached every peer, every peer's have reached me ----
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&ctl[1], 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if (threadIdx.x < world) {
            st_release_sys(peers.flags[threadIdx.x] + kDpMaxWorld + rank, seq);
            if (!wait_flag(peers.flags[rank] + kDpMaxWorld + threadIdx.x, seq)) ctl[2] = 1u;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            ctl[1] = 0u;
            ctl[0] = seq;
            __threadfence();
        }
    }
}

}  // namespace scvae

extern "C" int scvae_dp_reduce_adam(int world, int rank, const void *const *grad_ptrs, void *const *param_ptrs,
                                    void *const *flag_ptrs, float *m, float *v, int64_t n, const int64_t *step,
                                    float lr, float beta1, float beta2, float epsilon, float clip, float grad_scale,
                                    const float *scalars, void *ctl, int max_ctas, void *stream) {
    using namespace scvae;
    SCVAE_CHECK_ARG(world >= 1 && world <= kDpMaxWorld && rank >= 0 && rank < world,
                    "dp_reduce_adam: world %d / rank %d out of range (max %d ranks)", world, rank, kDpMaxWorld);
    SCVAE_CHECK_ARG(grad_ptrs && param_ptrs && flag_ptrs && m && v && step && ctl, "dp_reduce_adam: NULL pointer");
    SCVAE_CHECK_ARG(n >= 0 && n % 4 == 0, "dp_reduce_adam: the range must be a multiple of 4 floats");
    if (n == 0) return 0;
    DpPeers peers;
    for (int r = 0; r < kDpMaxWorld; ++r) {
        peers.grad[r] = r < world ? (const float *)grad_ptrs[r] : nullptr;
        peers.param[r] = r < world ? (float *)param_ptrs[r] : nullptr;
        peers.flags[r] = r < world ? (uint32_t *)flag_ptrs[r] : nullptr;
        if (r < world)
            SCVAE_CHECK_ARG(peers.grad[r] && peers.param[r] && peers.flags[r] && aligned16(peers.grad[r]) &&
                                aligned16(peers.param[r]),
                            "dp_reduce_adam: peer %d buffers missing or not 16-byte aligned", r);
    }
    SCVAE_CHECK_ARG(aligned16(m) && aligned16(v), "dp_reduce_adam: Adam slots must be 16-byte aligned");
    const int64_t per = ((n >> 2) + world - 1) / world;
    int64_t blocks = (per + 255) / 256;
    const int cap = max_ctas > 0 ? max_ctas : 148;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
#define DP_LAUNCH(W, UU)                                                                                      \
    dp_reduce_adam_kernel<W, UU><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(                               \
        peers, world, rank, m, v, n, step, lr, beta1, beta2, epsilon, clip, grad_scale, scalars, (uint32_t *)ctl)
    switch (world) {
        case 2: DP_LAUNCH(2, 8); break;
        case 4: DP_LAUNCH(4, 4); break;
        case 8: DP_LAUNCH(8, 2); break;
        default: DP_LAUNCH(0, 1); break;
    }
#undef DP_LAUNCH
    SCVAE_CHECK_LAUNCH("dp_reduce_adam");
    return 0;
}
