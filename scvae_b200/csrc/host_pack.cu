// Host side of the streamed data path (hotloop.PackedStream): assembles the packed row slab of one
// minibatch in (pinned) host memory from the per-row strings that were encoded once per data set.
// Replaces, for matrices that stay in host memory, the reference's per-step
// `x_train[batch_indices].toarray()` + feed_dict copy (VAE:985-1029) with a gather of ~2 bytes per
// non-zero.  Pure host code (no device work): a prefix sum over the rows' lengths, then the row
// strings are copied by a few threads.
#include <cstring>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

#include "common.cuh"

extern "C" int64_t scvae_packed_rows_offset(int B);

// store: all row strings back to back; row_off[n + 1]: their byte offsets; order[rows]: the rows of
// this slab.  Writes the slab (layout: scvae_csr_densify_packed) to dst; returns its size in bytes
// through *bytes_out.  Fails (1) when dst_capacity is too small.
extern "C" int scvae_pack_row_slab(const uint8_t *store, const int64_t *row_off, const float *row_const_all,
                                   const int64_t *order, int rows, int64_t n_rows, uint8_t *dst, int64_t dst_capacity,
                                   int threads, int64_t *bytes_out) {
    using namespace scvae;
    SCVAE_CHECK_ARG(store && row_off && row_const_all && order && dst && bytes_out && rows >= 0,
                    "pack_row_slab: bad arguments");
    const int64_t base = scvae_packed_rows_offset(rows);
    SCVAE_CHECK_ARG(dst_capacity >= base, "pack_row_slab: destination too small");
    int32_t *ro = reinterpret_cast<int32_t *>(dst);
    float *rc = reinterpret_cast<float *>(dst + 4 * (int64_t)(rows + 1));
    int64_t at = 0;
    for (int r = 0; r < rows; ++r) {
        const int64_t i = order[r];
        SCVAE_CHECK_ARG(i >= 0 && i < n_rows, "pack_row_slab: row index out of range");
        ro[r] = (int32_t)at;
        rc[r] = row_const_all[i];
        at += row_off[i + 1] - row_off[i];
        SCVAE_CHECK_ARG(at < ((int64_t)1 << 31), "pack_row_slab: slab beyond 2 GiB");
    }
    ro[rows] = (int32_t)at;
    SCVAE_CHECK_ARG(base + at <= dst_capacity, "pack_row_slab: destination too small");
    uint8_t *out = dst + base;
    // strings and destinations are 16-byte aligned 16-byte multiples: streaming (non-temporal) stores
    // keep the destination lines out of the cache (no read-for-ownership, the DMA reads them next),
    // and the next row's head is prefetched while this one is copied
    const bool aligned = ((reinterpret_cast<uintptr_t>(store) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    auto copy = [&](int r0, int r1) {
        for (int r = r0; r < r1; ++r) {
            const int64_t i = order[r];
            const size_t len = (size_t)(row_off[i + 1] - row_off[i]);
            const uint8_t *src = store + row_off[i];
#if defined(__x86_64__)
            if (r + 1 < r1) {
                const uint8_t *nxt = store + row_off[order[r + 1]];
                _mm_prefetch((const char *)nxt, _MM_HINT_T0);
                _mm_prefetch((const char *)nxt + 64, _MM_HINT_T0);
                _mm_prefetch((const char *)nxt + 128, _MM_HINT_T0);
                _mm_prefetch((const char *)nxt + 192, _MM_HINT_T0);
            }
            if (aligned && (len & 15u) == 0 && (row_off[i] & 15) == 0 && (ro[r] & 15) == 0) {
                const __m128i *s16 = reinterpret_cast<const __m128i *>(src);
                __m128i *d16 = reinterpret_cast<__m128i *>(out + ro[r]);
                const size_t n16 = len >> 4;
                size_t k = 0;
                for (; k + 4 <= n16; k += 4) {
                    const __m128i a = _mm_load_si128(s16 + k), b = _mm_load_si128(s16 + k + 1);
                    const __m128i c = _mm_load_si128(s16 + k + 2), d = _mm_load_si128(s16 + k + 3);
                    _mm_stream_si128(d16 + k, a);
                    _mm_stream_si128(d16 + k + 1, b);
                    _mm_stream_si128(d16 + k + 2, c);
                    _mm_stream_si128(d16 + k + 3, d);
                }
                for (; k < n16; ++k) _mm_stream_si128(d16 + k, _mm_load_si128(s16 + k));
                continue;
            }
#endif
            std::memcpy(out + ro[r], src, len);
        }
#if defined(__x86_64__)
        _mm_sfence();
#endif
    };
    if (threads < 1) threads = 1;
    if (threads > 16) threads = 16;
    if (threads == 1 || rows < 4 * threads) {
        copy(0, rows);
    } else {
        std::vector<std::thread> pool;
        pool.reserve(threads - 1);
        const int per = (rows + threads - 1) / threads;
        for (int t = 1; t < threads; ++t) {
            const int r0 = t * per, r1 = r0 + per < rows ? r0 + per : rows;
            if (r0 < r1) pool.emplace_back(copy, r0, r1);
        }
        copy(0, per < rows ? per : rows);
        for (auto &th : pool) th.join();
    }
    *bytes_out = base + at;
    return 0;
}

// The same slab assembled by the COPY ENGINE: the header (offsets, per-cell constants) is written into
// a small pinned staging area, then ONE cudaMemcpyBatchAsync ships the header and the B row strings
// from the pinned `store` straight to their places in the device slab -- no host gather, no second
// pass over the bytes.  `header` (pinned, >= scvae_packed_rows_offset(rows) bytes) must stay untouched
// until the batch has executed (the caller rotates a few of them).  Returns 2 when the runtime does
// not offer batched copies (the caller falls back to scvae_pack_row_slab + one copy).
extern "C" int scvae_packed_copy_batch(const uint8_t *store, const int64_t *row_off, const float *row_const_all,
                                       const int64_t *order, int rows, int64_t n_rows, uint8_t *header,
                                       uint8_t *slab_dev, int64_t slab_capacity, int device, void *stream,
                                       int64_t *bytes_out) {
    using namespace scvae;
    SCVAE_CHECK_ARG(store && row_off && row_const_all && order && header && slab_dev && bytes_out && rows > 0,
                    "packed_copy_batch: bad arguments");
#if CUDART_VERSION >= 12080
    const int64_t base = scvae_packed_rows_offset(rows);
    int32_t *ro = reinterpret_cast<int32_t *>(header);
    float *rc = reinterpret_cast<float *>(header + 4 * (int64_t)(rows + 1));
    static thread_local std::vector<void *> dsts, srcs;
    static thread_local std::vector<size_t> sizes;
    dsts.clear(); srcs.clear(); sizes.clear();
    dsts.reserve(rows + 1); srcs.reserve(rows + 1); sizes.reserve(rows + 1);
    int64_t at = 0;
    for (int r = 0; r < rows; ++r) {
        const int64_t i = order[r];
        SCVAE_CHECK_ARG(i >= 0 && i < n_rows, "packed_copy_batch: row index out of range");
        const int64_t len = row_off[i + 1] - row_off[i];
        ro[r] = (int32_t)at;
        rc[r] = row_const_all[i];
        if (len > 0) {
            dsts.push_back(slab_dev + base + at);
            srcs.push_back(const_cast<uint8_t *>(store + row_off[i]));
            sizes.push_back((size_t)len);
        }
        at += len;
        SCVAE_CHECK_ARG(at < ((int64_t)1 << 31), "packed_copy_batch: slab beyond 2 GiB");
    }
    ro[rows] = (int32_t)at;
    SCVAE_CHECK_ARG(base + at <= slab_capacity, "packed_copy_batch: destination too small");
    dsts.push_back(slab_dev);
    srcs.push_back(header);
    sizes.push_back((size_t)(8 * (int64_t)rows + 4));
    cudaMemcpyAttributes attr;
    std::memset(&attr, 0, sizeof(attr));
    attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;      // (pinned / device pointers: no location hints)
    (void)device;
    size_t attr_idx = 0, fail = 0;
    const cudaError_t e = cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), dsts.size(), &attr, &attr_idx, 1,
                                               &fail, (cudaStream_t)stream);
    if (e != cudaSuccess) cudaGetLastError();      // (not sticky: do not leave it for the next launch check)
    if (e == cudaErrorNotSupported || e == cudaErrorCallRequiresNewerDriver) return 2;
    SCVAE_CHECK_ARG(stream != nullptr || e == cudaSuccess,
                    "packed_copy_batch: batched copies need a non-default stream (%s)", cudaGetErrorString(e));
    SCVAE_CHECK_ARG(e == cudaSuccess, "packed_copy_batch: cudaMemcpyBatchAsync failed at copy %zu: %s", fail,
                    cudaGetErrorString(e));
    *bytes_out = base + at;
    return 0;
#else
    (void)slab_capacity; (void)device; (void)stream; (void)n_rows;
    return 2;
#endif
}
