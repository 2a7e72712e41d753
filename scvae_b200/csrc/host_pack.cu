// Host side of the streamed data path (hotloop.PackedStream): assembles the packed row slab of one
// minibatch in (pinned) host memory from the per-row strings that were encoded once per data set.
// Replaces, for matrices that stay in host memory, the reference's per-step
// `x_train[batch_indices].toarray()` + feed_dict copy (VAE:985-1029) with a gather of ~2 bytes per
// non-zero.  Pure host code (no device work): a prefix sum over the rows' lengths, then the row
// strings are copied by a few threads.
#include <cstring>
#include <thread>
#include <vector>

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
    auto copy = [&](int r0, int r1) {
        for (int r = r0; r < r1; ++r) {
            const int64_t i = order[r];
            std::memcpy(out + ro[r], store + row_off[i], (size_t)(row_off[i + 1] - row_off[i]));
        }
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
