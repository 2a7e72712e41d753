"""Minibatch loops around the step engine: device-resident and streamed CSR inputs,
CUDA-graph replay of the training step.

Replaces the reference's per-step host work (VAE:985-1029): ``x_train[idx].toarray()`` on
the host, a feed_dict copy of two dense (B, G) fp32 arrays and a ``session.run``.  Here the
count matrix stays CSR; either all of it lives in HBM and a step ships only row indices
(``ResidentCSR``), or each step's row slab is copied from pinned host memory on a side stream
while the previous step computes (``StreamedCSR``).  The step itself (densify -> noise ->
forward -> backward -> clip+Adam) is captured once per minibatch shape into a CUDA graph.
"""

import numpy
import torch

from . import kernels as K


def _as_csr_arrays(matrix):
    """scipy CSR (or anything with indptr/indices/data) -> (int64, int32, float32) arrays."""
    import scipy.sparse
    if not scipy.sparse.isspmatrix_csr(matrix):
        matrix = scipy.sparse.csr_matrix(matrix)
    return (numpy.ascontiguousarray(matrix.indptr, dtype=numpy.int64),
            numpy.ascontiguousarray(matrix.indices, dtype=numpy.int32),
            numpy.ascontiguousarray(matrix.data, dtype=numpy.float32),
            matrix.shape)


def _counts_fit_u16(data):
    """Integer counts below 65536: the 16-bit target copy of the fused heads is exact."""
    return bool(data.size == 0 or (data.min() >= 0 and data.max() <= 65504
                                   and numpy.array_equal(data, numpy.floor(data))))


class ResidentCSR:
    """The whole (cells x genes) count matrix in HBM as CSR; steps gather rows by index."""

    def __init__(self, matrix, device):
        indptr, indices, data, shape = _as_csr_arrays(matrix)
        self.shape = shape
        self.device = torch.device(device)
        self.indptr = torch.from_numpy(indptr).to(self.device)
        self.indices = torch.from_numpy(indices).to(self.device)
        self.values = torch.from_numpy(data).to(self.device)
        self.nbytes = indptr.nbytes + indices.nbytes + data.nbytes
        self.u16_ok = _counts_fit_u16(data)
        self.f16_exact = bool(self.u16_ok and (data.size == 0 or data.max() <= 2048))
        # per-cell constant of every count log-likelihood, once per data set
        self.row_const = torch.zeros(shape[0], dtype=torch.float32, device=self.device)
        K.csr_row_constants(self.indptr, self.values, self.row_const)
        # optional per-cell decoder features (set by the model shell): batch ids as floats,
        # normalised count sums
        self.targets = None                 # ResidentCSR of likelihood targets when they differ from x
        self.batch_index = None
        self.count_sum_feature = None
        self.count_sum_parameter = None     # raw count sums: N of the constrained Poisson

    def set_features(self, batch_indices=None, count_sum_feature=None, count_sum_parameter=None):
        if count_sum_parameter is not None:
            self.count_sum_parameter = torch.as_tensor(
                numpy.asarray(count_sum_parameter).reshape(-1), dtype=torch.float32).to(self.device)
        if batch_indices is not None:
            self.batch_index = torch.as_tensor(numpy.asarray(batch_indices).reshape(-1),
                                               dtype=torch.float32).to(self.device)
        if count_sum_feature is not None:
            self.count_sum_feature = torch.as_tensor(
                numpy.asarray(count_sum_feature).reshape(-1), dtype=torch.float32).to(self.device)
        return self

    @property
    def number_of_examples(self):
        return self.shape[0]


class StreamedCSR:
    """CSR kept in pinned host memory; ``fetch(i0, i1)`` copies the slab of rows [i0, i1) to
    one of two device staging buffers on a copy stream (double buffering)."""

    def __init__(self, matrix, device, max_rows):
        indptr, indices, data, shape = _as_csr_arrays(matrix)
        self.shape = shape
        self.device = torch.device(device)
        self.u16_ok = _counts_fit_u16(data)
        # compact wire format (4 instead of 8 bytes per non-zero) when columns and counts fit 16 bits
        self.compact = bool(self.u16_ok and shape[1] <= 65536)
        self.indptr = torch.from_numpy(indptr).pin_memory()
        if self.compact:
            self.indices = torch.from_numpy(indices.astype(numpy.uint16).view(numpy.int16)).pin_memory()
            self.values = torch.from_numpy(data.astype(numpy.uint16).view(numpy.int16)).pin_memory()
        else:
            self.indices = torch.from_numpy(indices).pin_memory()
            self.values = torch.from_numpy(data).pin_memory()
        self.bytes_per_nnz = 4 if self.compact else 8
        # per-cell constant sum_g lgamma(1 + x), once per data set; 4 bytes per row on the wire
        from scipy.special import gammaln
        terms = gammaln(1.0 + data.astype(numpy.float64))
        csum_t = numpy.concatenate([[0.0], numpy.cumsum(terms)])
        self.row_const = torch.from_numpy(
            (csum_t[indptr[1:]] - csum_t[indptr[:-1]]).astype(numpy.float32)).pin_memory()
        n = shape[0]
        row_nnz = numpy.diff(indptr)
        # worst-case slab of max_rows consecutive rows
        csum = numpy.concatenate([[0], numpy.cumsum(row_nnz)])
        hi = numpy.minimum(numpy.arange(n) + max_rows, n)
        self.max_nnz = int((csum[hi] - csum[numpy.arange(n)]).max()) if n else 0
        self.max_rows = max_rows
        self.u16_ok = _counts_fit_u16(data)
        self.f16_exact = bool(self.u16_ok and (data.size == 0 or data.max() <= 2048))
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.slots = []
        for _ in range(2):
            self.slots.append({
                "indptr": torch.empty(max_rows + 1, dtype=torch.int64, device=self.device),
                "indices": torch.empty(max(self.max_nnz, 1), dtype=self.indices.dtype,
                                       device=self.device),
                "values": torch.empty(max(self.max_nnz, 1), dtype=self.values.dtype,
                                      device=self.device),
                "row_const": torch.empty(max_rows, dtype=torch.float32, device=self.device),
                "ready": torch.cuda.Event(), "free": torch.cuda.Event(), "bytes": 0,
                "u16_ok": self.u16_ok, "f16_exact": self.f16_exact,
            })
        for s in self.slots:
            s["free"].record()

    def fetch(self, slot_id, i0, i1):
        """Enqueue the H2D copy of rows [i0, i1) into staging slot ``slot_id``."""
        s = self.slots[slot_id]
        lo, hi = int(self.indptr[i0]), int(self.indptr[i1])
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(s["free"])
            s["indptr"][: i1 - i0 + 1].copy_(self.indptr[i0:i1 + 1], non_blocking=True)
            s["indices"][: hi - lo].copy_(self.indices[lo:hi], non_blocking=True)
            s["values"][: hi - lo].copy_(self.values[lo:hi], non_blocking=True)
            s["row_const"][: i1 - i0].copy_(self.row_const[i0:i1], non_blocking=True)
            s["ready"].record(self.copy_stream)
        s["bytes"] = (i1 - i0 + 1) * 8 + (hi - lo) * self.bytes_per_nnz + (i1 - i0) * 4
        return s


class PackedStream:
    """The count matrix stays in HOST memory; every step's rows travel as ONE packed slab (see
    ``scvae_csr_densify_packed``: ~2 bytes per non-zero, a single host -> device copy per step)
    into one of two device staging buffers on a copy stream while the previous step computes.

    Every row is encoded ONCE (block counts, (index, count) byte pairs, an escape list for counts
    >= 255).  ``pack_epoch(order)`` fixes the epoch's row order; a feeder thread then assembles
    slab after slab in a ring of pinned buffers (``scvae_pack_row_slab``: a gather of the rows'
    strings, GIL released) ahead of the GPU, and ``fetch(slot, k)`` enqueues the copy of slab k.
    For matrices beyond HBM, and the host-to-device leg of `bench.py`'s end-to-end number."""

    BLOCK = 255          # genes per block: a block's non-zero count fits one byte
    RING = 4             # pinned slab buffers the feeder thread may run ahead by

    def __init__(self, matrix, device, minibatch_size, pack_threads=None, feeder=None,
                 host_fraction=None):
        indptr, indices, data, shape = _as_csr_arrays(matrix)
        if not (_counts_fit_u16(data) and shape[1] <= 65280):
            raise ValueError("the packed stream carries integer counts <= 65504 of <= 65280 genes")
        import time
        t0 = time.perf_counter()
        self.shape = shape
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self._copy_stream = None
        self.B = int(minibatch_size)
        import os
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")) or 1)
        cores_per_rank = (os.cpu_count() or 8) // max(local_world, 1)
        if pack_threads is None:          # host threads of the slab gather, shared between the ranks of a node
            pack_threads = max(1, min(8, cores_per_rank - 2))     # (two cores stay with the rank's own threads)
        self.pack_threads = int(pack_threads)
        self.u16_ok = True
        self.f16_exact = bool(data.size == 0 or data.max() <= 2048)
        n, G = shape
        self.nblk = nblk = -(-G // self.BLOCK)
        from scipy.special import gammaln
        csum = numpy.concatenate([[0.0], numpy.cumsum(gammaln(1.0 + data.astype(numpy.float64)))])
        self.row_const = numpy.ascontiguousarray((csum[indptr[1:]] - csum[indptr[:-1]]).astype(numpy.float32))
        # ---- encode every row once --------------------------------------------------------------
        nnz = numpy.diff(indptr)
        if nnz.size and nnz.max() > 65535:
            raise ValueError("the packed stream carries at most 65535 non-zeros per row")
        total = int(indices.size)
        row_of = numpy.repeat(numpy.arange(n, dtype=numpy.int64), nnz)
        pos_in_row = numpy.arange(total, dtype=numpy.int64) - numpy.repeat(indptr[:-1], nnz)
        blk = indices // self.BLOCK
        big = data >= 255
        nesc = numpy.bincount(row_of[big], minlength=n).astype(numpy.int64)
        lens = (4 + nblk + 2 * nnz + 4 * nesc + 15) & ~15        # strings are 16-byte multiples
        self.row_off = numpy.concatenate([[0], numpy.cumsum(lens)]).astype(numpy.int64)
        store = self._host_bytes(int(self.row_off[-1]))
        base = self.row_off[:-1]
        store[base], store[base + 1] = nesc & 255, nesc >> 8
        store[base + 2], store[base + 3] = nnz & 255, nnz >> 8
        blocks = numpy.bincount(row_of * nblk + blk, minlength=n * nblk)
        store[(base[:, None] + 4 + numpy.arange(nblk)[None, :]).reshape(-1)] = blocks.astype(numpy.uint8)
        ent = numpy.repeat(base + 4 + nblk, nnz) + 2 * pos_in_row
        store[ent] = (indices - blk * self.BLOCK).astype(numpy.uint8)
        store[ent + 1] = numpy.minimum(data, 255).astype(numpy.uint8)
        if big.any():
            rows_big = row_of[big]
            first_big = numpy.concatenate([[0], numpy.cumsum(nesc)])[:-1]
            k_in_row = numpy.arange(rows_big.size) - numpy.repeat(first_big, nesc)
            esc = (base + 4 + nblk + 2 * nnz)[rows_big] + 4 * k_in_row
            p, v = pos_in_row[big], data[big].astype(numpy.int64)
            store[esc], store[esc + 1] = p & 255, p >> 8
            store[esc + 2], store[esc + 3] = v & 255, v >> 8
        self.store = store
        self.lens = lens
        self.bytes_per_nonzero = float(store.size) / max(total, 1)
        # feeder "host": a thread gathers each minibatch's strings into a pinned slab
        # (scvae_pack_row_slab, ~9 GB/s per host thread) that ONE copy ships -- the default when the
        # rank has host cores to spare (>= 6); "device": the strings sit in pinned memory and the GPU
        # pulls the rows itself (scvae_packed_pull: 45 GB/s alone, but its CTAs share the SMs with the
        # step they run beside, which costs ~0.2 ms per step) -- no host work per step, the default
        # when several ranks share few cores (8 GPUs on a 16-core host: 2.6 x the host feeder)
        # "batch": the strings sit in pinned memory and ONE batched copy (cudaMemcpyBatchAsync: the
        # copy engine walks the B row strings) moves them into the device slab -- measured 1.5 ms
        # of driver time per 4096-row batch, so the host gather stays the default
        # "hybrid": both at once -- the last `host_fraction` of every minibatch's rows goes through
        # the host gather + copy, the rest is pulled by the GPU, each into its own device slab (two
        # densify launches).  Measured on 4 / 8 GPUs sharing 16 cores: 17.1 / 24.2 M cells/s against
        # 19.7 / 28.8 M for the pure pull (the host threads are the scarcer resource), so not a default
        self.feeder = feeder or ("host" if (cores_per_rank >= 6 or not self.cuda) else "device")
        if self.feeder not in ("device", "host", "batch", "hybrid") or (self.feeder != "host" and not self.cuda):
            raise ValueError("feeder: `hybrid` / `device` / `batch` (CUDA only) or `host`")
        if host_fraction is None:
            host_fraction = min(1.0, cores_per_rank / 6.0)
        # rows [0, split) of a full minibatch are pulled by the GPU, rows [split, B) come from the host
        self.split = 0
        if self.feeder == "hybrid":
            self.split = min(self.B - 1, max(1, int(round(self.B * (1.0 - float(host_fraction))))))
        if self.feeder in ("device", "batch", "hybrid"):
            self.store_pinned = torch.from_numpy(store).pin_memory()
            self.store = self.store_pinned.numpy()
        if self.feeder in ("device", "hybrid"):
            self.row_off_dev = torch.from_numpy(self.row_off).to(self.device)
            self.row_const_dev = torch.from_numpy(self.row_const).to(self.device)
        if self.feeder == "batch":
            hb = K.packed_rows_offset(self.B)
            self._headers = [torch.empty(hb, dtype=torch.uint8).pin_memory() for _ in range(self.RING)]
            self._headers_np = [h.numpy() for h in self._headers]
            self._header_done = [None] * self.RING
            probe = torch.empty(K.packed_rows_offset(1) + int(lens.max() if n else 16) + 16,
                                dtype=torch.uint8, device=self.device)
            with torch.cuda.stream(self.copy_stream):
                if n == 0 or K.packed_copy_batch(self.store, self.row_off, self.row_const,
                                                 numpy.zeros(1, dtype=numpy.int64), self._headers_np[0],
                                                 probe) is None:
                    self.feeder = "host"          # no batched copies in this runtime
            torch.cuda.synchronize(self.device)
        self.encode_seconds = time.perf_counter() - t0
        # the largest slab any B rows can make: the B longest strings
        top = numpy.sort(lens)[::-1][:self.B]
        self.max_slab_bytes = (K.packed_rows_offset(self.B) + int(top.sum()) + 15) & ~15
        self._ring = self._ring_np = None
        self.slots = None
        self.slabs = []
        self._thread = None
        self._cv = None
        self.pack_seconds = 0.0

    def _host_bytes(self, size):
        """Zeroed host bytes for the row strings, in transparent huge pages when the kernel grants
        them (the per-step gather reads ~3 KB strings at random: 4 KB pages mean a TLB miss each)."""
        import mmap
        if size >= (8 << 20) and hasattr(mmap, "MADV_HUGEPAGE"):
            try:
                mem = mmap.mmap(-1, (size + (2 << 20) - 1) & ~((2 << 20) - 1))
                mem.madvise(mmap.MADV_HUGEPAGE)
                self._store_mmap = mem
                return numpy.frombuffer(mem, dtype=numpy.uint8, count=size)
            except (OSError, ValueError):
                pass
        return numpy.zeros(size, dtype=numpy.uint8)

    @property
    def copy_stream(self):
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        return self._copy_stream

    @property
    def number_of_examples(self):
        return self.shape[0]

    def _buffers(self):
        if self._ring is not None:
            return
        ring = [] if self.feeder in ("device", "batch") else [
            torch.empty(self.max_slab_bytes, dtype=torch.uint8) for _ in range(self.RING)]
        if self.cuda:
            ring = [r.pin_memory() for r in ring]
            self.slots = []
            for _ in range(2):
                slot = {"packed": True,
                        "buf": torch.empty(self.max_slab_bytes, dtype=torch.uint8, device=self.device),
                        "ready": torch.cuda.Event(), "free": torch.cuda.Event(), "bytes": 0, "rows": 0,
                        "u16_ok": True, "f16_exact": self.f16_exact, "stream": self, "split": 0}
                if self.feeder == "hybrid":      # second slab: the rows the host gathers
                    slot["buf_host"] = torch.empty(self.max_slab_bytes, dtype=torch.uint8,
                                                   device=self.device)
                slot["free"].record()
                self.slots.append(slot)
        self._ring, self._ring_np = ring, [r.numpy() for r in ring]

    def pack_slab_host(self, k, dst=None):
        """Assemble slab ``k`` of the current epoch (numpy uint8 view of its bytes)."""
        slab = self.slabs[k]
        if dst is None:
            dst = numpy.empty(self.max_slab_bytes, dtype=numpy.uint8)
        first = self._split_of(slab)        # (hybrid: the host packs the rows behind the split)
        nbytes = K.pack_row_slab(self.store, self.row_off, self.row_const,
                                 numpy.ascontiguousarray(slab["order"][first:]), dst,
                                 threads=self.pack_threads)
        slab["bytes_host"] = nbytes
        if first == 0:
            slab["bytes"] = nbytes
        return dst[:nbytes]

    def _split_of(self, slab):
        """Rows of this slab the GPU pulls (hybrid feeder; a ragged last slab keeps >= 1 host row)."""
        return max(0, min(self.split, slab["rows"] - 1)) if self.feeder == "hybrid" else 0

    def pack_epoch(self, order=None):
        """Fix the epoch's row order (default: data order) and start assembling its slabs of ``B``
        rows ahead of the GPU.  Returns the number of slabs."""
        import threading
        self.close()
        n, B = self.shape[0], self.B
        order = numpy.arange(n, dtype=numpy.int64) if order is None else \
            numpy.ascontiguousarray(order, dtype=numpy.int64)
        self.slabs = [{"rows": min(B, order.size - i), "bytes": 0,
                       "order": numpy.ascontiguousarray(order[i:i + B])}
                      for i in range(0, order.size, B)]
        self._buffers()
        if self.feeder in ("device", "hybrid"):
            self.order_dev = torch.from_numpy(order).to(self.device)      # the epoch's row order
        if self.feeder in ("device", "batch"):
            self._taken_upto = 0
            return len(self.slabs)
        self._cv = threading.Condition()
        self._packed_upto = 0                 # slabs [0, _packed_upto) sit in the ring
        self._taken_upto = 0                  # slabs [0, _taken_upto) have had their copy enqueued
        self._released_upto = 0               # slabs [0, _released_upto) have left their ring buffer
        self._copied = {}                     # slab -> event behind its copy out of the ring
        self._stop = False
        self._error = None
        self.pack_seconds = 0.0
        self._thread = threading.Thread(target=self._feed, daemon=True)
        self._thread.start()
        return len(self.slabs)

    def _feed(self):
        import time
        try:
            for k in range(len(self.slabs)):
                j = k % self.RING
                with self._cv:
                    # (no CUDA call in this thread -- it runs beside stream captures: the main
                    # thread reports which slabs have left the host, see fetch)
                    while not self._stop and k - self._released_upto >= self.RING:
                        self._cv.wait()
                    if self._stop:
                        return
                t0 = time.perf_counter()
                self.pack_slab_host(k, self._ring_np[j])
                self.pack_seconds += time.perf_counter() - t0
                with self._cv:
                    self._packed_upto = k + 1
                    self._cv.notify_all()
        except BaseException as exc:          # surfaced by the next fetch
            with self._cv:
                self._error = exc
                self._cv.notify_all()

    def close(self):
        """Stop the feeder thread of the current epoch (if any)."""
        if self._thread is not None:
            with self._cv:
                self._stop = True
                self._cv.notify_all()
            self._thread.join()
            self._thread = None

    def fetch(self, slot_id, k):
        """Enqueue the host -> device copy of slab ``k`` (slabs are taken in order) into staging
        slot ``slot_id``."""
        if self.feeder in ("device", "batch"):
            if k != self._taken_upto:
                raise ValueError("packed slabs are fetched in epoch order")
            self._taken_upto = k + 1
            slab, slot = self.slabs[k], self.slots[slot_id]
            rows = slab["rows"]
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(slot["free"])
                if self.feeder == "device":
                    K.packed_pull(self.store_pinned, self.row_off_dev, self.row_const_dev,
                                  self.order_dev[k * self.B:k * self.B + rows], slot["buf"])
                    slab["bytes"] = K.packed_rows_offset(rows) + int(self.lens[slab["order"]].sum())
                else:
                    j = k % self.RING
                    if self._header_done[j] is not None:
                        self._header_done[j].synchronize()     # (four steps old: already executed)
                    slab["bytes"] = K.packed_copy_batch(self.store, self.row_off, self.row_const,
                                                        slab["order"], self._headers_np[j], slot["buf"])
                    self._header_done[j] = torch.cuda.Event()
                    self._header_done[j].record(self.copy_stream)
                slot["ready"].record(self.copy_stream)
            slot["bytes"], slot["rows"] = slab["bytes"], rows
            return slot
        with self._cv:
            if k != self._taken_upto:
                raise ValueError("packed slabs are fetched in epoch order")
            while self._packed_upto <= k and self._error is None:
                self._cv.wait()
            if self._error is not None:
                raise self._error
        slab, j = self.slabs[k], k % self.RING
        if not self.cuda:
            out = {"packed": True, "host": self._ring_np[j][:slab["bytes"]].copy(), "rows": slab["rows"],
                   "bytes": slab["bytes"]}
            with self._cv:
                self._taken_upto = self._released_upto = k + 1
                self._cv.notify_all()
            return out
        slot = self.slots[slot_id]
        first = self._split_of(slab)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(slot["free"])
            if first:       # hybrid: the GPU pulls rows [0, first) itself, the host's rows follow by copy
                K.packed_pull(self.store_pinned, self.row_off_dev, self.row_const_dev,
                              self.order_dev[k * self.B:k * self.B + first], slot["buf"])
                slot["buf_host"][:slab["bytes_host"]].copy_(self._ring[j][:slab["bytes_host"]],
                                                            non_blocking=True)
                slab["bytes"] = slab["bytes_host"] + K.packed_rows_offset(first) + int(
                    self.lens[slab["order"][:first]].sum())
            else:
                slot["buf"][:slab["bytes"]].copy_(self._ring[j][:slab["bytes"]], non_blocking=True)
            slot["split"] = first
            slot["ready"].record(self.copy_stream)
            done = torch.cuda.Event()
            done.record(self.copy_stream)
        self._copied[k] = done
        # slabs whose copy was enqueued two fetches ago have long left the host: release their ring
        # buffers to the feeder (the wait below returns at once)
        released = self._released_upto
        while released <= k - 2:
            self._copied.pop(released).synchronize()
            released += 1
        with self._cv:
            self._released_upto = released
            self._taken_upto = k + 1
            self._cv.notify_all()
        slot["bytes"], slot["rows"] = slab["bytes"], slab["rows"]
        return slot


class ReconstructionSink:
    """Streams the (N, G) reconstruction of an evaluation pass to the host (VAE:1939-2052, where
    the reference concatenates ``p_x_mean`` of every ``session.run`` in host memory).

    The moments kernel of minibatch b writes ``p_x_mean`` straight into one of two device staging
    buffers (row pitch = the host row pitch, so the copy is ONE contiguous transfer); a copy
    stream moves it into its rows of the pinned (N, G) result while minibatch b + 1 computes.
    ``dtype`` float16 halves the transfer (an option the reference does not have; default fp32).
    """

    def __init__(self, number_of_rows, number_of_features, minibatch_size, device, dtype="float32"):
        self.n, self.G, self.B = int(number_of_rows), int(number_of_features), int(minibatch_size)
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        if dtype not in ("float32", "float16"):
            raise ValueError("reconstruction dtype: float32 or float16")
        self.half = dtype == "float16"
        # fp16 rows are written 8 columns per thread (scvae_f32_to_f16): pitch a multiple of 8
        self.ld = -(-self.G // 8) * 8 if self.half else self.G
        tdtype = torch.float16 if self.half else torch.float32
        self.host = torch.empty(self.n, self.ld, dtype=tdtype, pin_memory=self.cuda)
        self.mean32 = [torch.empty(self.B, self.ld, dtype=torch.float32, device=self.device)
                       for _ in range(2 if not self.half else 1)]
        self.stage16 = ([torch.empty(self.B, self.ld, dtype=tdtype, device=self.device)
                         for _ in range(2)] if self.half else None)
        self._copy_stream = torch.cuda.Stream(device=self.device) if self.cuda else None
        self._free = [None, None]
        self.bytes_copied = 0
        self.count = 0

    def mean_buffer(self):
        """Device buffer (B, ld) the moments kernel writes this minibatch's ``p_x_mean`` into."""
        slot = self.count % 2
        if self.cuda and not self.half and self._free[slot] is not None:
            torch.cuda.current_stream().wait_event(self._free[slot])   # its last copy has left
        return self.mean32[0 if self.half else slot]

    def push(self, start, rows):
        """Enqueue the transfer of the rows just written by the moments kernel."""
        slot = self.count % 2
        self.count += 1
        if self.half:
            if self.cuda and self._free[slot] is not None:
                torch.cuda.current_stream().wait_event(self._free[slot])
            K.f32_to_f16(self.mean32[0][:rows], self.ld, self.stage16[slot][:rows])
            src = self.stage16[slot]
        else:
            src = self.mean32[slot]
        dst = self.host[start:start + rows]
        self.bytes_copied += dst.numel() * dst.element_size()
        if not self.cuda:
            dst.copy_(src[:rows])
            return
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(ready)
            dst.copy_(src[:rows], non_blocking=True)
            if self._free[slot] is None:
                self._free[slot] = torch.cuda.Event()
            self._free[slot].record(self._copy_stream)

    def finish(self):
        """Wait for the last transfer; the (N, G) result as a numpy view of the pinned buffer."""
        if self.cuda:
            self._copy_stream.synchronize()
        out = self.host.numpy()
        return out if self.ld == self.G else out[:, :self.G]


class TrainLoop:
    """One (engine, minibatch size) training loop with a CUDA-graph-captured step."""

    def __init__(self, engine, minibatch_size, R=1, S=1, seed=1, use_graph=True):
        self.engine = engine
        self.B = int(minibatch_size)
        self.R, self.S = int(R), int(S)
        self.seed = int(seed)
        self.use_graph = bool(use_graph)
        self.plan = engine._plan(self.B, self.R * self.S)
        self.rows = torch.zeros(self.B, dtype=torch.int64, device=engine.device)
        self._graph = None
        self._graph_key = None
        self._src = None

    # the captured body: everything that is identical from step to step
    def _body(self, src, lr, w):
        eng, p = self.engine, self.plan
        if self.R == 1 and hasattr(eng, "fork_shadows"):
            eng.fork_shadows(p)
        if isinstance(src, ResidentCSR):
            eng.set_batch_csr(p, src.indptr, src.indices, src.values, self.rows,
                              u16_ok=src.u16_ok, f16_exact=src.f16_exact, train16=self.R == 1,
                              row_const_all=src.row_const)
        elif src.get("packed"):      # staging slot of a PackedStream: one slab = this minibatch
            split = src.get("split", 0)
            if split:                # hybrid feeder: pulled rows, then the host-gathered rows
                eng.set_batch_packed(p, src["buf"], f16_exact=src["f16_exact"], rows=(0, split))
                eng.set_batch_packed(p, src["buf_host"], f16_exact=src["f16_exact"], rows=(split, p.B))
            else:
                eng.set_batch_packed(p, src["buf"], f16_exact=src["f16_exact"])
        else:  # staging slot of a StreamedCSR
            eng.set_batch_csr(p, src["indptr"], src["indices"], src["values"], None, rebase=True,
                              u16_ok=src["u16_ok"], f16_exact=src["f16_exact"],
                              train16=self.R == 1, row_const_all=src["row_const"])
        targets = getattr(src, "targets", None) if isinstance(src, ResidentCSR) else None
        if targets is not None:
            # likelihood targets that differ from the network input (binarised values of the
            # Bernoulli likelihood, VAE:854-857): a second resident matrix, same rows
            if p.T is None:
                p.T = torch.zeros(p.B, eng.Gp, dtype=torch.float32, device=eng.device)
            K.csr_densify(targets.indptr, targets.indices, targets.values, self.rows, eng.G, p.T, None)
            p.use_T = True
        if getattr(eng, "constrained", False):
            if not isinstance(src, ResidentCSR) or src.count_sum_parameter is None:
                raise NotImplementedError("the constrained Poisson needs a resident data set with "
                                          "count sums (set_features(count_sum_parameter=...))")
            K.gather_f32(src.count_sum_parameter, self.rows, p.count_sum_parameter)
        if getattr(eng, "n_extra", 0):
            # decoder-input extras (batch correction / count-sum feature) of this minibatch
            if not isinstance(src, ResidentCSR):
                raise NotImplementedError("batch correction / count-sum features need a resident "
                                          "data set")
            if eng.number_of_batches:
                K.gather_f32(src.batch_index, self.rows, p.batch_index)
            if eng.count_sum_feature:
                K.gather_f32(src.count_sum_feature, self.rows, p.count_sum)
        if self._device_scalars(src):
            p.eps_source = (self.seed, eng.store.step)      # drawn inside vae_mid_fwd
        else:
            p.eps_source = None
            K.fill_normal(p.eps, self.seed, 0, eng.store.step)
        eng.train_step(p, self.R, self.S, lr, w)
        p.eps_source = None

    def _device_scalars(self, src):
        """The step reads learning rate, warm-up weight and noise offset on the device."""
        eng = self.engine
        u16_ok = src.u16_ok if isinstance(src, ResidentCSR) else src["u16_ok"]
        fn = getattr(eng, "step_on_device_scalars", None)
        return bool(fn is not None and fn(self.plan, self.R, self.S, u16_ok))

    def step(self, src, lr, warm_up_weight=1.0):
        """Run one optimiser step on the minibatch described by ``src`` (+ ``self.rows``)."""
        eng = self.engine
        if hasattr(eng, "set_step_scalars"):
            eng.set_step_scalars(lr, warm_up_weight)       # read on the device by the step's kernels
        # the learning rate always comes from the device; the warm-up weight too on the fused path,
        # so one graph per minibatch source serves a whole warm-up schedule
        key = (id(src), None if self._device_scalars(src) else float(warm_up_weight))
        if not self.use_graph:
            self._body(src, lr, warm_up_weight)
            return self.plan.bound
        if self._graph is None or self._graph_key != key:
            if self._graph is None or self._graph_key is None:
                pass
            graphs = getattr(self, "_graphs", None)
            if graphs is None:
                graphs = self._graphs = {}
            if key not in graphs:
                # eager warm-up allocates every lazily created buffer before capture; the
                # optimiser state is snapshotted so that warm-up + capture are side-effect free
                eng = self.engine
                snap = (eng.store.param.clone(), eng.store.m.clone(), eng.store.v.clone(),
                        eng.store.step.clone(),
                        [(l.moving_mean.clone(), l.moving_var.clone())
                         for l in eng.enc + eng.dec if l.bn])
                self._body(src, lr, warm_up_weight)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._body(src, lr, warm_up_weight)
                if hasattr(eng, "invalidate_shadows"):
                    eng.invalidate_shadows()         # the snapshot restores the masters only
                eng.store.param.copy_(snap[0])
                eng.store.m.copy_(snap[1])
                eng.store.v.copy_(snap[2])
                eng.store.step.copy_(snap[3])
                for l, (mm, mv) in zip([l for l in eng.enc + eng.dec if l.bn], snap[4]):
                    l.moving_mean.copy_(mm)
                    l.moving_var.copy_(mv)
                graphs[key] = (g, src)   # keep src alive: its buffers are baked into the graph
            self._graph, self._graph_key = graphs[key][0], key
        if getattr(eng, "_shadow_valid", True) is False and getattr(eng, "W16", None) is not None \
                and getattr(eng, "world_size", 1) == 1:
            # the captured step holds no shadow refresh (the optimiser kernel keeps the fp16 weight
            # copies current): bring them up to date after an outside change of the parameters
            eng._refresh_shadows(self.plan, self.plan.M)
        self._graph.replay()
        return self.plan.bound
