"""Development aid: host gather rate of hotloop.PackedStream (scvae_pack_row_slab) by thread count."""
import sys, time
import numpy, torch
sys.path.insert(0, ".")
import bench
from scvae_b200.hotloop import PackedStream
csr = bench.make_csr(68000, 20000, 0.07, seed=60, device=torch.device("cuda:0"))
for threads in (1, 2, 4, 8, 12):
    st = PackedStream(csr, "cuda:0", 4096, pack_threads=threads, feeder="host")
    st._buffers()
    order = numpy.random.RandomState(1).permutation(68000)
    st.slabs = [{"rows": 4096, "bytes": 0, "order": numpy.ascontiguousarray(order[i:i + 4096])} for i in range(0, 65536, 4096)]
    for rep in range(2):
        t0 = time.perf_counter()
        for k in range(16):
            st.pack_slab_host(k, st._ring_np[k % 4])
        dt = (time.perf_counter() - t0) / 16
    print("threads %2d: %.3f ms per 4096-row slab (%.1f MB) -> %.1f GB/s" % (
        threads, dt * 1e3, st.slabs[0]["bytes"] / 1e6, st.slabs[0]["bytes"] / dt / 1e9), flush=True)
