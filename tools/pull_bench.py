"""Development aid: PCIe rate of scvae_packed_pull (the GPU reading row strings from pinned host memory),
alone and beside a running training step."""
import sys, time
import numpy, torch
sys.path.insert(0, ".")
import bench
from scvae_b200 import kernels as K
from scvae_b200.hotloop import PackedStream
dev = torch.device("cuda:0")
csr = bench.make_csr(68000, 20000, 0.07, seed=60, device=dev)
st = PackedStream(csr, dev, 4096, feeder="device")
order = numpy.random.RandomState(1).permutation(68000)
st.pack_epoch(order[:65536])
for prio in (0, -1):
    stream = torch.cuda.Stream(device=dev, priority=prio)
    with torch.cuda.stream(stream):
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for k in range(16):
                K.packed_pull(st.store_pinned, st.row_off_dev, st.row_const_dev,
                              st.order_dev[k * 4096:(k + 1) * 4096], st.slots[k % 2]["buf"])
            e1.record(stream)
            stream.synchronize()
            ms = e0.elapsed_time(e1) / 16
        nbytes = K.packed_rows_offset(4096) + int(st.lens[order[:4096]].sum())
        print("priority %d: %.3f ms per slab of %.1f MB -> %.1f GB/s (alone)" % (prio, ms, nbytes / 1e6, nbytes / ms / 1e6), flush=True)
