"""Diagnostic: first Adam step from zero moments must move every parameter with a non-negligible
gradient by ~lr; reports elements that moved by something else (packed / streamed / resident)."""
import sys
import numpy, torch, scipy.sparse
sys.path.insert(0, ".")
from scvae_b200.engine import VAEEngine
from scvae_b200.hotloop import ResidentCSR, StreamedCSR, PackedStream, TrainLoop

def run(source, B, G, L, lik, trial):
    dev = torch.device("cuda:0")
    rng = numpy.random.RandomState(23 + trial)
    N = 2 * B
    x = ((rng.rand(N, G) < 0.07) * rng.randint(1, 500, size=(N, G))).astype(numpy.float32)
    csr = scipy.sparse.csr_matrix(x)
    eng = VAEEngine(G, L, [100], lik, device=dev, seed=4)
    loop = TrainLoop(eng, B, seed=31, use_graph=True)
    if source == "resident":
        src = ResidentCSR(csr, dev)
        loop.rows.copy_(torch.from_numpy(rng.permutation(N)[:B].astype(numpy.int64)).to(dev))
    elif source == "packed":
        st = PackedStream(csr, dev, B); st.pack_epoch(rng.permutation(N)); src = st.fetch(1, 1)
        torch.cuda.current_stream().wait_event(src["ready"])
    else:
        st = StreamedCSR(csr, dev, B); src = st.fetch(0, B // 2, B // 2 + B)
        torch.cuda.current_stream().wait_event(src["ready"])
    before = eng.store.param.clone()
    lr = 1e-3
    loop.step(src, lr, 0.8)
    torch.cuda.synchronize()
    g = eng.store.grad.clone()
    d = (eng.store.param - before).abs()
    sel = g.abs() > 1e-5
    bad = sel & ((d - lr).abs() > 0.02 * lr)
    print(source, "trial", trial, "step", eng.store.step.item(), "counter", eng._adam_counter.item(),
          "bad", int(bad.sum()), "of", int(sel.sum()),
          "ratios", (d[bad][:5] / lr).tolist() if bad.any() else [],
          "first bad index", int(bad.nonzero()[0]) if bad.any() else -1, "total", eng.store.total)
    loop.step(src, lr, 0.8); torch.cuda.synchronize()
    print("   after 2nd step: step", eng.store.step.item(), "counter", eng._adam_counter.item())

for trial in range(3):
    for source in ("packed", "streamed", "resident"):
        run(source, 512, 28000, 100, "zero-inflated negative binomial", trial)
