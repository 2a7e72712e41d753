"""Development aid: phase timeline (ns, %globaltimer) of vae_mid_fwd / vae_mid_bwd at the C2 shape."""
import os, sys
if "--no-timeline" not in sys.argv:
    os.environ["SCVAE_MID_TIMELINE"] = "1"
import numpy, torch, scipy.sparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scvae_b200.engine import VAEEngine
from scvae_b200.hotloop import ResidentCSR, TrainLoop
B, G, L = 4096, 20000, 50
for a in sys.argv[1:]:
    if a.isdigit():
        B = int(a)
dev = torch.device("cuda:0")
rng = numpy.random.RandomState(1)
x = ((rng.rand(2 * B, G) < 0.07) * numpy.floor(1 - numpy.log(rng.rand(2 * B, G)) * 1.2)).astype(numpy.float32)
data = ResidentCSR(scipy.sparse.csr_matrix(x), dev)
eng = VAEEngine(G, L, [100], "negative binomial", device=dev, seed=0)
loop = TrainLoop(eng, B, seed=1, use_graph=False)
for it in range(300):          # long enough for the clocks to ramp up
    loop.rows.copy_(torch.arange(B, device=dev) + (it % 2) * B)
    loop.step(data, 1e-4, 1.0)
torch.cuda.synchronize()
if "--no-timeline" in sys.argv:
    sys.exit(0)
for key in ("_mid_fwd", "_mid_bwd"):
    tl = getattr(loop.plan, key + "_timeline").cpu().numpy().reshape(160, 32)
    grid = int((tl[:, 0] > 0).sum())
    t = tl[:grid].astype(numpy.float64)
    t0 = t[:, 0].min()
    print(key, "grid", grid)
    for k in range(15):
        col = t[:, k]
        if (col > 0).all():
            print("  stamp %2d: first CTA %8.2f us  median %8.2f  last %8.2f" % (
                k, (col.min() - t0) / 1e3, (numpy.median(col) - t0) / 1e3, (col.max() - t0) / 1e3))
