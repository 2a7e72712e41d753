"""Development aid: one training step at a given shape, every bound term / gradient against the oracle."""
import os, sys
import numpy, torch, scipy.sparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import scvae_oracle as O
from scvae_b200.engine import VAEEngine
from scvae_b200.hotloop import ResidentCSR, TrainLoop

B, G, L = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
hidden = [int(h) for h in sys.argv[4].split(",")]
lik = sys.argv[5] if len(sys.argv) > 5 else "negative binomial"
dev = torch.device("cuda:0")
rng = numpy.random.RandomState(23)
mask = rng.rand(2 * B, G) < (rng.rand(G) * 0.14)
x_all = numpy.minimum(mask * numpy.floor(1.0 - numpy.log(rng.rand(2 * B, G)) * 1.2), 500).astype(numpy.float32)
csr = scipy.sparse.csr_matrix(x_all)
eng = VAEEngine(G, L, hidden, lik, device=dev, seed=4)
gen = torch.Generator().manual_seed(9)
params = eng.export_parameters()
for k in params:
    if k.endswith("biases") or k.endswith("beta"):
        params[k] = torch.randn(params[k].shape, generator=gen) * 0.1
eng.import_parameters(params)
loop = TrainLoop(eng, B, seed=31, use_graph=False)
data = ResidentCSR(csr, dev)
rows = torch.from_numpy(rng.permutation(2 * B)[:B].astype(numpy.int64)).to(dev)
loop.rows.copy_(rows)
before = {k: v.double() for k, v in eng.export_parameters().items()}
bound = loop.step(data, 1e-3, 0.8).cpu().numpy()
torch.cuda.synchronize()
plan = loop.plan
print("fused", plan.fused_done, "mid", getattr(plan, "mid_done", None), "mid_err", eng.mid_error(plan))
eps = plan.eps.cpu().double().reshape(1, B, L)
x = torch.tensor(x_all[rows.cpu().numpy()], dtype=torch.float64)
cfg = O.VAEConfig(G, L, hidden, lik)
state = O.AdamState(before)
ref = {k: v.clone() for k, v in before.items()}
out, grads = O.train_step(cfg, ref, state, x, x, eps, 1e-3, warm_up_weight=0.8)
for i, key in enumerate(["lower_bound", "lower_bound_weighted", "reconstruction_error", "kl_divergence"]):
    r = out[key].item()
    print("%-24s gpu %.6f ref %.6f rel %.2e" % (key, bound[i], r, abs(bound[i] - r) / abs(r)))
def rel(a, b):
    return ((a.double() - b).abs().max() / b.abs().max()).item()
print("q_z_mean rel", rel(plan.PH[:, :L].cpu(), out["q_z_mean"]), " logp rel", rel(plan.logp.cpu(), out["log_p_x_given_z"].reshape(-1)))
got = eng.export_gradients()
for k, g in grads.items():
    e = (got[k].double() - g).abs()
    print("%-40s err %.3e  max|g| %.3e  rel %.2e  argmax %s" % (k, e.max().item(), g.abs().max().item(), e.max().item() / (g.abs().max().item() + 1e-30), numpy.unravel_index(e.argmax().item(), tuple(e.shape))))
