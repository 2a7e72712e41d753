"""Development aid: lean evaluation pass through vae_mid_fwd (moving statistics) vs the oracle, layer by layer."""
import os, sys
import numpy, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import scvae_oracle as O
from scvae_b200.engine import VAEEngine
G, L, hidden, lik, B = 512, 7, [48, 24], "zero-inflated negative binomial", 200
cfg = O.VAEConfig(G, L, hidden, lik, "gaussian", 1, 1, True, True, kl_weight=0.7)
params = O.vae_init_params(cfg, seed=3, dtype=torch.float64)
gen = torch.Generator().manual_seed(11)
for k in params:
    if k.endswith("biases") or k.endswith("beta"):
        params[k] = torch.randn(params[k].shape, generator=gen, dtype=torch.float64) * 0.1
x, _ = O.synthetic_counts(B, G, n_types=3, seed=5, target_zero_fraction=0.8)
x = numpy.minimum(x, 500.0)
eps = torch.randn(1, B, L, generator=gen, dtype=torch.float64)
x64 = torch.tensor(x, dtype=torch.float64)
upd = []
O.vae_forward(cfg, params, x64, x64, eps, True, bn_updates=upd)
for scope, mean, var in upd:
    params[scope + "/BATCH_NORM/moving_mean"] = mean[0] * 0.9
    params[scope + "/BATCH_NORM/moving_variance"] = var[0] * 1.1
for mid in ("1", "0"):
    os.environ["SCVAE_MID_FUSED"] = mid
    eng = VAEEngine(G, L, hidden, lik, "gaussian", True, kl_weight=0.7, device="cuda:0", tensor_cores=True)
    eng.import_parameters(params)
    plan = eng._plan(B, 1)
    eng.set_batch_dense(plan, torch.tensor(x).cuda())
    plan.eps.copy_(eps.reshape(B, L).float())
    ev = O.vae_forward(cfg, params, x64, x64, eps, is_training=False)
    eng.forward(plan, False, 1, 1, 1.0, keep_heads=False)
    torch.cuda.synchronize()
    b = plan.bound.cpu().numpy()
    print("mid", mid, "bound", b, "ref", ev["lower_bound"].item(), ev["reconstruction_error"].item(), ev["kl_divergence"].item())
    mu = plan.PH[:, :L].cpu().double()
    print("   mu rel", ((mu - ev["q_z_mean"]).abs().max() / ev["q_z_mean"].abs().max()).item(),
          "log_sigma abs", (plan.PH[:, L:2 * L].cpu().double().clamp(-3, 3) - ev["log_sigma"]).abs().max().item(),
          "kl_row rel", ((plan.kl_row.cpu().double() - ev["kl"].reshape(-1)).abs().max() / ev["kl"].abs().max()).item())
    # training-mode forward for comparison
    tr = O.vae_forward(cfg, params, x64, x64, eps, is_training=True)
    eng.forward(plan, True, 1, 1, 1.0, keep_heads=False, update_moving=False)
    torch.cuda.synchronize()
    mu = plan.PH[:, :L].cpu().double()
    print("   train-mode mu rel", ((mu - tr["q_z_mean"]).abs().max() / tr["q_z_mean"].abs().max()).item(), "kl", plan.bound.cpu().numpy()[3], tr["kl_divergence"].item())
