"""GPU parity of the GMVAE step engine against the CPU oracle (forward, gradients, Adam step,
evaluate moments), including cluster-chunked decoder passes."""
import numpy
import pytest
import torch

from oracle import scvae_oracle as O

pytestmark = pytest.mark.gpu

CASES = [
    # name, G, L, K, hidden, likelihood, R, S, B, prior, free_nats, head_buffer_bytes
    ("nb-uniform", 96, 5, 4, [24], "negative binomial", 1, 1, 40, "uniform", 0.0, 4 << 30),
    ("zinb-chunked", 64, 4, 5, [20, 12], "zero-inflated negative binomial", 1, 2, 24, "uniform", 0.0, 60000),
    ("poisson-learn", 80, 3, 3, [16], "poisson", 2, 1, 32, "learn", 0.0, 4 << 30),
    ("nb-freenats", 72, 4, 6, [16], "negative binomial", 1, 1, 30, "uniform", 0.9, 4 << 30),
    # minibatches of 128 rows: the tensor-core mode runs the cluster passes through the fused heads
    ("nb-fused", 96, 5, 3, [24], "negative binomial", 1, 1, 128, "uniform", 0.0, 4 << 30),
    ("zinb-fused-chunked", 64, 4, 4, [20], "zero-inflated negative binomial", 1, 2, 128, "uniform", 0.0,
     800000),
]


def _setup(case, tensor_cores):
    from scvae_b200.gmvae_engine import GMVAEEngine
    name, G, L, Kc, hidden, lik, R, S, B, prior, free_nats, hbb = case
    cfg = O.GMVAEConfig(G, L, Kc, hidden, lik, R, S, True, kl_weight=0.8,
                        prior_probabilities_method=prior,
                        proportion_of_free_nats_for_y_kl_divergence=free_nats)
    params = O.gmvae_init_params(cfg, seed=2, dtype=torch.float64)
    gen = torch.Generator().manual_seed(5)
    for k in params:
        if k.endswith("biases") or k.endswith("beta") or k.endswith("LOGITS"):
            params[k] = torch.randn(params[k].shape, generator=gen, dtype=torch.float64) * 0.2
    x, _ = O.synthetic_counts(B, G, n_types=3, seed=8, target_zero_fraction=0.8)
    x = numpy.minimum(x, 300.0)
    eps = torch.randn(Kc, R * S, B, L, generator=gen, dtype=torch.float64)
    eng = GMVAEEngine(G, L, Kc, hidden, lik, True, 0.8, prior, None, free_nats, device="cuda:0",
                      tensor_cores=tensor_cores, head_buffer_bytes=hbb)
    eng.import_parameters(params)
    plan = eng._plan(B, R * S)
    eng.set_batch_dense(plan, torch.tensor(x).cuda())
    plan.eps.copy_(eps.reshape(-1, L).float())
    return cfg, params, torch.tensor(x, dtype=torch.float64), eps, eng, plan


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("tensor_cores", [False, True], ids=["fp32", "tf32"])
def test_gmvae_forward_backward_step(case, tensor_cores):
    name, G, L, Kc, hidden, lik, R, S, B, prior, free_nats, hbb = case
    cfg, params, x, eps, eng, plan = _setup(case, tensor_cores)
    if "chunked" in name:
        assert plan.chunk < Kc
    w = 0.7
    tol = 5e-5 if not tensor_cores else 2e-3
    etol = 5e-5 if not tensor_cores else 1e-3      # ELBO terms: the north-star bound
    state = O.AdamState(params)
    ref = {k: v.clone() for k, v in params.items()}
    out, grads = O.train_step(cfg, ref, state, x, x, eps, 1e-3, warm_up_weight=w)
    bound = eng.train_step(plan, R, S, 1e-3, warm_up_weight=w).cpu().numpy()
    torch.cuda.synchronize()
    assert plan.fused_done == (tensor_cores and "fused" in name)
    names = ["lower_bound", "lower_bound_weighted", "reconstruction_error", "kl_divergence_z",
             "kl_divergence_y"]
    for i, n in enumerate(names):
        assert abs(bound[i] - out[n].item()) <= etol * abs(out[n].item()) + 1e-5, (n, bound[i], out[n].item())
    assert (plan.logits[:, :Kc].cpu().double() - out["q_y_logits"]).abs().max().item() <= \
        tol * out["q_y_logits"].abs().max().item() + 1e-5
    lp_ref = out["log_p_x_given_z"].reshape(-1)
    assert ((plan.logp.cpu().double() - lp_ref).abs().max() / lp_ref.abs().max()).item() <= tol
    klz_ref = out["kl_z"].reshape(-1)
    assert ((plan.klz.cpu().double() - klz_ref).abs().max() / klz_ref.abs().max()).item() <= tol
    got = eng.export_gradients()
    gtol = 3e-4 if not tensor_cores else 1e-2
    gmax = max(g.abs().max().item() for g in grads.values())
    for k, g in grads.items():
        err = (got[k].double() - g).abs().max().item()
        # (absolute floor for the exactly-zero gradients of batch-normed biases: fp32 noise, or
        # the fp16 operand rounding of the fused heads path)
        floor = (1e-4 if plan.fused_done else 1e-5) * gmax
        assert err <= gtol * g.abs().max().item() + floor, (k, err, g.abs().max().item())
    new = eng.export_parameters()
    noise = (1e-3 if not tensor_cores else 3e-2) * gmax
    for k, v in ref.items():
        diff = (new[k].double() - v).abs()
        if k in grads:
            diff = diff * (grads[k].abs() > noise)
        rtol = 5e-5 if "moving" in k else 1e-5   # fp32 batch variance of raw-count layers
        assert diff.max().item() <= rtol * max(v.abs().max().item(), 1.0), (k, diff.max().item())


def test_gmvae_evaluate_mode():
    case = CASES[0]
    name, G, L, Kc, hidden, lik, R, S, B, prior, free_nats, hbb = case
    cfg, params, x, eps, eng, plan = _setup(case, False)
    upd = []
    O.gmvae_forward(cfg, params, x, x, eps, True, bn_updates=upd)
    for scope, mean, var in upd:            # plausible moving statistics
        params[scope + "/BATCH_NORM/moving_mean"] = mean[0] * 0.9
        params[scope + "/BATCH_NORM/moving_variance"] = var[0] * 1.1
    eng.import_parameters(params)
    out = O.gmvae_forward(cfg, params, x, x, eps, is_training=False, moments=True)
    eng.forward(plan, False, R, S, 1.0)
    m = eng.moments(plan, R, S)
    zm = eng.z_mean(plan)
    torch.cuda.synchronize()
    b = plan.bound.cpu().numpy()
    assert abs(b[0] - out["lower_bound"].item()) <= 5e-5 * abs(out["lower_bound"].item())
    assert (zm.cpu().double() - out["z_mean"]).abs().max().item() <= 5e-5 * out["z_mean"].abs().max().item() + 1e-6
    scale = out["p_x_mean"].abs().max().item()
    assert (m[0].cpu().double() - out["p_x_mean"]).abs().max().item() <= 1e-4 * scale
    assert (m[1].cpu().double() - out["p_x_stddev"]).abs().max().item() <= 1e-4 * out["p_x_stddev"].abs().max().item()
    assert (m[2].cpu().double() - out["stddev_of_p_x_given_z_mean"]).abs().max().item() <= 1e-4 * scale


@pytest.mark.parametrize("tensor_cores", [False, True], ids=["fp32", "tc"])
@pytest.mark.parametrize("B", [40, 128])
def test_gmvae_batch_correction_and_count_sum_feature(tensor_cores, B):
    """Decoder-input extras concatenated to every z_k (GMVAE:3097-3132): forward, gradients and
    one optimiser step against the oracle (B = 128 also takes the fused heads path)."""
    from scvae_b200.gmvae_engine import GMVAEEngine
    G, L, Kc, hidden, lik = 96, 4, 3, [24], "negative binomial"
    opts = dict(number_of_batches=3, count_sum_feature=True)
    cfg = O.GMVAEConfig(G, L, Kc, hidden, lik, 1, 1, True, kl_weight=1.0, **opts)
    params = O.gmvae_init_params(cfg, seed=4, dtype=torch.float64)
    gen = torch.Generator().manual_seed(7)
    for k in params:
        if k.endswith("biases") or k.endswith("beta"):
            params[k] = torch.randn(params[k].shape, generator=gen, dtype=torch.float64) * 0.2
    x, _ = O.synthetic_counts(B, G, n_types=3, seed=9, target_zero_fraction=0.8)
    x = numpy.minimum(x, 100.0)
    x64 = torch.tensor(x, dtype=torch.float64)
    eps = torch.randn(Kc, 1, B, L, generator=gen, dtype=torch.float64)
    batch = torch.randint(0, 3, (B, 1), generator=gen)
    cs = x64.sum(dim=1)
    cs = ((cs - cs.min()) / (cs.max() - cs.min())).reshape(B, 1)
    feats = dict(batch_indices=batch, count_sum_feature=cs)
    state = O.AdamState(params)
    ref = {k: v.clone() for k, v in params.items()}
    out, grads = O.train_step(cfg, ref, state, x64, x64, eps, 1e-3, **feats)
    eng = GMVAEEngine(G, L, Kc, hidden, lik, True, 1.0, "uniform", None, 0.0, device="cuda:0",
                      tensor_cores=tensor_cores, **opts)
    eng.import_parameters(params)
    plan = eng._plan(B, 1)
    eng.set_batch_dense(plan, torch.tensor(x).cuda())
    eng.set_batch_features(plan, batch.cuda(), cs.float().cuda())
    plan.eps.copy_(eps.reshape(-1, L).float())
    bound = eng.train_step(plan, 1, 1, 1e-3).cpu().numpy()
    torch.cuda.synchronize()
    assert plan.fused_done == (tensor_cores and B % 128 == 0)
    tol = 5e-5 if not tensor_cores else 1e-3
    for i, n in enumerate(["lower_bound", "lower_bound_weighted", "reconstruction_error"]):
        assert abs(bound[i] - out[n].item()) <= tol * abs(out[n].item()) + 1e-5, (n, bound[i], out[n].item())
    got = eng.export_gradients()
    gtol = 3e-4 if not tensor_cores else 1e-2
    gmax = max(g.abs().max().item() for g in grads.values())
    floor = (1e-4 if plan.fused_done else 1e-5) * gmax
    for k, g in grads.items():
        assert got[k].shape == g.shape, (k, got[k].shape, g.shape)
        err = (got[k].double() - g).abs().max().item()
        assert err <= gtol * g.abs().max().item() + floor, (k, err, g.abs().max().item())


@pytest.mark.parametrize("lik", ["negative binomial", "zero-inflated poisson"])
def test_gmvae_piecewise_categorical(lik):
    """`-k` for the GMVAE (head P_K, GMVAE:3192-3220): step parity and the y-marginalised moments."""
    from scvae_b200.gmvae_engine import GMVAEEngine
    G, L, Kc, hidden, B, k_max = 80, 4, 3, [20], 36, 2
    cfg = O.GMVAEConfig(G, L, Kc, hidden, lik, 1, 2, True, kl_weight=1.0,
                        number_of_reconstruction_classes=k_max)
    params = O.gmvae_init_params(cfg, seed=5, dtype=torch.float64)
    gen = torch.Generator().manual_seed(8)
    for k in params:
        if k.endswith("biases") or k.endswith("beta"):
            params[k] = torch.randn(params[k].shape, generator=gen, dtype=torch.float64) * 0.2
    x, _ = O.synthetic_counts(B, G, n_types=3, seed=10, target_zero_fraction=0.7)
    x = numpy.minimum(x, 60.0)
    x64 = torch.tensor(x, dtype=torch.float64)
    eps = torch.randn(Kc, 2, B, L, generator=gen, dtype=torch.float64)
    state = O.AdamState(params)
    ref = {k: v.clone() for k, v in params.items()}
    out_m = O.gmvae_forward(cfg, params, x64, x64, eps, is_training=True, moments=True)
    out, grads = O.train_step(cfg, ref, state, x64, x64, eps, 1e-3)
    eng = GMVAEEngine(G, L, Kc, hidden, lik, True, 1.0, "uniform", None, 0.0, device="cuda:0",
                      tensor_cores=False, number_of_reconstruction_classes=k_max)
    eng.import_parameters(params)
    plan = eng._plan(B, 2)
    eng.set_batch_dense(plan, torch.tensor(x).cuda())
    plan.eps.copy_(eps.reshape(-1, L).float())
    eng.forward(plan, True, 1, 2, 1.0, update_moving=False)
    m = eng.moments(plan, 1, 2)
    torch.cuda.synchronize()
    scale = out_m["p_x_mean"].abs().max().item()
    assert (m[0].cpu().double() - out_m["p_x_mean"]).abs().max().item() <= 1e-4 * scale
    sd_scale = out_m["p_x_stddev"].abs().max().item()
    assert (m[1].cpu().double() - out_m["p_x_stddev"]).abs().max().item() <= 2e-4 * sd_scale
    bound = eng.train_step(plan, 1, 2, 1e-3).cpu().numpy()
    torch.cuda.synchronize()
    for i, n in enumerate(["lower_bound", "lower_bound_weighted", "reconstruction_error"]):
        assert abs(bound[i] - out[n].item()) <= 5e-5 * abs(out[n].item()) + 1e-5, (n, bound[i], out[n].item())
    got = eng.export_gradients()
    gmax = max(g.abs().max().item() for g in grads.values())
    for k, g in grads.items():
        assert got[k].shape == g.shape, (k, got[k].shape, g.shape)
        err = (got[k].double() - g).abs().max().item()
        assert err <= 3e-4 * g.abs().max().item() + 1e-5 * gmax, (k, err, g.abs().max().item())



@pytest.mark.parametrize("K_,B,L,RS", [(3, 5, 4, 2), (4, 37, 50, 1), (2, 8, 33, 3)])
def test_full_covariance_latent_kernels(K_, B, L, RS):
    """csrc/gmvae_full.cu (multivariate-Gaussian q(z|x,y) / p(z|y), f4) against the autograd
    restatement in tests/kernel_standins.py (fp64; fill_triangular, sampled KL with a triangular
    solve, gradients w.r.t. the head pre-activations of both distributions, covariance means)."""
    import kernel_standins as C
    from scvae_b200 import kernels as K
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(K_ * 100 + L)
    T = L * (L + 1) // 2
    M = K_ * RS * B
    ld = (L + T + 3) & ~3
    qh = torch.zeros(K_ * B, ld)
    qh[:, :L + T] = torch.randn(K_ * B, L + T, generator=gen) * 0.7
    pz = torch.zeros(K_, ld)
    pz[:, :L + T] = torch.randn(K_, L + T, generator=gen) * 0.7
    eps = torch.randn(M, L, generator=gen)
    dz = torch.zeros(M, (L + 4) & ~3)
    dz[:, :L] = torch.randn(M, L, generator=gen) * 0.1
    coef = torch.rand(M, generator=gen) * 0.05
    Zp = (L + 4) & ~3

    def run(mod, to):
        pl = to(torch.zeros(K_, L, L))
        z, klz, w, cu = to(torch.zeros(M, Zp)), to(torch.zeros(M)), to(torch.zeros(M, L)), to(torch.zeros(M, L))
        dqh, dpz, cov = to(torch.zeros(K_ * B, ld)), to(torch.zeros(K_, ld)), to(torch.zeros(K_, L, L))
        mod.gmvae_full_prior(to(pz), K_, L, pl)
        mod.gmvae_latent_full_fwd(to(qh), to(pz), pl, K_, B, L, RS, to(eps), z, klz, w)
        mod.gmvae_latent_full_bwd(to(qh), to(pz), pl, K_, B, L, RS, to(eps), to(dz), to(coef), w, cu, dqh, dpz)
        mod.gmvae_full_covariance_mean(to(qh), K_, B, L, cov)
        return [t.cpu().double() for t in (z, klz, w, dqh, dpz, cov, pl)]

    got = run(K, lambda t: t.to(dev))
    torch.cuda.synchronize()
    want = run(C, lambda t: t.clone())
    for name, g, r in zip(("z", "klz", "w", "dqh", "dpz", "cov", "pl"), got, want):
        err = (g - r).abs().max().item()
        assert err <= 2e-4 * max(r.abs().max().item(), 1.0), (name, err, r.abs().max().item())
