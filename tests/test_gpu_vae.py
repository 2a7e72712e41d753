"""GPU parity of the VAE step engine (forward, gradients, clip+Adam step, evaluate moments)
against the CPU oracle on identical weights, inputs and noise."""
import numpy
import pytest
import torch

from oracle import scvae_oracle as O

pytestmark = pytest.mark.gpu

CASES = [
    # name, G, L, hidden, likelihood, R, S, B, bn, latent
    ("c1-poisson", 100, 10, [100], "poisson", 1, 1, 100, True, "gaussian"),
    ("nb-two-layers", 212, 7, [48, 24], "negative binomial", 1, 1, 64, True, "gaussian"),
    ("zinb-iw", 96, 5, [32], "zero-inflated negative binomial", 3, 2, 40, True, "gaussian"),
    ("zip-nobn", 77, 4, [20], "zero-inflated poisson", 1, 2, 33, False, "gaussian"),
    ("nb-unitvar", 64, 6, [16], "negative binomial", 2, 1, 21, True, "unit-variance gaussian"),
]


def _setup(case, tensor_cores):
    from scvae_b200.engine import VAEEngine
    name, G, L, hidden, lik, R, S, B, bn, latent = case
    cfg = O.VAEConfig(G, L, hidden, lik, latent, R, S, bn, True, kl_weight=0.7)
    params = O.vae_init_params(cfg, seed=3, dtype=torch.float64)
    gen = torch.Generator().manual_seed(11)
    for k in params:   # non-trivial biases / BN state so that every term is exercised
        if k.endswith("biases") or k.endswith("beta"):
            params[k] = torch.randn(params[k].shape, generator=gen, dtype=torch.float64) * 0.1
        if k.endswith("moving_mean"):
            params[k] = torch.randn(params[k].shape, generator=gen, dtype=torch.float64) * 0.2
        if k.endswith("moving_variance"):
            params[k] = torch.rand(params[k].shape, generator=gen, dtype=torch.float64) + 0.5
    x, _ = O.synthetic_counts(B, G, n_types=3, seed=5, target_zero_fraction=0.8)
    x = numpy.minimum(x, 500.0)
    if not bn:  # keep un-normalised activations out of the sigmoid-saturation hazard (Q1)
        x = numpy.minimum(x, 6.0)
        for k in params:
            if k.endswith("weights"):
                params[k] = params[k] * 0.3
    eps = torch.randn(R * S, B, L, generator=gen, dtype=torch.float64)
    if bn:  # plausible moving statistics: the batch statistics of this batch, perturbed
        upd = []
        O.vae_forward(cfg, params, torch.tensor(x, dtype=torch.float64),
                      torch.tensor(x, dtype=torch.float64), eps, True, bn_updates=upd)
        for scope, mean, var in upd:
            params[scope + "/BATCH_NORM/moving_mean"] = mean[0] * 0.9
            params[scope + "/BATCH_NORM/moving_variance"] = var[0] * 1.1
    eng = VAEEngine(G, L, hidden, lik, latent, bn, kl_weight=0.7, device="cuda:0",
                    tensor_cores=tensor_cores)
    eng.import_parameters(params)
    plan = eng._plan(B, R * S)
    eng.set_batch_dense(plan, torch.tensor(x).cuda())
    plan.eps.copy_(eps.reshape(R * S * B, L).float())
    return cfg, params, torch.tensor(x, dtype=torch.float64), eps, eng, plan


def _rel(a, b):
    a = numpy.asarray(a, dtype=numpy.float64)
    b = numpy.asarray(b, dtype=numpy.float64)
    return numpy.abs(a - b).max() / (numpy.abs(b).max() + 1e-30)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("tensor_cores", [False, True], ids=["fp32", "tf32"])
def test_vae_forward_backward_step(case, tensor_cores):
    name, G, L, hidden, lik, R, S, B, bn, latent = case
    cfg, params, x, eps, eng, plan = _setup(case, tensor_cores)
    w = 0.6
    tol = 5e-5 if not tensor_cores else 2e-3   # per-cell tensors; tf32 operands: 2^-10 truncation per product
    etol = 5e-5 if not tensor_cores else 1e-3  # ELBO terms: the north-star bound on every path
    state = O.AdamState(params)
    ref_params = {k: v.clone() for k, v in params.items()}
    out, grads = O.train_step(cfg, ref_params, state, x, x, eps, 1e-3, warm_up_weight=w)

    bound = eng.train_step(plan, R, S, 1e-3, warm_up_weight=w)
    torch.cuda.synchronize()
    bound = bound.cpu().numpy()
    # ELBO / reconstruction error / KL within 1e-3 relative (BASELINE north_star), much tighter here
    assert abs(bound[0] - out["lower_bound"].item()) <= etol * abs(out["lower_bound"].item())
    assert abs(bound[1] - out["lower_bound_weighted"].item()) <= etol * abs(out["lower_bound_weighted"].item())
    assert abs(bound[2] - out["reconstruction_error"].item()) <= etol * abs(out["reconstruction_error"].item())
    assert abs(bound[3] - out["kl_divergence"].item()) <= etol * abs(out["kl_divergence"].item()) + 1e-6
    # per-cell latent means and log-likelihoods
    assert _rel(plan.PH[:, :L].cpu(), out["q_z_mean"]) <= tol
    assert _rel(plan.logp.cpu(), out["log_p_x_given_z"].reshape(-1)) <= tol
    # raw gradients of every variable
    got = eng.export_gradients()
    gtol = 2e-4 if not tensor_cores else 1e-2
    gmax = max(g.abs().max().item() for g in grads.values())
    for k, g in grads.items():
        scale = g.abs().max().item()
        err = (got[k].double() - g).abs().max().item()
        # biases in front of a batch norm have an exactly-zero gradient: absolute floor
        assert err <= gtol * scale + 1e-5 * gmax, (k, err, scale)
    # parameters after clip + Adam, BN moving statistics.  Adam's first step is
    # lr * g / (|g| + eps'): where |g| is at fp32-noise level the sign is noise in ANY fp32
    # implementation (the reference included), so those entries are excluded.
    new = eng.export_parameters()
    noise = (1e-3 if not tensor_cores else 3e-2) * gmax
    for k, v in ref_params.items():
        diff = (new[k].double() - v).abs()
        if k in grads:
            diff = diff * (grads[k].abs() > noise)
        rtol = 5e-5 if "moving" in k else 1e-5   # fp32 batch variance of raw-count layers
        assert diff.max().item() <= rtol * max(v.abs().max().item(), 1.0), (k, diff.max().item())
    assert eng.global_step == 1


@pytest.mark.parametrize("case", CASES[:3], ids=[c[0] for c in CASES[:3]])
def test_vae_evaluate_mode(case):
    name, G, L, hidden, lik, R, S, B, bn, latent = case
    cfg, params, x, eps, eng, plan = _setup(case, False)
    out = O.vae_forward(cfg, params, x, x, eps, is_training=False, moments=True)
    eng.forward(plan, False, R, S, 1.0)
    m = eng.moments(plan, R, S)
    kn = eng.kl_neurons(plan)
    torch.cuda.synchronize()
    b = plan.bound.cpu().numpy()
    assert abs(b[0] - out["lower_bound"].item()) <= 5e-5 * abs(out["lower_bound"].item())
    assert abs(b[2] - out["reconstruction_error"].item()) <= 5e-5 * abs(out["reconstruction_error"].item())
    assert _rel(kn.cpu(), out["kl_divergence_neurons"]) <= 5e-5
    assert _rel(m[0].cpu(), out["p_x_mean"]) <= 1e-4
    assert _rel(m[1].cpu(), out["p_x_stddev"]) <= 1e-4
    sd_err = (m[2].cpu().double() - out["stddev_of_p_x_given_z_mean"]).abs().max().item()
    assert sd_err <= 1e-4 * out["p_x_mean"].abs().max().item()
    # deterministic z = q_z_mean (use_deterministic_z, VAE:2353-2362)
    out_d = O.vae_forward(cfg, params, x, x, eps, is_training=False, use_deterministic_z=True)
    eng.forward(plan, False, R, S, 1.0, deterministic=True)
    torch.cuda.synchronize()
    assert abs(plan.bound.cpu().numpy()[0] - out_d["lower_bound"].item()) <= 5e-5 * abs(out_d["lower_bound"].item())


def test_vae_csr_input_matches_dense():
    import scipy.sparse
    case = CASES[1]
    name, G, L, hidden, lik, R, S, B, bn, latent = case
    cfg, params, x, eps, eng, plan = _setup(case, False)
    eng.forward(plan, True, R, S, 1.0, update_moving=False)
    torch.cuda.synchronize()
    ref = plan.bound.cpu().numpy().copy()
    csr = scipy.sparse.csr_matrix(x.numpy().astype(numpy.float32))
    dev = torch.device("cuda:0")
    eng.set_batch_csr(plan, torch.tensor(csr.indptr.astype(numpy.int64)).to(dev),
                      torch.tensor(csr.indices.astype(numpy.int32)).to(dev),
                      torch.tensor(csr.data).to(dev), torch.arange(B, device=dev))
    eng.forward(plan, True, R, S, 1.0, update_moving=False)
    torch.cuda.synchronize()
    assert numpy.allclose(plan.bound.cpu().numpy(), ref, rtol=2e-6)


@pytest.mark.parametrize("R,S,deterministic", [(1, 1, False), (2, 2, False), (1, 1, True)])
def test_vae_lean_evaluation_pass_matches_full_path(R, S, deterministic):
    """Per-epoch evaluation passes: 16-bit minibatch + forward-only fused heads (keep_heads=False)
    against the fp32 path that materialises the head pre-activations."""
    import scipy.sparse
    from scvae_b200.engine import VAEEngine
    from scvae_b200.hotloop import ResidentCSR
    G, L, H, B = 256, 8, [32], 128
    dev = torch.device("cuda:0")
    x, _ = O.synthetic_counts(B, G, n_types=3, seed=60, target_zero_fraction=0.85)
    x = numpy.minimum(x, 300.0)
    data = ResidentCSR(scipy.sparse.csr_matrix(x.astype(numpy.float32)), dev)
    eng = VAEEngine(G, L, H, "negative binomial", device=dev, tensor_cores=True, seed=3)
    plan = eng._plan(B, R * S)
    idx = torch.arange(B, device=dev)
    eng.sample_noise(plan, 5, 0)
    results = []
    for lean in (False, True):
        eng.set_batch_csr(plan, data.indptr, data.indices, data.values, idx, u16_ok=data.u16_ok,
                          f16_exact=data.f16_exact, train16=lean, row_const_all=data.row_const)
        # batch statistics (an untrained model's moving averages do not normalise raw counts)
        eng.forward(plan, True, R, S, 1.0, deterministic=deterministic, update_moving=False,
                    keep_heads=not lean)
        torch.cuda.synchronize()
        results.append((plan.bound.cpu().numpy().copy(), plan.logp.cpu().numpy().copy()))
    (b0, lp0), (b1, lp1) = results
    assert plan.have_t16          # the lean pass really took the 16-bit / fused route
    assert abs(b0[0] - b1[0]) <= 1e-3 * abs(b0[0])
    assert abs(b0[2] - b1[2]) <= 1e-3 * abs(b0[2])
    m = B if deterministic else R * S * B
    assert numpy.abs(lp0[:m] - lp1[:m]).max() <= 3e-3 * numpy.abs(lp0[:m]).max()


def test_fused_training_step_is_bit_reproducible():
    """Same parameters, data and noise -> bit-identical parameters after several steps: every
    reduction on the training path has a fixed order (no atomics, no multi-contributor
    reduce-adds)."""
    import scipy.sparse
    from scvae_b200.engine import VAEEngine
    from scvae_b200.hotloop import ResidentCSR, TrainLoop
    G, L, H, B, N = 1024, 8, [48], 256, 512
    dev = torch.device("cuda:0")
    x, _ = O.synthetic_counts(N, G, n_types=3, seed=61, target_zero_fraction=0.9)
    x = numpy.minimum(x, 500.0)
    csr = scipy.sparse.csr_matrix(x.astype(numpy.float32))
    finals = []
    for _ in range(2):
        eng = VAEEngine(G, L, H, "negative binomial", device=dev, tensor_cores=True, seed=5)
        loop = TrainLoop(eng, B, seed=3, use_graph=True)
        data = ResidentCSR(csr, dev)
        for i in range(4):
            loop.rows.copy_(torch.arange(B, device=dev) + (i % 2) * B)
            loop.step(data, 1e-3, 1.0)
        torch.cuda.synchronize()
        assert loop.plan.fused_done
        finals.append((eng.store.param.clone(), eng.store.grad.clone(), loop.plan.bound.clone()))
    assert torch.equal(finals[0][1], finals[1][1]), "gradients differ between identical runs"
    assert torch.equal(finals[0][0], finals[1][0]), "parameters differ between identical runs"
    assert torch.equal(finals[0][2], finals[1][2])


EXTRA_CASES = [
    # name, G, L, hidden, likelihood, B, engine/oracle options
    ("batch-correction", 96, 6, [32], "negative binomial", 48, dict(number_of_batches=3)),
    ("count-sum-feature", 96, 6, [32], "poisson", 48, dict(count_sum_feature=True)),
    ("both-two-layers", 104, 5, [40, 24], "zero-inflated negative binomial", 40,
     dict(number_of_batches=4, count_sum_feature=True)),
    ("lfm-generative", 96, 6, [32], "negative binomial", 48,
     dict(generative_architecture="LFM", number_of_batches=2)),
    ("lfm-inference", 96, 6, [32], "negative binomial", 48, dict(inference_architecture="LFM")),
    ("lfm-both", 96, 6, [32], "poisson", 48,
     dict(inference_architecture="LFM", generative_architecture="LFM", count_sum_feature=True)),
    ("constrained-poisson", 96, 6, [32], "constrained poisson", 48, dict()),
    ("constrained-poisson-extras", 100, 5, [24], "constrained poisson", 40,
     dict(number_of_batches=2, count_sum_feature=True)),
    ("piecewise-nb-k2", 96, 6, [32], "negative binomial", 48, dict(number_of_reconstruction_classes=2)),
    ("piecewise-zip-k3-extras", 100, 5, [24], "zero-inflated poisson", 40,
     dict(number_of_reconstruction_classes=3, number_of_batches=2)),
]


@pytest.mark.parametrize("case", EXTRA_CASES, ids=[c[0] for c in EXTRA_CASES])
@pytest.mark.parametrize("tensor_cores", [False, True], ids=["fp32", "tc16"])
def test_vae_decoder_extras_and_linear_factor_architectures(case, tensor_cores):
    """Batch correction (one-hot batch index) and the count-sum feature concatenated to z
    (VAE:2400-2441), LFM inference / generative architectures (VAE:2221-2239, :2443-2462):
    forward, every gradient and one clip + Adam step against the oracle."""
    from scvae_b200.engine import VAEEngine
    name, G, L, hidden, lik, B, opts = case
    cfg = O.VAEConfig(G, L, hidden, lik, "gaussian", 1, 1, True, True, kl_weight=1.0, **opts)
    params = O.vae_init_params(cfg, seed=4, dtype=torch.float64)
    gen = torch.Generator().manual_seed(12)
    for k in params:
        if k.endswith("biases") or k.endswith("beta"):
            params[k] = torch.randn(params[k].shape, generator=gen, dtype=torch.float64) * 0.1
        if cfg.inference_architecture == "LFM" and k.startswith("POSTERIOR") and k.endswith("weights"):
            params[k] = params[k] * 0.05        # raw counts feed the posterior heads directly
    x, _ = O.synthetic_counts(B, G, n_types=3, seed=6, target_zero_fraction=0.8)
    x = numpy.minimum(x, 30.0)
    x64 = torch.tensor(x, dtype=torch.float64)
    eps = torch.randn(1, B, L, generator=gen, dtype=torch.float64)
    feats = {}
    if cfg.number_of_batches:
        feats["batch_indices"] = torch.randint(0, cfg.number_of_batches, (B, 1), generator=gen)
    if cfg.count_sum_feature:
        cs = x64.sum(dim=1)
        feats["count_sum_feature"] = ((cs - cs.min()) / (cs.max() - cs.min())).reshape(B, 1)
    constrained = lik == "constrained poisson"
    if constrained:
        feats["count_sum"] = x64.sum(dim=1, keepdim=True)
    state = O.AdamState(params)
    ref_params = {k: v.clone() for k, v in params.items()}
    out, grads = O.train_step(cfg, ref_params, state, x64, x64, eps, 1e-3, **feats)

    eng = VAEEngine(G, L, hidden, lik, "gaussian", True, device="cuda:0", tensor_cores=tensor_cores,
                    **opts)
    eng.import_parameters(params)
    plan = eng._plan(B, 1)
    eng.set_batch_dense(plan, torch.tensor(x).cuda())
    eng.set_batch_features(plan, feats["batch_indices"].cuda() if "batch_indices" in feats else None,
                           feats["count_sum_feature"].float().cuda() if "count_sum_feature" in feats
                           else None)
    if constrained:
        eng.set_batch_count_sum_parameter(plan, feats["count_sum"].float().cuda())
    plan.eps.copy_(eps.reshape(B, L).float())
    if not tensor_cores:     # evaluation mode (moving statistics) with the extras, exact path
        out_e = O.vae_forward(cfg, params, x64, x64, eps, is_training=False, **feats)
        eng.forward(plan, False, 1, 1, 1.0)
        torch.cuda.synchronize()
        be = plan.bound.cpu().numpy()
        assert abs(be[0] - out_e["lower_bound"].item()) <= 5e-5 * abs(out_e["lower_bound"].item())
    bound = eng.train_step(plan, 1, 1, 1e-3).cpu().numpy()
    torch.cuda.synchronize()
    if tensor_cores:
        # the 16-bit fused heads path handles the wider decoder input (not the softmax-coupled
        # constrained Poisson, which has its own row kernel)
        assert plan.fused_done == (not constrained and not cfg.k_max)
    tol = 5e-5 if not tensor_cores else 2e-3
    etol = 5e-5 if not tensor_cores else 1e-3
    assert abs(bound[0] - out["lower_bound"].item()) <= etol * abs(out["lower_bound"].item())
    assert _rel(plan.PH[:, :L].cpu(), out["q_z_mean"]) <= tol
    assert _rel(plan.logp.cpu(), out["log_p_x_given_z"].reshape(-1)) <= tol
    got = eng.export_gradients()
    gtol = 2e-4 if not tensor_cores else 1e-2
    gmax = max(g.abs().max().item() for g in grads.values())
    assert set(got) >= set(grads)
    for k, g in grads.items():
        assert got[k].shape == g.shape, (k, got[k].shape, g.shape)
        err = (got[k].double() - g).abs().max().item()
        assert err <= gtol * g.abs().max().item() + 1e-5 * gmax, (k, err)
    new = eng.export_parameters()
    noise = (1e-3 if not tensor_cores else 3e-2) * gmax
    for k, v in ref_params.items():
        if "moving" in k:
            continue
        diff = (new[k].double() - v).abs()
        if k in grads:
            diff = diff * (grads[k].abs() > noise)
        assert diff.max().item() <= 1e-5 * max(v.abs().max().item(), 1.0) + (2e-3 if tensor_cores else 0), k


MID_CASES = [
    # name, G, L, hidden, likelihood, B, bn, engine/oracle options
    ("one-layer", 256, 10, [64], "negative binomial", 128, True, {}),
    ("ragged-slabs", 256, 10, [100], "poisson", 333, True, {}),
    ("two-layers", 512, 7, [48, 24], "zero-inflated negative binomial", 200, True, {}),
    ("three-layers-wide-latent", 256, 100, [96, 64, 127], "negative binomial", 150, True, {}),
    ("no-batch-norm", 256, 6, [40], "zero-inflated poisson", 96, False, {}),
    ("batch-correction-count-sum", 256, 6, [32, 16], "negative binomial", 130, True,
     dict(number_of_batches=3, count_sum_feature=True)),
    ("reference-default-minibatch", 1000, 10, [100], "negative binomial", 100, True, {}),
    # the first-layer weight gradient from dY1 as fp16 + rounding remainder (SCVAE_DY1_SPLIT=1; the
    # default is one loss-scaled fp16 dY1 with the bias column summed in fp32 by vae_mid_bwd)
    ("one-layer-split-dy1", 256, 10, [64], "negative binomial", 128, True, dict(dy1_split=True)),
    ("no-batch-norm-split-dy1", 256, 6, [40], "zero-inflated poisson", 96, False, dict(dy1_split=True)),
]


@pytest.mark.parametrize("case", MID_CASES, ids=[c[0] for c in MID_CASES])
def test_fused_middle_training_step_matches_oracle(case):
    """The 16-bit training step with the persistent middle kernels (vae_mid_fwd / vae_mid_bwd:
    batch norm, ReLU, posterior clip + sample + KL fused behind exact-fp32 products) against the
    oracle: bound terms, per-cell tensors, raw gradients, variables after clip + Adam."""
    from scvae_b200.engine import VAEEngine
    name, G, L, hidden, lik, B, bn, opts = case
    opts = dict(opts)
    dy1_split = opts.pop("dy1_split", None)
    cfg = O.VAEConfig(G, L, hidden, lik, "gaussian", 1, 1, bn, True, kl_weight=0.7, **opts)
    params = O.vae_init_params(cfg, seed=3, dtype=torch.float64)
    gen = torch.Generator().manual_seed(11)
    for k in params:
        if k.endswith("biases") or k.endswith("beta"):
            params[k] = torch.randn(params[k].shape, generator=gen, dtype=torch.float64) * 0.1
        if not bn and k.endswith("weights"):
            params[k] = params[k] * 0.3
    x, _ = O.synthetic_counts(B, G, n_types=3, seed=5, target_zero_fraction=0.8)
    x = numpy.minimum(x, 500.0 if bn else 6.0)
    eps = torch.randn(1, B, L, generator=gen, dtype=torch.float64)
    features = {}
    if opts.get("number_of_batches"):
        features["batch_indices"] = torch.randint(0, opts["number_of_batches"], (B, 1), generator=gen)
    if opts.get("count_sum_feature"):
        cs = torch.tensor(x.sum(axis=1, keepdims=True), dtype=torch.float64)
        features["count_sum_feature"] = cs / cs.max()
    x64 = torch.tensor(x, dtype=torch.float64)
    if bn:  # plausible moving statistics: the batch statistics of this batch, perturbed
        upd = []
        O.vae_forward(cfg, params, x64, x64, eps, True, bn_updates=upd, **features)
        for scope, mean, var in upd:
            params[scope + "/BATCH_NORM/moving_mean"] = mean[0] * 0.9
            params[scope + "/BATCH_NORM/moving_variance"] = var[0] * 1.1
    eng = VAEEngine(G, L, hidden, lik, "gaussian", bn, kl_weight=0.7, device="cuda:0",
                    tensor_cores=True, **opts)
    if dy1_split is not None:
        eng.dy1_split = dy1_split
    eng.import_parameters(params)
    plan = eng._plan(B, 1)
    eng.set_batch_dense(plan, torch.tensor(x).cuda())
    if features:
        eng.set_batch_features(
            plan, features["batch_indices"].float().cuda() if "batch_indices" in features else None,
            features["count_sum_feature"].float().cuda() if "count_sum_feature" in features else None)
    plan.eps.copy_(eps.reshape(B, L).float())
    w = 0.6
    state = O.AdamState(params)
    ref = {k: v.clone() for k, v in params.items()}
    out, grads = O.train_step(cfg, ref, state, x64, x64, eps, 1e-3, warm_up_weight=w, **features)
    bound = eng.train_step(plan, 1, 1, 1e-3, warm_up_weight=w).cpu().numpy()
    torch.cuda.synchronize()
    assert plan.fused_done and plan.mid_done, "the fused middle was not taken"
    assert not eng.mid_error(plan)
    for i, key in enumerate(["lower_bound", "lower_bound_weighted", "reconstruction_error",
                             "kl_divergence"]):
        r = out[key].item()
        assert abs(bound[i] - r) <= 2e-4 * abs(r) + 1e-6, (key, bound[i], r)
    # the middle is exact fp32: only the fp16 rounding of the first layer's weights reaches mu
    assert _rel(plan.PH[:, :L].cpu(), out["q_z_mean"]) <= 5e-4
    assert _rel(plan.logp.cpu(), out["log_p_x_given_z"].reshape(-1)) <= 2e-4
    got = eng.export_gradients()
    gmax = max(g.abs().max().item() for g in grads.values())
    for k, g in grads.items():
        err = (got[k].double() - g).abs().max().item()
        assert err <= 1e-2 * g.abs().max().item() + 1e-4 * gmax, (k, err, g.abs().max().item())
    new = eng.export_parameters()
    noise = 3e-2 * gmax
    for k, v in ref.items():
        diff = (new[k].double() - v).abs()
        if k in grads:
            diff = diff * (grads[k].abs() > noise)
        rtol = 1e-4 if "moving" in k else 1e-5
        assert diff.max().item() <= rtol * max(v.abs().max().item(), 1.0), (k, diff.max().item())
    # lean evaluation pass (moving statistics, forward-only heads) through the same middle kernel,
    # against the oracle on the ENGINE's variables (Adam's first step is lr * sign(g): where |g| is
    # at rounding level the two updates may differ by 2 lr, which is not what this part checks)
    mine = {k: v.double() for k, v in eng.export_parameters().items()}
    ev = O.vae_forward(cfg, mine, x64, x64, eps, is_training=False, **features)
    eng.forward(plan, False, 1, 1, 1.0, keep_heads=False)
    torch.cuda.synchronize()
    b = plan.bound.cpu().numpy()
    assert abs(b[0] - ev["lower_bound"].item()) <= 1e-3 * abs(ev["lower_bound"].item())
    assert abs(b[3] - ev["kl_divergence"].item()) <= 1e-3 * abs(ev["kl_divergence"].item()) + 1e-6
    assert _rel(eng.kl_neurons(plan).cpu(), ev["kl_divergence_neurons"]) <= 1e-3


CONTINUOUS_CASES = [
    ("gaussian", False), ("softplus gaussian", True), ("gamma", False), ("bernoulli", True),
    ("lomax", False), ("log-normal", True), ("exponentially_modified_gaussian", False),
]


@pytest.mark.parametrize("lik,tensor_cores", CONTINUOUS_CASES, ids=[c[0] for c in CONTINUOUS_CASES])
def test_vae_training_step_with_continuous_likelihoods(lik, tensor_cores):
    """One training step with each reconstruction distribution outside the count family against
    the oracle: bound terms, per-cell log p, gradients of every head."""
    from scvae_b200.engine import VAEEngine
    G, L, hidden, B = 72, 5, [20], 40
    cfg = O.VAEConfig(G, L, hidden, lik, "gaussian", 1, 1, True, True, kl_weight=0.9)
    params = O.vae_init_params(cfg, seed=7, dtype=torch.float64)
    gen = torch.Generator().manual_seed(3)
    for k in params:
        if k.endswith("biases") or k.endswith("beta"):
            params[k] = torch.randn(params[k].shape, generator=gen, dtype=torch.float64) * 0.1
    rng = numpy.random.RandomState(2)
    if lik == "bernoulli":
        x = (rng.rand(B, G) < 0.3).astype(numpy.float32)
    else:
        x = (rng.gamma(2.0, 1.0, (B, G)) + 0.05).astype(numpy.float32)
    eps = torch.randn(1, B, L, generator=gen, dtype=torch.float64)
    eng = VAEEngine(G, L, hidden, lik, "gaussian", True, kl_weight=0.9, device="cuda:0",
                    tensor_cores=tensor_cores)
    eng.import_parameters(params)
    plan = eng._plan(B, 1)
    eng.set_batch_dense(plan, torch.tensor(x).cuda())
    plan.eps.copy_(eps.reshape(B, L).float())
    x64 = torch.tensor(x, dtype=torch.float64)
    state = O.AdamState(params)
    ref = {k: v.clone() for k, v in params.items()}
    out, grads = O.train_step(cfg, ref, state, x64, x64, eps, 1e-3)
    bound = eng.train_step(plan, 1, 1, 1e-3).cpu().numpy()
    torch.cuda.synchronize()
    tol = 1e-4 if not tensor_cores else 1e-3
    for i, key in enumerate(["lower_bound", "lower_bound_weighted", "reconstruction_error", "kl_divergence"]):
        r = out[key].item()
        assert abs(bound[i] - r) <= tol * abs(r) + 1e-6, (key, bound[i], r)
    assert _rel(plan.logp.cpu(), out["log_p_x_given_z"].reshape(-1)) <= tol
    got = eng.export_gradients()
    gmax = max(g.abs().max().item() for g in grads.values())
    for k, g in grads.items():
        err = (got[k].double() - g).abs().max().item()
        assert err <= (2e-3 if not tensor_cores else 1e-2) * g.abs().max().item() + 1e-5 * gmax, (k, err)
    m = eng.moments(plan, 1, 1)
    assert all(t.shape == (B, G) for t in m)
