"""Pins of the CPU oracle (the reference ships no tests or golden vectors: SURVEY §8c).

Closed forms against scipy.stats, the zero-inflation identities of
scvae/distributions/zero_inflated.py:180-199, analytic KL vs sampled KL, autograd vs fp64
finite differences, TF-Adam / TF-batch-norm semantics."""
import math

import numpy
import pytest
import scipy.special
import scipy.stats
import torch

from oracle import scvae_oracle as O

D = torch.float64


def test_poisson_matches_scipy():
    x = torch.arange(0, 60, dtype=D)
    for ll in (-3.0, 0.0, 1.7, 4.0):
        got = O.poisson_log_prob(x, torch.tensor(ll, dtype=D))
        ref = scipy.stats.poisson.logpmf(x.numpy(), math.exp(ll))
        assert numpy.allclose(got.numpy(), ref, rtol=1e-12, atol=1e-12)


def test_negative_binomial_matches_scipy():
    # tfp NegativeBinomial(total_count=r, probs=p) == scipy nbinom(n=r, p=1-p)  
    x = torch.arange(0, 80, dtype=D)
    for log_r in (-2.0, 0.3, 3.0):
        for p in (0.05, 0.5, 0.93):
            got = O.nb_log_prob(x, torch.tensor(p, dtype=D), torch.tensor(log_r, dtype=D))
            ref = scipy.stats.nbinom.logpmf(x.numpy(), math.exp(log_r), 1 - p)
            assert numpy.allclose(got.numpy(), ref, rtol=1e-10, atol=1e-10)
            theta = {"p": torch.tensor(p, dtype=D), "log_r": torch.tensor(log_r, dtype=D)}
            m, v = O.likelihood_moments("negative binomial", theta)
            assert math.isclose(m.item(), scipy.stats.nbinom.mean(math.exp(log_r), 1 - p), rel_tol=1e-10)
            assert math.isclose(v.item(), scipy.stats.nbinom.var(math.exp(log_r), 1 - p), rel_tol=1e-10)


@pytest.mark.parametrize("kind", ["zero-inflated poisson", "zero-inflated negative binomial"])
def test_zero_inflation_identities(kind):
    pi = torch.tensor(0.3, dtype=D)
    theta = {"pi": pi, "log_lambda": torch.tensor(0.8, dtype=D), "p": torch.tensor(0.6, dtype=D),
             "log_r": torch.tensor(1.1, dtype=D)}
    base = kind.replace("zero-inflated ", "")
    x = torch.arange(0, 400, dtype=D)
    lp = O.likelihood_log_prob(kind, x, theta)
    lp_base = O.likelihood_log_prob(base, x, theta)
    # log_prob(0) = log(pi + (1 - pi) p0); log_prob(x > 0) = log(1 - pi) + log p(x)
    assert math.isclose(lp[0].item(), math.log(0.3 + 0.7 * math.exp(lp_base[0].item())), rel_tol=1e-12)
    assert torch.allclose(lp[1:], math.log(0.7) + lp_base[1:], rtol=1e-12)
    # normalisation and moments by direct summation
    prob = torch.exp(lp)
    assert abs(prob.sum().item() - 1) < 1e-9
    m, v = O.likelihood_moments(kind, theta)
    mean = (prob * x).sum().item()
    assert math.isclose(m.item(), mean, rel_tol=1e-8)
    assert math.isclose(v.item(), (prob * (x - mean) ** 2).sum().item(), rel_tol=1e-7)


def test_head_clipping_matches_float32_bounds():
    a = torch.tensor([-200.0, -11.0, 0.0, 11.0, 200.0], dtype=D)
    assert O._clip_head(a, "log_r").tolist() == [-10.0, -10.0, 0.0, 10.0, 10.0]
    p = O._clip_head(a, "p")
    assert p[0].item() == pytest.approx(O.TINY) and p[-1].item() == 1.0   # quirk Q1


def test_analytic_kl_matches_sampled_and_monte_carlo():
    torch.manual_seed(0)
    cfg_a = O.VAEConfig(30, 4, [8], "poisson", analytical_kl_term=True, number_of_monte_carlo_samples=4000)
    cfg_s = O.VAEConfig(30, 4, [8], "poisson", analytical_kl_term=False, number_of_monte_carlo_samples=4000)
    params = O.vae_init_params(cfg_a, 1, D)
    x = torch.tensor(O.synthetic_counts(6, 30, seed=1)[0], dtype=D).clamp(max=20)
    eps = torch.randn(4000, 6, 4, dtype=D)
    a = O.vae_forward(cfg_a, params, x, x, eps)
    s = O.vae_forward(cfg_s, params, x, x, eps)
    assert abs(a["kl_divergence"].item() - s["kl_divergence"].item()) < 0.05 * a["kl_divergence"].item() + 0.02
    # closed form of KL(N(mu, sigma) || N(0, 1))
    mu, ls = a["q_z_mean"], a["log_sigma"]
    ref = (0.5 * (mu ** 2 + torch.exp(2 * ls) - 1) - ls).mean(0).sum()
    assert math.isclose(a["kl_divergence"].item(), ref.item(), rel_tol=1e-12)


def test_log_mean_exp():
    a = torch.randn(5, 7, dtype=D) * 30
    ref = torch.log(torch.exp(a - a.max()).mean(0)) + a.max()
    assert torch.allclose(O.log_mean_exp(a, 0), ref, rtol=1e-10)


@pytest.mark.parametrize("kind", list(O.LIKELIHOODS))
def test_vae_autograd_matches_finite_differences(kind):
    cfg = O.VAEConfig(12, 3, [5], kind, number_of_importance_samples=2, number_of_monte_carlo_samples=2)
    params = O.vae_init_params(cfg, 2, D)
    x = torch.tensor(O.synthetic_counts(4, 12, seed=3)[0], dtype=D).clamp(max=30)
    if kind in ("log-normal", "gamma"):
        x = x + 0.5          # strictly positive support
    eps = torch.randn(4, 4, 3, generator=torch.Generator().manual_seed(0), dtype=D)
    state = O.AdamState(params)
    base = {k: v.clone() for k, v in params.items()}
    _, grads = O.train_step(cfg, {k: v.clone() for k, v in base.items()}, state, x, x, eps, 1e-3, 0.5)
    rng = numpy.random.RandomState(0)
    for name in O.trainable_names(base):
        flat = base[name].reshape(-1)
        for idx in rng.choice(flat.numel(), size=min(3, flat.numel()), replace=False):
            h = 1e-6
            vals = []
            for sgn in (+1, -1):
                p2 = {k: v.clone() for k, v in base.items()}
                p2[name].reshape(-1)[idx] += sgn * h
                vals.append(-O.vae_forward(cfg, p2, x, x, eps, True, 0.5)["lower_bound_weighted"].item())
            fd = (vals[0] - vals[1]) / (2 * h)
            an = grads[name].reshape(-1)[idx].item()
            assert abs(fd - an) <= 1e-5 * max(1.0, abs(fd)), (name, idx, fd, an)


def test_gmvae_bound_decomposition_and_gradients_exist():
    cfg = O.GMVAEConfig(14, 3, 4, [6], "negative binomial", number_of_monte_carlo_samples=2)
    params = O.gmvae_init_params(cfg, 0, D)
    x = torch.tensor(O.synthetic_counts(5, 14, seed=2)[0], dtype=D).clamp(max=30)
    eps = torch.randn(4, 2, 5, 3, dtype=D)
    out = O.gmvae_forward(cfg, params, x, x, eps, moments=True)
    assert math.isclose(out["lower_bound"].item(),
                        (out["reconstruction_error"] - out["kl_divergence_z"] - out["kl_divergence_y"]).item(),
                        rel_tol=1e-12)
    assert torch.allclose(out["y"].sum(1), torch.ones(5, dtype=D))
    # KL_y = log K - H[q(y|x)] >= 0 for a uniform prior
    assert (out["kl_y"] >= -1e-12).all()
    state = O.AdamState(params)
    _, grads = O.train_step(cfg, params, state, x, x, eps, 1e-3)
    assert all(g is not None for g in grads.values())


def test_tf_adam_semantics():
    # theta -= lr sqrt(1-b2^t)/(1-b1^t) m / (sqrt(v) + eps): epsilon outside the bias correction
    params = {"w/weights": torch.tensor([1.0, -2.0], dtype=D)}
    state = O.AdamState(params)
    g = torch.tensor([0.5, -3.0], dtype=D)     # second entry is clipped to -1
    O.adam_clip_step(params, {"w/weights": g}, state, 0.1)
    gc = torch.tensor([0.5, -1.0], dtype=D)
    m, v = 0.1 * gc, 0.001 * gc ** 2
    lr_t = 0.1 * math.sqrt(1 - 0.999) / (1 - 0.9)
    ref = torch.tensor([1.0, -2.0], dtype=D) - lr_t * m / (torch.sqrt(v) + 1e-8)
    assert torch.allclose(params["w/weights"], ref, rtol=1e-12)


def test_tf_batch_norm_semantics():
    y = torch.tensor([[1.0, 2.0], [3.0, 6.0], [5.0, 1.0]], dtype=D)
    params = {"s/BATCH_NORM/beta": torch.tensor([0.5, -0.5], dtype=D),
              "s/BATCH_NORM/moving_mean": torch.zeros(2, dtype=D),
              "s/BATCH_NORM/moving_variance": torch.ones(2, dtype=D)}
    upd = []
    out = O.batch_norm(y, "s", params, True, upd)
    mean, var = y.mean(0), y.var(0, unbiased=False)
    assert torch.allclose(out, (y - mean) / torch.sqrt(var + 1e-3) + params["s/BATCH_NORM/beta"])
    O.apply_bn_updates(params, upd)
    assert torch.allclose(params["s/BATCH_NORM/moving_mean"], 0.001 * mean)
    assert torch.allclose(params["s/BATCH_NORM/moving_variance"], 0.999 + 0.001 * y.var(0, unbiased=True))


def test_synthetic_counts_shape_and_sparsity():
    x, labels = O.synthetic_counts(500, 40, n_types=4, target_zero_fraction=0.9)
    assert x.shape == (500, 40) and x.dtype == numpy.float32 and labels.shape == (500,)
    assert numpy.array_equal(x, numpy.round(x)) and x.min() >= 0
    assert 0.8 < (x == 0).mean() < 0.99


def test_constrained_poisson_matches_scipy():
    """Poisson(rate = softmax(a) * N): log-pmf against scipy, moments = rate."""
    import scipy.stats
    rng = numpy.random.RandomState(3)
    a = torch.tensor(rng.randn(5, 9), dtype=D)
    n = torch.tensor(rng.randint(1, 60, size=(5, 1)), dtype=D)
    x = torch.tensor(rng.poisson(3.0, size=(5, 9)), dtype=D)
    theta = {"lambda": O._clip_head(a, "lambda")}
    lp = O.likelihood_log_prob("constrained poisson", x, theta, n)
    rate = (torch.softmax(a, dim=-1) * n).numpy()
    assert numpy.allclose(lp.numpy(), scipy.stats.poisson.logpmf(x.numpy(), rate), rtol=1e-12, atol=1e-12)
    m, v = O.likelihood_moments("constrained poisson", theta, n)
    assert numpy.allclose(m.numpy(), rate) and numpy.allclose(v.numpy(), rate)
    assert numpy.allclose(m.sum(dim=1).numpy(), n.reshape(-1).numpy())      # rates sum to N


@pytest.mark.parametrize("kind", ["poisson", "negative binomial", "zero-inflated negative binomial"])
@pytest.mark.parametrize("k_max", [1, 3])
def test_piecewise_categorical_normalisation_and_moments(kind, k_max):
    """Categorised (CAT:210-274): the pmf sums to one and its mean / variance equal the closed
    forms, by direct summation over the support."""
    rng = numpy.random.RandomState(5)
    n = 6
    heads = O.LIKELIHOODS[kind]
    a = {h: torch.tensor(rng.randn(n) * 0.7, dtype=D) for h in heads}
    theta = {h: O._clip_head(v, h) for h, v in a.items()}
    cat = torch.log_softmax(torch.tensor(rng.randn(n, k_max + 1), dtype=D), dim=-1)
    xs = torch.arange(0, 4000, dtype=D)
    theta_b = {h: v.unsqueeze(-1) for h, v in theta.items()}
    lp = O.piecewise_log_prob(kind, xs.unsqueeze(0).expand(n, -1), theta_b, cat.unsqueeze(1).expand(-1, len(xs), -1), k_max)
    p = torch.exp(lp)
    assert torch.allclose(p.sum(dim=1), torch.ones(n, dtype=D), atol=1e-9)
    m, v = O.likelihood_moments(kind, theta)
    mean, var = O.piecewise_moments(m, v, cat, k_max)
    mean_direct = (p * xs).sum(dim=1)
    var_direct = (p * xs * xs).sum(dim=1) - mean_direct ** 2
    assert torch.allclose(mean, mean_direct, rtol=1e-8, atol=1e-9)
    assert torch.allclose(var, var_direct, rtol=1e-7, atol=1e-8)


def test_dropout_sites_and_gradients():
    """Dropout restated for the next build step (MU:45-50; independent masks per head,
    VAE:2286,2487): identity in evaluation, inverted scaling in training, autograd vs finite
    differences with the masks held fixed."""
    cfg = O.VAEConfig(12, 3, [6, 5], "negative binomial", dropout_keep_probabilities=[0.8, 0.7, 0.9])
    params = O.vae_init_params(cfg, 2, D)
    x = torch.tensor(O.synthetic_counts(5, 12, seed=3)[0], dtype=D).clamp(max=30)
    eps = torch.randn(1, 5, 3, generator=torch.Generator().manual_seed(0), dtype=D)
    drop = {"generator": torch.Generator().manual_seed(4)}
    out = O.vae_forward(cfg, params, x, x, eps, True, dropout=drop)
    sites = set(drop["masks"])
    assert sites == {"ENCODER/1", "ENCODER/2", "POSTERIOR/MU", "POSTERIOR/LOG_SIGMA", "DECODER/2",
                     "DECODER/1", "X_TILDE/P", "X_TILDE/LOG_R"}
    assert drop["masks"]["ENCODER/1"].shape == x.shape            # keep_x on the input counts
    assert not torch.equal(drop["masks"]["X_TILDE/P"], drop["masks"]["X_TILDE/LOG_R"])
    # evaluation mode and all-ones masks are the un-dropped graph (up to the 1 / keep scaling)
    plain = O.vae_forward(cfg, params, x, x, eps, False)["lower_bound"]
    assert torch.equal(O.vae_forward(cfg, params, x, x, eps, False, dropout=drop)["lower_bound"], plain)
    assert not torch.equal(out["lower_bound"], O.vae_forward(cfg, params, x, x, eps, True)["lower_bound"])
    # gradients with the masks fixed
    names = O.trainable_names(params)
    leaves = {k: params[k].clone().requires_grad_(True) for k in names}
    local = {k: leaves.get(k, v) for k, v in params.items()}
    loss = -O.vae_forward(cfg, local, x, x, eps, True, dropout=drop)["lower_bound_weighted"]
    grads = dict(zip(names, torch.autograd.grad(loss, [leaves[k] for k in names])))
    rng = numpy.random.RandomState(1)
    for name in names:
        flat = params[name].reshape(-1)
        idx = int(rng.randint(flat.numel()))
        vals = []
        for sgn in (+1, -1):
            p2 = {k: v.clone() for k, v in params.items()}
            p2[name].reshape(-1)[idx] += sgn * 1e-6
            vals.append(-O.vae_forward(cfg, p2, x, x, eps, True, dropout=drop)["lower_bound_weighted"].item())
        fd = (vals[0] - vals[1]) / 2e-6
        assert abs(fd - grads[name].reshape(-1)[idx].item()) <= 1e-5 * max(1.0, abs(fd)), name


# ---- continuous / binary reconstruction distributions against scipy.stats -------------
def _theta_for(kind, rng, n):
    import torch
    t = lambda a: torch.tensor(a, dtype=torch.float64)
    if kind == "gaussian":
        return {"mu": t(rng.randn(n) * 2), "log_sigma": t(rng.uniform(-2, 2, n))}
    if kind in ("softplus gaussian", "modified gaussian"):
        return {"mean": t(rng.randn(n) * 2), "softplus_scale": t(rng.randn(n) * 2)}
    if kind == "log-normal":
        return {"mean": t(rng.randn(n)), "variance": t(rng.uniform(0.1, 2.0, n))}
    if kind == "gamma":
        return {"concentration": t(rng.uniform(0.3, 5.0, n)), "rate": t(rng.uniform(0.2, 3.0, n))}
    if kind == "bernoulli":
        return {"logits": t(rng.randn(n) * 3)}
    if kind == "lomax":
        return {"log_concentration": t(rng.uniform(-1.0, 2.0, n)), "log_scale": t(rng.uniform(-1.0, 2.0, n))}
    if kind == "exponentially_modified_gaussian":
        return {"location": t(rng.randn(n)), "scale": t(rng.uniform(0.3, 2.0, n)),
                "rate": t(rng.uniform(0.3, 3.0, n))}
    raise KeyError(kind)


@pytest.mark.parametrize("kind", ["gaussian", "softplus gaussian", "log-normal", "gamma", "bernoulli",
                                  "lomax", "exponentially_modified_gaussian"])
def test_continuous_likelihoods_match_scipy(kind):
    """log-density and moments of the restated closed forms against scipy.stats (an independent
    implementation): Normal, lognorm, gamma, bernoulli, lomax, exponnorm."""
    import scipy.stats as st
    import torch
    rng = numpy.random.RandomState(4)
    n = 200
    th = _theta_for(kind, rng, n)
    if kind == "bernoulli":
        x = rng.randint(0, 2, n).astype(float)
    elif kind in ("gaussian", "softplus gaussian", "exponentially_modified_gaussian"):
        x = rng.randn(n) * 3
    else:
        x = rng.gamma(2.0, 1.5, n)
    lp = O.continuous_log_prob(kind, torch.tensor(x), th).numpy()
    m, v = (a.numpy() for a in O.continuous_moments(kind, th))
    g = {k: a.numpy() for k, a in th.items()}
    if kind == "gaussian":
        d = st.norm(g["mu"], numpy.exp(g["log_sigma"]))
    elif kind == "softplus gaussian":
        d = st.norm(g["mean"], numpy.sqrt(numpy.log1p(numpy.exp(g["softplus_scale"]))))
    elif kind == "log-normal":
        d = st.lognorm(s=numpy.sqrt(g["variance"]), scale=numpy.exp(g["mean"]))
    elif kind == "gamma":
        d = st.gamma(a=g["concentration"], scale=1.0 / g["rate"])
    elif kind == "bernoulli":
        d = st.bernoulli(1.0 / (1.0 + numpy.exp(-g["logits"])))
    elif kind == "lomax":
        d = st.lomax(c=numpy.exp(g["log_concentration"]), scale=numpy.exp(g["log_scale"]))
    else:
        # exponnorm(K = 1 / (scale rate), loc, scale)
        d = st.exponnorm(K=1.0 / (g["scale"] * g["rate"]), loc=g["location"], scale=g["scale"])
    ref = d.logpmf(x) if kind == "bernoulli" else d.logpdf(x)
    ok = numpy.ones(n, bool)
    if kind == "exponentially_modified_gaussian":
        # the reference clips erfc(w) at float32 tiny before the logarithm
        # (exponentially_modified_normal.py:218-222): beyond w ~ 9.3 its density is a floor, not
        # the distribution's -- compared only where the clip is inactive
        u, v_ = g["rate"] * (x - g["location"]), g["rate"] * g["scale"]
        ok = (v_ * v_ - u) / (numpy.sqrt(2.0) * v_) < 9.0
        assert ok.sum() > n // 2
    assert numpy.allclose(lp[ok], ref[ok], rtol=1e-9, atol=1e-9), numpy.abs(lp - ref)[ok].max()
    if kind == "lomax":      # moments only where they exist (the reference returns nan / inf elsewhere)
        c = numpy.exp(g["log_concentration"])
        ok = c > 1.0
        assert numpy.allclose(m[ok], d.mean()[ok], rtol=1e-9)
        assert numpy.isnan(m[~ok]).all()
        assert numpy.isinf(v[(c > 1.0) & (c <= 2.0)]).all() and numpy.isnan(v[c <= 1.0]).all()
    else:
        assert numpy.allclose(m, d.mean(), rtol=1e-9)
        assert numpy.allclose(v, d.var(), rtol=1e-9)


@pytest.mark.parametrize("kind", ["gaussian", "softplus gaussian", "gamma", "bernoulli", "lomax",
                                  "exponentially_modified_gaussian", "log-normal"])
def test_vae_forward_with_continuous_likelihoods_runs_and_differentiates(kind):
    """The VAE graph with every continuous head specification: activations, clips and a finite,
    differentiable bound (autograd vs central finite differences on one weight)."""
    import torch
    cfg = O.VAEConfig(12, 3, [8], kind)
    params = O.vae_init_params(cfg, seed=1, dtype=torch.float64)
    rng = numpy.random.RandomState(0)
    if kind == "bernoulli":
        x = torch.tensor((rng.rand(10, 12) < 0.3).astype(float))
    else:
        x = torch.tensor(rng.gamma(2.0, 1.0, (10, 12)) + 0.05)
    eps = torch.randn(1, 10, 3, generator=torch.Generator().manual_seed(2), dtype=torch.float64)
    name = "X_TILDE/{}/DENSE/weights".format(O.LIKELIHOODS[kind][0].upper())
    w = params[name].clone().requires_grad_(True)
    local = dict(params)
    local[name] = w
    out = O.vae_forward(cfg, local, x, x, eps, is_training=True)
    assert torch.isfinite(out["lower_bound"])
    out["lower_bound"].backward()
    h = 1e-6
    fd = []
    for sign in (+1, -1):
        pert = dict(params)
        pert[name] = params[name].clone()
        pert[name][0, 0] += sign * h
        fd.append(O.vae_forward(cfg, pert, x, x, eps, is_training=True)["lower_bound"].item())
    assert abs((fd[0] - fd[1]) / (2 * h) - w.grad[0, 0].item()) <= 1e-5 * max(1.0, abs(w.grad[0, 0].item()))
