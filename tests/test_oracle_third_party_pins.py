"""The oracle's restatement of the TensorFlow / TFP primitives against INDEPENDENT third-party
implementations of the same documented arithmetic (CPU only).

TensorFlow 1.15 cannot be installed in this image, so the oracle's primitive
arithmetic cannot be compared with TensorFlow itself.  What can be done is to compare it with
implementations written by other people whose formulas are the ones TensorFlow documents:

* ``tf.train.AdamOptimizer`` ("epsilon hat" form: lr_t = lr sqrt(1 - b2^t) / (1 - b1^t),
  theta -= lr_t m / (sqrt(v) + eps))  <->  scikit-learn's ``AdamOptimizer`` (the same form);
* fused batch norm (biased variance in the normalisation, Bessel-corrected variance in the moving
  average, moving <- decay * moving + (1 - decay) * batch)  <->  ATen's ``batch_norm``;
* ``tfp.distributions.kl_divergence(Normal, Normal)``  <->  ``torch.distributions``;
* the max-shifted ``log_mean_exp`` (MU:129-137)  <->  ``scipy.special.logsumexp``;
* ``xavier_initializer(uniform=True)`` bound  <->  ``torch.nn.init.xavier_uniform_``.

These pin the restated formulas, not TensorFlow's code: the parity cap stated in DESIGN section 2
remains.
"""
import math

import numpy
import pytest
import scipy.special
import torch

from oracle import scvae_oracle as O

D = torch.float64


def test_adam_against_scikit_learn_over_many_steps():
    from sklearn.neural_network._stochastic_optimizers import AdamOptimizer
    rng = numpy.random.RandomState(0)
    w0 = [rng.randn(7, 5), rng.randn(5)]
    sk_params = [w.copy() for w in w0]
    sk = AdamOptimizer(sk_params, learning_rate_init=3e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-8)
    params = {"a/weights": torch.tensor(w0[0], dtype=D), "a/biases": torch.tensor(w0[1], dtype=D)}
    state = O.AdamState(params)
    for step in range(60):
        # inside [-1, 1]: the reference's clip_by_value (VAE:2751-2755) is then the identity;
        # a few tiny entries exercise the epsilon placement
        g = [rng.uniform(-1, 1, size=w.shape) for w in w0]
        g[0][0, 0] = 1e-9 * (step + 1)
        g[1][2] = 0.0
        sk.update_params(sk_params, g)
        O.adam_clip_step(params, {"a/weights": torch.tensor(g[0], dtype=D),
                                  "a/biases": torch.tensor(g[1], dtype=D)}, state, 3e-3)
        assert numpy.allclose(params["a/weights"].numpy(), sk_params[0], rtol=1e-12, atol=1e-15), step
        assert numpy.allclose(params["a/biases"].numpy(), sk_params[1], rtol=1e-12, atol=1e-15), step
    assert state.step == 60


def test_adam_clips_before_the_moments():
    """clip_by_value(g, -1, 1) feeds Adam (VAE:2751-2759): equal to scikit-learn's Adam on the
    clipped gradient, different from clipping the update."""
    from sklearn.neural_network._stochastic_optimizers import AdamOptimizer
    w0 = numpy.array([0.3, -0.2, 1.5])
    g = numpy.array([5.0, -0.4, -7.5])
    sk_params = [w0.copy()]
    sk = AdamOptimizer(sk_params, learning_rate_init=1e-2)
    sk.update_params(sk_params, [numpy.clip(g, -1, 1)])
    params = {"a/weights": torch.tensor(w0, dtype=D)}
    O.adam_clip_step(params, {"a/weights": torch.tensor(g, dtype=D)}, O.AdamState(params), 1e-2)
    assert numpy.allclose(params["a/weights"].numpy(), sk_params[0], rtol=1e-13)


@pytest.mark.parametrize("rows", [2, 3, 64])
def test_batch_norm_against_aten(rows):
    gen = torch.Generator().manual_seed(rows)
    y = torch.randn(rows, 5, generator=gen, dtype=D) * 3.0 + 1.0
    beta = torch.randn(5, generator=gen, dtype=D)
    mm = torch.randn(5, generator=gen, dtype=D)
    mv = torch.rand(5, generator=gen, dtype=D) + 0.5
    params = {"s/BATCH_NORM/beta": beta, "s/BATCH_NORM/moving_mean": mm.clone(),
              "s/BATCH_NORM/moving_variance": mv.clone()}
    upd = []
    out = O.batch_norm(y, "s", params, True, upd)
    O.apply_bn_updates(params, upd)
    # ATen: biased variance for the output, unbiased for the running average,
    # running <- (1 - momentum) running + momentum batch; center=True, scale=False -> weight None
    rm, rv = mm.clone(), mv.clone()
    ref = torch.nn.functional.batch_norm(y, rm, rv, weight=None, bias=beta, training=True,
                                         momentum=1.0 - O.BN_DECAY, eps=O.BN_EPSILON)
    assert torch.allclose(out, ref, rtol=1e-12, atol=1e-12)
    assert torch.allclose(params["s/BATCH_NORM/moving_mean"], rm, rtol=1e-12, atol=1e-14)
    assert torch.allclose(params["s/BATCH_NORM/moving_variance"], rv, rtol=1e-12, atol=1e-14)
    # evaluation: the moving statistics
    out_eval = O.batch_norm(y, "s", params, False, None)
    ref_eval = torch.nn.functional.batch_norm(y, rm, rv, weight=None, bias=beta, training=False,
                                              eps=O.BN_EPSILON)
    assert torch.allclose(out_eval, ref_eval, rtol=1e-12, atol=1e-12)


def test_batch_norm_groups_are_separate_ops():
    """One batch-norm op per cluster on shared variables (GMVAE:2859-2877): each row group is
    normalised with its own statistics and the moving averages are updated group after group."""
    gen = torch.Generator().manual_seed(4)
    y = torch.randn(12, 3, generator=gen, dtype=D)
    params = {"s/BATCH_NORM/beta": torch.zeros(3, dtype=D), "s/BATCH_NORM/moving_mean": torch.zeros(3, dtype=D),
              "s/BATCH_NORM/moving_variance": torch.ones(3, dtype=D)}
    upd = []
    out = O.batch_norm(y, "s", params, True, upd, groups=3)
    O.apply_bn_updates(params, upd)
    rm, rv = torch.zeros(3, dtype=D), torch.ones(3, dtype=D)
    refs = [torch.nn.functional.batch_norm(y[4 * k:4 * k + 4], rm, rv, training=True,
                                           momentum=1.0 - O.BN_DECAY, eps=O.BN_EPSILON) for k in range(3)]
    assert torch.allclose(out, torch.cat(refs), rtol=1e-12, atol=1e-12)
    assert torch.allclose(params["s/BATCH_NORM/moving_mean"], rm, rtol=1e-12, atol=1e-15)
    assert torch.allclose(params["s/BATCH_NORM/moving_variance"], rv, rtol=1e-12, atol=1e-15)


def test_analytic_gaussian_kl_against_torch_distributions():
    """KL(N(mu, sigma) || N(0, 1)) summed as the reference sums it (VAE:2624-2652): the oracle's
    `kl_divergence` output of a forward pass against torch.distributions on the same q(z|x)."""
    cfg = O.VAEConfig(14, 4, [6], "poisson")
    params = O.vae_init_params(cfg, seed=1, dtype=D)
    x = torch.tensor(O.synthetic_counts(9, 14, seed=2)[0], dtype=D).clamp(max=30)
    eps = torch.randn(1, 9, 4, generator=torch.Generator().manual_seed(3), dtype=D)
    out = O.vae_forward(cfg, params, x, x, eps, is_training=True)
    q = torch.distributions.Normal(out["q_z_mean"], torch.exp(out["log_sigma"]))
    p = torch.distributions.Normal(torch.zeros_like(q.loc), torch.ones_like(q.scale))
    kl = torch.distributions.kl_divergence(q, p)                   # (B, L)
    assert math.isclose(out["kl_divergence"].item(), kl.sum(dim=1).mean().item(), rel_tol=1e-12)


def test_log_mean_exp_against_scipy():
    rng = numpy.random.RandomState(2)
    a = rng.randn(5, 7, 3) * 40.0            # needs the max shift
    got = O.log_mean_exp(torch.tensor(a, dtype=D), 0).numpy()
    ref = scipy.special.logsumexp(a, axis=0) - math.log(a.shape[0])
    assert numpy.allclose(got, ref, rtol=1e-13, atol=1e-13)


def test_xavier_uniform_bound_matches_torch():
    gen = torch.Generator().manual_seed(0)
    w = O.xavier_uniform(gen, 300, 200, D)
    assert w.shape == (300, 200)
    ref = torch.empty(200, 300, dtype=D)
    torch.nn.init.xavier_uniform_(ref, generator=torch.Generator().manual_seed(1))
    bound = math.sqrt(6.0 / 500.0)
    assert w.abs().max().item() <= bound and ref.abs().max().item() <= bound
    # both fill the interval: the largest of 60 000 uniform draws lies within 0.1 % of the bound
    assert w.abs().max().item() > 0.999 * bound and ref.abs().max().item() > 0.999 * bound
    assert abs(w.var().item() - bound ** 2 / 3.0) < 0.02 * bound ** 2 / 3.0


# ---------------------------------------------------------------------------------------------
# the TF stand-in that produced the reference-graph fixtures (oracle/tf1_standin.py): the same two
# primitives, independently of the oracle
# ---------------------------------------------------------------------------------------------
def test_standin_adam_against_scikit_learn_with_carried_slots():
    from sklearn.neural_network._stochastic_optimizers import AdamOptimizer
    from oracle import tf1_standin as T
    rng = numpy.random.RandomState(5)
    w0 = rng.randn(4, 3)
    sk_params = [w0.copy()]
    sk = AdamOptimizer(sk_params, learning_rate_init=2e-3)
    w, m, v = w0.copy(), numpy.zeros_like(w0), numpy.zeros_like(w0)
    for step in range(12):
        c = rng.uniform(-1, 1, size=w0.shape)           # d loss / d w of loss = sum(c * w)
        c[0, 0] = 1e-9
        T.STATE.reset(initial={"L/weights": w, "__adam_m__": {"L/weights": m},
                               "__adam_v__": {"L/weights": v}, "__adam_step__": step})
        var = T.get_variable("L/weights", None)
        opt = T.AdamOptimizer(2e-3)
        pairs = opt.compute_gradients((var * torch.tensor(c)).sum())
        opt.apply_gradients([(T.clip_by_value(g, -1.0, 1.0), x) for g, x in pairs])
        w = T.STATE.updates["L/weights"].numpy()
        m = T.STATE.updates["__adam_m__/L/weights"].numpy()
        v = T.STATE.updates["__adam_v__/L/weights"].numpy()
        sk.update_params(sk_params, [c])
        assert numpy.allclose(w, sk_params[0], rtol=1e-12, atol=1e-15), step


@pytest.mark.parametrize("rows", [2, 5, 33])
def test_standin_batch_norm_against_aten(rows):
    from oracle import tf1_standin as T
    gen = torch.Generator().manual_seed(100 + rows)
    y = torch.randn(rows, 4, generator=gen, dtype=D) * 2.0 - 0.5
    beta = torch.randn(4, generator=gen, dtype=D)
    mm = torch.randn(4, generator=gen, dtype=D)
    mv = torch.rand(4, generator=gen, dtype=D) + 0.5
    T.STATE.reset(initial={"L/BATCH_NORM/beta": beta.numpy(), "L/BATCH_NORM/moving_mean": mm.numpy(),
                           "L/BATCH_NORM/moving_variance": mv.numpy()})
    with T.variable_scope("L"):
        out = T.batch_norm(y, is_training=True, scope="BATCH_NORM")
    rm, rv = mm.clone(), mv.clone()
    ref = torch.nn.functional.batch_norm(y, rm, rv, weight=None, bias=beta, training=True,
                                         momentum=0.001, eps=1e-3)
    assert torch.allclose(torch.as_tensor(out.detach()), ref, rtol=1e-12, atol=1e-12)
    assert torch.allclose(torch.as_tensor(T.STATE.updates["L/BATCH_NORM/moving_mean"]), rm, rtol=1e-12, atol=1e-14)
    assert torch.allclose(torch.as_tensor(T.STATE.updates["L/BATCH_NORM/moving_variance"]), rv, rtol=1e-12, atol=1e-14)
    with T.variable_scope("L"):
        out_eval = T.batch_norm(y, is_training=False, scope="BATCH_NORM")
    ref_eval = torch.nn.functional.batch_norm(y, mm, mv, weight=None, bias=beta, training=False, eps=1e-3)
    assert torch.allclose(torch.as_tensor(out_eval.detach()), ref_eval, rtol=1e-12, atol=1e-12)


def test_standin_normal_kl_against_torch_distributions():
    from oracle import tf1_standin as T
    gen = torch.Generator().manual_seed(8)
    mu, sigma = torch.randn(6, 3, generator=gen, dtype=D), torch.rand(6, 3, generator=gen, dtype=D) + 0.2
    T.STATE.reset()
    kl = T.kl_divergence(T.Normal(mu, sigma), T.Normal(torch.zeros_like(mu), torch.ones_like(mu)))
    ref = torch.distributions.kl_divergence(torch.distributions.Normal(mu, sigma),
                                            torch.distributions.Normal(torch.zeros_like(mu), torch.ones_like(mu)))
    assert torch.allclose(torch.as_tensor(kl), ref, rtol=1e-12, atol=1e-14)
    z = torch.randn(6, 3, generator=gen, dtype=D)
    assert torch.allclose(torch.as_tensor(T.Normal(mu, sigma).log_prob(z)),
                          torch.distributions.Normal(mu, sigma).log_prob(z), rtol=1e-12, atol=1e-13)
