"""The oracle against golden vectors produced by the REFERENCE'S OWN graph code.

``tests/golden/reference_graph_*.npz`` were recorded by ``oracle/make_golden_graph.py``: the
reference's ``VariationalAutoencoder`` / ``GaussianMixtureVariationalAutoencoder`` classes
(VAE:2219-2770, GMVAE:2788-3470, MU:38-137, DU:30-306, ZI:180-199, CAT:210-274), imported
unmodified from /root/reference and executed over an eager stand-in for the TF-1.x / TFP-0.7
primitives (``oracle/tf1_standin.py``), in fp64.  They pin the oracle's graph composition --
variable names, layer / head order, clips, tiling over (R, S, B), zero-inflated and
piecewise-categorical formulas, KL / ELBO aggregation, moments, clip + Adam, batch-norm update
order -- to the reference's code; TF's own primitive arithmetic stays restated (see the
stand-in's docstring).  Nothing here reads /root/reference.
"""
import glob
import json
import os
from collections import OrderedDict

import numpy
import pytest
import torch

from oracle import scvae_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[len("reference_graph_"):-len(".npz")]
               for p in glob.glob(os.path.join(GOLDEN, "reference_graph_*.npz")))
D = torch.float64
RTOL, ATOL = 1e-9, 1e-10      # fp64 oracle against fp64 golden vectors


def load_case(name):
    data = numpy.load(os.path.join(GOLDEN, "reference_graph_{}.npz".format(name)))
    meta = json.loads(str(data["meta"]))
    groups = {}
    for key in data.files:
        if key == "meta":
            continue
        head, _, rest = key.partition("/")
        if head == "in":
            head, _, rest = rest.partition("/")
            head = "in_" + head
        groups.setdefault(head, OrderedDict())[rest] = data[key]
    return meta, groups


def oracle_config(meta):
    kw = dict(meta["kwargs"])
    common = dict(
        feature_size=meta["G"], latent_size=kw["latent_size"], hidden_sizes=kw["hidden_sizes"],
        reconstruction_distribution=kw["reconstruction_distribution"],
        number_of_importance_samples=meta["R"], number_of_monte_carlo_samples=meta["S"],
        minibatch_normalisation=kw.get("minibatch_normalisation", True),
        kl_weight=kw.get("kl_weight", 1.0),
        number_of_batches=kw.get("number_of_batches", 0) if kw.get("batch_correction") else 0,
        count_sum_feature=kw.get("count_sum", False),
        number_of_reconstruction_classes=kw.get("number_of_reconstruction_classes", 0))
    if meta["model"] == "GMVAE":
        return O.GMVAEConfig(
            number_of_latent_clusters=kw["number_of_latent_clusters"],
            prior_probabilities_method=kw.get("prior_probabilities_method", "uniform"),
            prior_probabilities=kw.get("prior_probabilities"),
            proportion_of_free_nats_for_y_kl_divergence=kw.get(
                "proportion_of_free_nats_for_y_kl_divergence", 0.0),
            dropout_keep_probabilities=kw.get("dropout_keep_probabilities"),
            latent_distribution=kw.get("latent_distribution", "gaussian mixture"), **common)
    return O.VAEConfig(
        latent_distribution=kw.get("latent_distribution", "gaussian"),
        # VAE:186-192: analytic KL by default only for the plain gaussian latent distribution
        analytical_kl_term=kw.get("analytical_kl_term",
                                  kw.get("latent_distribution", "gaussian") == "gaussian"),
        inference_architecture=kw.get("inference_architecture", "MLP"),
        generative_architecture=kw.get("generative_architecture", "MLP"),
        dropout_keep_probabilities=kw.get("dropout_keep_probabilities"), **common)


def oracle_inputs(meta, groups):
    """(params in the reference's creation order, x, eps, feature kwargs, dropout masks)."""
    params = OrderedDict((name, torch.as_tensor(groups["in_var"][name], dtype=D))
                         for name, _, _ in meta["variables"])
    feeds = groups["in_feed"]
    x = torch.as_tensor(feeds["X"], dtype=D)
    eps = [torch.as_tensor(v, dtype=D) for v in groups.get("in_eps", {}).values()]
    kw = meta["kwargs"]
    features = {}
    if kw.get("batch_correction"):
        features["batch_indices"] = torch.as_tensor(feeds["batch_indices"])
    if kw.get("count_sum"):
        features["count_sum_feature"] = torch.as_tensor(feeds["count_sum_feature"], dtype=D)
    if meta["model"] == "GMVAE":
        eps = torch.stack(eps)                       # (K, R*S, B, L)
        if kw["reconstruction_distribution"] == "constrained poisson":
            features["count_sum"] = torch.as_tensor(feeds["count_sum"], dtype=D)
        if "in_dropout" in groups:
            features["dropout"] = {"masks": {site: torch.as_tensor(m, dtype=D)
                                             for site, m in groups["in_dropout"].items()}}
    else:
        eps = eps[0] if eps else None                # (R*S, B, L)
        if kw["reconstruction_distribution"] == "constrained poisson":
            features["count_sum"] = torch.as_tensor(feeds["count_sum"], dtype=D)
        masks = {site: torch.as_tensor(m, dtype=D)
                 for site, m in groups.get("in_dropout", {}).items()}
        if masks:
            features["dropout"] = {"masks": masks}
    return params, x, eps, features


def close(got, want, what, rtol=RTOL, atol=ATOL):
    got = got.detach().numpy() if torch.is_tensor(got) else numpy.asarray(got)
    want = numpy.asarray(want)
    assert got.size == want.size, "{}: {} vs {}".format(what, got.shape, want.shape)
    got = got.reshape(want.shape)
    scale = max(1.0, float(numpy.abs(want).max())) if want.size else 1.0
    error = numpy.abs(got - want)
    assert numpy.all(error <= atol * scale + rtol * numpy.abs(want)), \
        "{}: max |diff| {:.3e} (scale {:.3e})".format(what, float(error.max()), scale)


def test_golden_cases_present():
    assert len(CASES) >= 25
    assert sum(c.startswith("gmvae") for c in CASES) >= 8


@pytest.mark.parametrize("name", CASES)
def test_variable_layout_matches_reference_graph(name):
    """Variable names, shapes, trainability and creation order of the reference's graph."""
    meta, groups = load_case(name)
    cfg = oracle_config(meta)
    init = O.gmvae_init_params if meta["model"] == "GMVAE" else O.vae_init_params
    params = init(cfg, seed=0, dtype=D)
    assert [(k, list(v.shape)) for k, v in params.items()] == \
        [(n, s) for n, s, _ in meta["variables"]]
    assert O.trainable_names(params) == [n for n, _, t in meta["variables"] if t]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_graph(name):
    meta, groups = load_case(name)
    cfg = oracle_config(meta)
    params, x, eps, features = oracle_inputs(meta, groups)
    feeds = groups["in_feed"]
    warm_up = float(feeds["warm_up_weight"])
    forward = O.gmvae_forward if meta["model"] == "GMVAE" else O.vae_forward
    extra = {} if meta["model"] == "GMVAE" else {
        "use_deterministic_z": meta["use_deterministic_z"]}
    out = forward(cfg, params, x, x, eps, is_training=meta["is_training"],
                  warm_up_weight=warm_up, moments=True, **extra, **features)
    for key, want in groups["out"].items():
        close(out[key], want, name + " " + key)

    if not meta["is_training"]:
        return
    # one optimiser step: raw gradients, batch-norm moving statistics, clip + Adam
    state = O.AdamState(params)
    if meta["adam_step"]:
        state.step = meta["adam_step"]
        for key in state.m:
            state.m[key] = torch.as_tensor(groups["in_adam_m"][key], dtype=D)
            state.v[key] = torch.as_tensor(groups["in_adam_v"][key], dtype=D)
    _, grads = O.train_step(cfg, params, state, x, x, eps,
                            float(feeds["learning_rate"]), warm_up_weight=warm_up, **extra,
                            **features)
    assert set(groups["grad"]) == set(state.m)
    for key, want in groups["grad"].items():
        got = grads[key] if grads[key] is not None else torch.zeros_like(params[key])
        close(got, want, name + " grad " + key, rtol=1e-8, atol=1e-9)
    assert set(groups["new"]) == {k for k in params
                                  if k in state.m or k.endswith(("moving_mean",
                                                                 "moving_variance"))}
    gmax = max(float(numpy.abs(g).max()) for g in groups["grad"].values())
    for key, want in groups["new"].items():
        got = params[key]
        if key in groups["grad"]:
            # a bias in front of a batch norm has an exactly-zero gradient; what either side
            # computes instead is rounding noise (~1e-16 gmax) that Adam's g / (|g| + 1e-8)
            # turns into a visible step -- leave those entries out
            live = torch.as_tensor(numpy.abs(groups["grad"][key]) > 1e-10 * gmax)
            got = torch.where(live, got, torch.as_tensor(want, dtype=D))
        close(got, want, name + " new " + key)


def test_reference_sample_and_update_order():
    """Order of the reference's sampling ops and batch-norm update ops (K-fold reuse)."""
    meta, _ = load_case("gmvae_nb_train")
    K = meta["kwargs"]["number_of_latent_clusters"]
    assert [c[0] for c in meta["sample_calls"]] == ["Categorical"] + ["Normal"] * (2 * K)
    # q(y|x) encoder once, then per k the q(z|x,y) encoder, then per k the decoder
    assert meta["bn_update_order"] == (["Y/CATEGORICAL/ENCODER/LAYER_1/BATCH_NORM"]
                                       + ["Z/Q/ENCODER/LAYER_1/BATCH_NORM"] * K
                                       + ["X/DECODER/LAYER_1/BATCH_NORM"] * K)
    meta, _ = load_case("vae_nb_dropout_train")
    assert meta["dropout_sites"] == ["ENCODER/1", "POSTERIOR/MU", "POSTERIOR/LOG_SIGMA",
                                     "DECODER/1", "X_TILDE/P", "X_TILDE/LOG_R"]
    # GMVAE: every build of a shared layer (one per cluster) is a dropout op of its own, the
    # one-hot input of the p(z|y) heads is dropped too (fourth keep probability)
    meta, groups = load_case("gmvae_nb_dropout_train")
    K, B, G = 3, meta["B"], meta["G"]
    sites = meta["dropout_sites"]
    assert sites[:3] == ["Y/CATEGORICAL/ENCODER/LAYER_1", "Y/CATEGORICAL/ENCODER/LAYER_2",
                         "Y/CATEGORICAL/LOGITS"]
    per_cluster = ["Z/Q/ENCODER/LAYER_1", "Z/Q/ENCODER/LAYER_2", "Z/Q/SOFTPLUS_GAUSSIAN/MEAN",
                   "Z/Q/SOFTPLUS_GAUSSIAN/SOFTPLUS_SCALE", "Z/P/SOFTPLUS_GAUSSIAN/MEAN",
                   "Z/P/SOFTPLUS_GAUSSIAN/SOFTPLUS_SCALE"]
    decoder = ["X/DECODER/LAYER_1", "X/DECODER/LAYER_2", "X/DISTRIBUTION/P",
               "X/DISTRIBUTION/LOG_R"]
    expected = list(sites[:3])
    for group in (per_cluster, decoder):
        for k in range(K):
            expected += [s if k == 0 else "{}#{}".format(s, k) for s in group]
    assert sites == expected and len(sites) == 3 + K * 10
    masks = groups["in_dropout"]
    assert masks["Z/Q/ENCODER/LAYER_1#2"].shape == (B, G + K)          # [x, e_k] together
    assert masks["Z/P/SOFTPLUS_GAUSSIAN/MEAN#1"].shape == (1, K)         # the one-hot y itself
    assert masks["X/DECODER/LAYER_1"].shape == (meta["R"] * meta["S"] * B, 3)


def test_standin_primitives_match_scipy():
    """The TFP closed forms restated in ``oracle/tf1_standin.py`` (the part of the golden vectors
    that is NOT the reference's own code) against scipy, independently of the oracle."""
    import scipy.stats
    from oracle import tf1_standin as T
    T.STATE.reset()
    x = torch.arange(0, 70, dtype=D)
    for rate in (0.05, 1.3, 17.0):
        got = T.Poisson(rate=torch.tensor(rate, dtype=D)).log_prob(x).numpy()
        assert numpy.allclose(got, scipy.stats.poisson.logpmf(x.numpy(), rate), rtol=1e-12,
                              atol=1e-12)
    for r in (0.2, 1.0, 9.5):
        for p in (0.03, 0.5, 0.97):
            nb = T.NegativeBinomial(total_count=torch.tensor(r, dtype=D),
                                    probs=torch.tensor(p, dtype=D))
            ref = scipy.stats.nbinom(r, 1.0 - p)      # tfp probs = 1 - scipy p
            assert numpy.allclose(nb.log_prob(x).numpy(), ref.logpmf(x.numpy()), rtol=1e-10,
                                  atol=1e-10)
            assert numpy.isclose(nb.mean().item(), ref.mean(), rtol=1e-12)
            assert numpy.isclose(nb.variance().item(), ref.var(), rtol=1e-12)
    a = T.Normal(torch.tensor([0.3, -1.2], dtype=D), torch.tensor([0.7, 2.0], dtype=D))
    b = T.Normal(torch.tensor([0.0, 0.5], dtype=D), torch.tensor([1.0, 0.4], dtype=D))
    z = torch.tensor([0.1, 0.9], dtype=D)
    assert numpy.allclose(a.log_prob(z).numpy(), scipy.stats.norm.logpdf(z.numpy(), [0.3, -1.2],
                                                                         [0.7, 2.0]))
    grid = torch.linspace(-40, 40, 400001, dtype=D).unsqueeze(-1)
    numeric = (torch.exp(a.log_prob(grid)) * (a.log_prob(grid) - b.log_prob(grid))).sum(0) * (
        grid[1, 0] - grid[0, 0])
    assert numpy.allclose(T.kl_divergence(a, b).numpy(), numeric.numpy(), rtol=1e-8)
    logits = torch.tensor([[0.2, -1.0, 3.0], [0.0, 0.0, 0.0]], dtype=D)
    cat = T.Categorical(logits=logits)
    probs = torch.softmax(logits, -1).numpy()
    assert numpy.allclose(cat.entropy().numpy(), [scipy.stats.entropy(p) for p in probs])
    assert numpy.allclose(cat.log_prob(torch.tensor([2, 1])).numpy(),
                          numpy.log([probs[0, 2], probs[1, 1]]))
    other = T.Categorical(logits=torch.tensor([[1.0, 1.0, -2.0]], dtype=D))
    q = torch.softmax(other.logits, -1).numpy()[0]
    assert numpy.allclose(T.kl_divergence(cat, other).numpy(),
                          [scipy.stats.entropy(p, q) for p in probs])


def test_standin_layers_follow_tf_contrib_semantics():
    """fully_connected / batch_norm / dropout / Adam of the stand-in on hand-computed cases."""
    from oracle import tf1_standin as T
    x = numpy.array([[1.0, 2.0], [3.0, 6.0], [5.0, 1.0]])
    T.STATE.reset(initial={"L/DENSE/weights": numpy.array([[1.0, 0.0], [0.0, 2.0]]),
                           "L/DENSE/biases": numpy.array([0.5, -1.0])})
    with T.variable_scope("L"):
        y = T.fully_connected(torch.as_tensor(x), 2, scope="DENSE")
        assert numpy.allclose(y.detach().numpy(), x * [1.0, 2.0] + [0.5, -1.0])
        out = T.batch_norm(y, is_training=True, scope="BATCH_NORM")
    mean, var = y.detach().numpy().mean(0), y.detach().numpy().var(0)
    assert numpy.allclose(out.detach().numpy(), (y.detach().numpy() - mean) / numpy.sqrt(var + 1e-3))
    new_mean = T.STATE.updates["L/BATCH_NORM/moving_mean"].numpy()
    new_var = T.STATE.updates["L/BATCH_NORM/moving_variance"].numpy()
    assert numpy.allclose(new_mean, 0.001 * mean)                       # 0 - (0 - mean)(1 - .999)
    assert numpy.allclose(new_var, 1.0 - 0.001 * (1.0 - var * 3 / 2))   # Bessel-corrected
    # Adam, first step from zero slots: theta - lr * g / (|g| + eps * sqrt(1 - b2)) ~ lr sign(g)
    w = T.STATE.variables["L/DENSE/weights"]
    opt = T.AdamOptimizer(0.01)
    pairs = opt.compute_gradients((w * torch.tensor([[2.0, -3.0], [0.5, 0.0]])).sum())
    opt.apply_gradients([(T.clip_by_value(g, -1.0, 1.0), v) for g, v in pairs])
    step = w.detach().numpy() - T.STATE.updates["L/DENSE/weights"].numpy()
    assert numpy.allclose(step, 0.01 * numpy.array([[1.0, -1.0], [1.0, 0.0]]), atol=1e-8)


@pytest.mark.parametrize("name", CASES)
def test_fixtures_are_reachable_in_fp32(name):
    """The fixtures are fp64; the CUDA engines compute in fp32 and are held to them at 2e-4
    (tests/test_zz_gpu_reference_graph.py).  The oracle run in fp32 must land well inside that,
    otherwise a fixture sits in an ill-conditioned corner (saturated sigmoids, values on a clip
    boundary) and says nothing about an fp32 implementation."""
    if "clipped" in name:
        pytest.skip("drives the heads onto their clips on purpose (fp64 comparison only)")
    meta, groups = load_case(name)
    cfg = oracle_config(meta)
    params, x, eps, features = oracle_inputs(meta, groups)
    f32 = lambda v: v.float() if torch.is_tensor(v) and v.is_floating_point() else v  # noqa: E731
    params = OrderedDict((k, f32(v)) for k, v in params.items())
    features = {k: ({"masks": {s: f32(m) for s, m in v["masks"].items()}} if k == "dropout"
                    else f32(v)) for k, v in features.items()}
    forward = O.gmvae_forward if meta["model"] == "GMVAE" else O.vae_forward
    extra = {} if meta["model"] == "GMVAE" else {
        "use_deterministic_z": meta["use_deterministic_z"]}
    out = forward(cfg, params, x.float(), x.float(), f32(eps), is_training=meta["is_training"],
                  warm_up_weight=float(groups["in_feed"]["warm_up_weight"]), moments=True,
                  **extra, **features)
    for key, want in groups["out"].items():
        got = out[key].detach().double().numpy().reshape(-1)
        want = want.reshape(-1)
        scale = max(float(numpy.abs(want).max()), 0.05)
        assert numpy.abs(got - want).max() <= 5e-5 * scale, (key, numpy.abs(got - want).max(), scale)


def _engine_for(meta, device="cpu"):
    kw = meta["kwargs"]
    extras = dict(
        number_of_batches=kw.get("number_of_batches", 0) if kw.get("batch_correction") else 0,
        count_sum_feature=kw.get("count_sum", False),
        number_of_reconstruction_classes=kw.get("number_of_reconstruction_classes", 0))
    if meta["model"] == "GMVAE":
        from scvae_b200.gmvae_engine import GMVAEEngine
        return GMVAEEngine(
            meta["G"], kw["latent_size"], kw["number_of_latent_clusters"], kw["hidden_sizes"],
            kw["reconstruction_distribution"], kw.get("minibatch_normalisation", True),
            kw.get("kl_weight", 1.0), kw.get("prior_probabilities_method", "uniform"),
            kw.get("prior_probabilities"),
            kw.get("proportion_of_free_nats_for_y_kl_divergence", 0.0), device=device,
            tensor_cores=False, latent_distribution=kw.get("latent_distribution", "gaussian mixture"),
            **extras)
    from scvae_b200.engine import VAEEngine
    return VAEEngine(
        meta["G"], kw["latent_size"], kw["hidden_sizes"], kw["reconstruction_distribution"],
        kw.get("latent_distribution", "gaussian"), kw.get("minibatch_normalisation", True),
        kl_weight=kw.get("kl_weight", 1.0), device=device, tensor_cores=False,
        inference_architecture=kw.get("inference_architecture", "MLP"),
        generative_architecture=kw.get("generative_architecture", "MLP"), **extras)


@pytest.mark.parametrize("name", CASES)
def test_engine_parameter_layout_round_trips_reference_variables(name):
    """Host logic of the product, no kernels involved: the engines' flat parameter store
    (transposed, bias-augmented, head-concatenated, class-major P_K, split one-hot rows) must
    take the reference's variables by their TF names and give back exactly the same names,
    shapes and values (fp32)."""
    meta, groups = load_case(name)
    engine = _engine_for(meta)
    variables = OrderedDict((n, torch.as_tensor(groups["in_var"][n], dtype=torch.float32))
                            for n, _, _ in meta["variables"])
    engine.import_parameters(variables)
    exported = engine.export_parameters()
    assert sorted(exported) == sorted(variables)
    for key, value in variables.items():
        assert tuple(exported[key].shape) == tuple(value.shape), key
        assert torch.equal(exported[key].float().cpu(), value), key
    gradients = engine.export_gradients()
    assert sorted(gradients) == sorted(n for n, _, t in meta["variables"] if t)
    for key, value in gradients.items():
        assert tuple(value.shape) == tuple(variables[key].shape), key
