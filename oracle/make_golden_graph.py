"""Golden vectors from the REFERENCE'S OWN graph code (test infrastructure, build container only).

Imports ``scvae/models/variational_autoencoder.py``,
``scvae/models/gaussian_mixture_variational_autoencoder.py``, ``scvae/models/utilities.py`` and
``scvae/distributions/*.py`` unmodified from ``/root/reference`` with ``oracle/tf1_standin.py``
installed under the TensorFlow / TFP module names, constructs the reference's model classes on
injected inputs (one construction = one ``session.run``; see the stand-in's docstring for what
this does and does not pin) and writes one ``tests/golden/reference_graph_<case>.npz`` per
case: inputs (``in/...``: variables by TF name, x, eps, dropout masks, feeds) and what the
reference's graph produced (``out/...``: fetched tensors, ``grad/...``: raw gradients of
``-lower_bound_weighted``, ``new/...``: variables after the batch-norm updates and the
clip + Adam step).

Run:  python oracle/make_golden_graph.py        (needs /root/reference; outputs are committed)
"""

import importlib
import json
import os
import sys
import types

import numpy
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf1_standin  # noqa: E402

REFERENCE = os.environ.get("SCVAE_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

G, B = 24, 6      # genes, cells per minibatch of every case


def import_reference_classes():
    from make_golden import _MockFinder
    tf1_standin.install()
    finder = _MockFinder()
    finder.ROOTS = ("loompy", "tables", "matplotlib", "seaborn", "mpl_toolkits")
    sys.meta_path.insert(0, finder)
    resources = types.ModuleType("importlib_resources")
    resources.open_text = lambda package, name: open(
        os.path.join(REFERENCE, package.replace(".", os.sep), name))
    sys.modules["importlib_resources"] = resources
    for name, directory in (("scvae", "scvae"), ("scvae.data", "scvae/data")):
        package = types.ModuleType(name)
        package.__path__ = [os.path.join(REFERENCE, directory)]
        sys.modules[name] = package
    data_set = types.ModuleType("scvae.data.data_set")
    data_set.DataSet = object
    sys.modules["scvae.data.data_set"] = data_set
    return {
        "VAE": importlib.import_module(
            "scvae.models.variational_autoencoder").VariationalAutoencoder,
        "GMVAE": importlib.import_module(
            "scvae.models.gaussian_mixture_variational_autoencoder"
        ).GaussianMixtureVariationalAutoencoder,
    }


def counts(rng, rows, genes, gentle=False, outliers=True):
    """Small zero-heavy counts with a few large values (exercises lgamma far from the origin).
    ``gentle``: counts below ten only -- without batch norm nothing rescales the activations and
    large counts saturate the sigmoid heads to exactly one, where the reference itself yields
    NaN."""
    rate = rng.gamma(0.6, 4.0, size=(1, genes))
    x = rng.poisson(rate * rng.gamma(2.0, 0.5, size=(rows, 1))).astype(numpy.float64)
    x[rng.uniform(size=x.shape) < 0.4] = 0.0
    if gentle:
        return numpy.minimum(x, 9.0)
    if not outliers:
        return x
    x[0, 0] = 157.0
    x[rows - 1, genes - 1] = 1203.0
    return x


# name, model, constructor kwargs, run options
CASES = [
    # BASELINE configs[0] at its real shape: 100 genes, Poisson, latent 10, hidden [100] (the
    # defaults), one minibatch of 100 cells (defaults.json:51) of the reference's own synthetic
    # `development`-style counts
    ("vae_c1_poisson_train", "VAE", dict(reconstruction_distribution="poisson", latent_size=10,
                                         hidden_sizes=[100]), dict(G=100, B=100, outliers=False, scale={"POSTERIOR/LOG_SIGMA": 0.3})),
    ("vae_c1_poisson_eval", "VAE", dict(reconstruction_distribution="poisson", latent_size=10,
                                        hidden_sizes=[100]),
     dict(G=100, B=100, outliers=False, scale={"POSTERIOR/LOG_SIGMA": 0.3},
          is_training=False)),
    ("vae_poisson_train", "VAE", dict(reconstruction_distribution="poisson"), dict()),
    ("vae_nb_train", "VAE", dict(reconstruction_distribution="negative binomial"), dict()),
    ("vae_nb_eval", "VAE", dict(reconstruction_distribution="negative binomial"),
     dict(is_training=False)),
    ("vae_nb_eval_deterministic", "VAE", dict(reconstruction_distribution="negative binomial"),
     dict(is_training=False, use_deterministic_z=True)),
    # training steps WITHOUT sampling noise (z = q_z_mean while is_training): nothing random is
    # left in the graph, so a machine with TensorFlow 1.15 can replay these through the real
    # reference and pin the TF / TFP primitives too (oracle/check_with_tensorflow.py)
    ("vae_nb_train_deterministic", "VAE", dict(reconstruction_distribution="negative binomial"),
     dict(use_deterministic_z=True)),
    ("vae_zinb_train_deterministic", "VAE",
     dict(reconstruction_distribution="zero-inflated negative binomial", hidden_sizes=[8, 5]),
     dict(use_deterministic_z=True)),
    ("vae_poisson_eval_deterministic", "VAE", dict(reconstruction_distribution="poisson"),
     dict(is_training=False, use_deterministic_z=True)),
    ("vae_zip_train", "VAE", dict(reconstruction_distribution="zero-inflated poisson"), dict()),
    ("vae_zinb_train", "VAE", dict(reconstruction_distribution="zero-inflated negative binomial",
                                   hidden_sizes=[8, 5]), dict()),
    ("vae_zinb_eval_iw", "VAE", dict(reconstruction_distribution="zero-inflated negative binomial"),
     dict(is_training=False, R=3, S=2)),
    ("vae_nb_train_iw_warmup", "VAE", dict(reconstruction_distribution="negative binomial",
                                            kl_weight=0.7), dict(R=2, S=3, warm_up_weight=0.25)),
    ("vae_nb_train_second_step", "VAE", dict(reconstruction_distribution="negative binomial"),
     dict(adam_step=4)),
    ("vae_nb_no_bn_train", "VAE", dict(reconstruction_distribution="negative binomial",
                                        minibatch_normalisation=False), dict(gentle=True)),
    ("vae_constrained_poisson_train", "VAE",
     dict(reconstruction_distribution="constrained poisson"), dict()),
    ("vae_nb_k3_train", "VAE", dict(reconstruction_distribution="negative binomial",
                                     number_of_reconstruction_classes=3), dict()),
    ("vae_poisson_k2_eval", "VAE", dict(reconstruction_distribution="poisson",
                                         number_of_reconstruction_classes=2),
     dict(is_training=False, R=2, S=2)),
    ("vae_nb_bc_count_sum_train", "VAE", dict(reconstruction_distribution="negative binomial",
                                               batch_correction=True, number_of_batches=3,
                                               count_sum=True), dict()),
    ("vae_nb_lfm_inference_train", "VAE", dict(reconstruction_distribution="negative binomial",
                                                inference_architecture="LFM"),
     # raw counts feed the posterior heads directly: small counts, small head weights
     dict(small_counts=True, scale={"POSTERIOR": 0.05})),
    ("vae_nb_lfm_generative_train", "VAE", dict(reconstruction_distribution="negative binomial",
                                                 generative_architecture="LFM"),
     # z feeds the likelihood heads directly: keep sigma = exp(log_sigma) moderate
     dict(scale={"POSTERIOR": 0.3})),
    ("vae_nb_sampled_kl_train", "VAE", dict(reconstruction_distribution="negative binomial",
                                             analytical_kl_term=False), dict(R=2, S=2)),
    ("vae_nb_sampled_kl_eval_deterministic", "VAE",
     dict(reconstruction_distribution="negative binomial", analytical_kl_term=False),
     dict(is_training=False, use_deterministic_z=True)),
    ("vae_nb_sampled_kl_eval_iw", "VAE",
     dict(reconstruction_distribution="negative binomial", analytical_kl_term=False,
          latent_distribution="unit-variance gaussian"), dict(is_training=False, R=3, S=2)),
    ("vae_nb_unit_variance_train", "VAE", dict(reconstruction_distribution="negative binomial",
                                                latent_distribution="unit-variance gaussian"),
     dict()),
    ("vae_nb_dropout_train", "VAE", dict(reconstruction_distribution="negative binomial",
                                          dropout_keep_probabilities=[0.8, 0.9, 0.7]), dict()),
    ("vae_zinb_dropout_deep_train", "VAE",
     dict(reconstruction_distribution="zero-inflated negative binomial", hidden_sizes=[8, 5],
          batch_correction=True, number_of_batches=3, count_sum=True,
          dropout_keep_probabilities=[0.8, 0.9, 0.7]), dict(R=1, S=2)),
    ("vae_poisson_k2_dropout_train", "VAE",
     dict(reconstruction_distribution="poisson", number_of_reconstruction_classes=2,
          latent_distribution="unit-variance gaussian", analytical_kl_term=True,
          dropout_keep_probabilities=0.75), dict(R=2, S=1)),
    # large head weights: log_lambda / log_r reach their +-10 clips, log_sigma its +-3 clips
    # (zero gradient through a clipped unit)
    ("vae_poisson_clipped_train", "VAE", dict(reconstruction_distribution="poisson"),
     dict(scale={"X_TILDE": 12.0, "POSTERIOR/LOG_SIGMA": 12.0})),
    ("vae_nb_clipped_train", "VAE", dict(reconstruction_distribution="negative binomial"),
     dict(scale={"X_TILDE/LOG_R": 15.0, "POSTERIOR/LOG_SIGMA": 12.0})),
    ("gmvae_nb_train", "GMVAE", dict(reconstruction_distribution="negative binomial",
                                      number_of_latent_clusters=3), dict()),
    ("gmvae_nb_eval", "GMVAE", dict(reconstruction_distribution="negative binomial",
                                     number_of_latent_clusters=3), dict(is_training=False)),
    ("gmvae_zinb_train_mc", "GMVAE",
     dict(reconstruction_distribution="zero-inflated negative binomial",
          number_of_latent_clusters=4, hidden_sizes=[8, 5]), dict(R=2, S=2, warm_up_weight=0.5)),
    ("gmvae_poisson_learn_train", "GMVAE", dict(reconstruction_distribution="poisson",
                                                 number_of_latent_clusters=3,
                                                 prior_probabilities_method="learn"), dict()),
    ("gmvae_nb_custom_prior_train", "GMVAE",
     dict(reconstruction_distribution="negative binomial", number_of_latent_clusters=3,
          prior_probabilities_method="custom", prior_probabilities=[0.5, 0.3, 0.2]), dict()),
    ("gmvae_nb_free_nats_train", "GMVAE",
     dict(reconstruction_distribution="negative binomial", number_of_latent_clusters=3,
          proportion_of_free_nats_for_y_kl_divergence=0.9), dict()),
    ("gmvae_nb_k2_train", "GMVAE", dict(reconstruction_distribution="negative binomial",
                                         number_of_latent_clusters=2,
                                         number_of_reconstruction_classes=2), dict()),
    ("gmvae_nb_bc_count_sum_train", "GMVAE",
     dict(reconstruction_distribution="negative binomial", number_of_latent_clusters=2,
          batch_correction=True, number_of_batches=2, count_sum=True), dict()),
    # every build of a shared layer is a dropout op of its own (one mask per cluster), and the
    # GMVAE takes a fourth keep probability for the one-hot input of the p(z|y) heads
    ("gmvae_nb_dropout_train", "GMVAE",
     dict(reconstruction_distribution="negative binomial", number_of_latent_clusters=3,
          hidden_sizes=[8, 5], dropout_keep_probabilities=[0.8, 0.9, 0.7, 0.6]), dict(R=1, S=2)),
    # multivariate-Gaussian q(z|x,y) / p(z|y) with fill_triangular scale matrices (DU:75-93,
    # multivariate_normal.py:90-150)
    ("gmvae_nb_full_covariance_train", "GMVAE",
     dict(reconstruction_distribution="negative binomial", number_of_latent_clusters=3,
          latent_distribution="full-covariance gaussian mixture"), dict(R=1, S=2)),
    ("gmvae_poisson_full_covariance_eval", "GMVAE",
     dict(reconstruction_distribution="poisson", number_of_latent_clusters=2,
          latent_distribution="full-covariance gaussian mixture"), dict(is_training=False, R=2, S=1)),
    ("gmvae_constrained_poisson_train", "GMVAE",
     dict(reconstruction_distribution="constrained poisson", number_of_latent_clusters=3),
     dict(R=1, S=2)),
    ("gmvae_constrained_poisson_eval", "GMVAE",
     dict(reconstruction_distribution="constrained poisson", number_of_latent_clusters=3),
     dict(is_training=False, R=2, S=1)),
    ("gmvae_nb_no_bn_eval", "GMVAE", dict(reconstruction_distribution="negative binomial",
                                           number_of_latent_clusters=3,
                                           minibatch_normalisation=False),
     dict(is_training=False, R=2, S=1, gentle=True)),
]

VAE_FETCHES = ["lower_bound", "lower_bound_weighted", "reconstruction_error", "kl_divergence",
               "kl_divergence_neurons", "q_z_mean", "z", "p_x_mean", "p_x_stddev",
               "stddev_of_p_x_given_z_mean"]
GMVAE_FETCHES = ["lower_bound", "lower_bound_weighted", "reconstruction_error", "kl_divergence",
                 "kl_divergence_z", "kl_divergence_y", "kl_divergence_neurons",
                 "kl_divergence_z_neurons", "z_mean", "y", "q_y_logits", "q_y_probabilities",
                 "p_y_probabilities", "p_x_mean", "p_x_stddev", "stddev_of_p_x_given_z_mean"]


def discover_variables(cls, kwargs, feeds, G=G):
    """First construction with default initialisers: the variable names and shapes the
    reference's graph code creates, in creation order."""
    tf1_standin.STATE.reset(feeds=feeds, seed=11)
    cls(feature_size=G, log_directory="log", **kwargs)
    state = tf1_standin.STATE
    return [(name, tuple(value.shape), name in state.trainable)
            for name, value in state.variables.items()]


def randomised_variables(rng, layout):
    """Non-trivial values for every variable (zero biases / unit moving variances would hide
    mistakes): Xavier-scaled weights, small biases / beta / moving means, moving variances
    around one."""
    values = {}
    for name, shape, _ in layout:
        if name == "global_step":
            continue
        if name.endswith("/weights"):
            limit = numpy.sqrt(6.0 / sum(shape))
            values[name] = rng.uniform(-limit, limit, size=shape)
        elif name.endswith("moving_variance"):
            values[name] = rng.uniform(0.5, 1.5, size=shape)
        else:
            values[name] = 0.1 * rng.standard_normal(size=shape)
    return values


def run_case(classes, name, model, kwargs, options, seed):
    rng = numpy.random.RandomState(seed)
    kwargs = dict(kwargs)
    kwargs.setdefault("latent_size", 3)
    kwargs.setdefault("hidden_sizes", [8])
    G, B = options.get("G", globals()["G"]), options.get("B", globals()["B"])
    R, S = options.get("R", 1), options.get("S", 1)
    is_training = options.get("is_training", True)
    deterministic = options.get("use_deterministic_z", False)
    x = counts(rng, B, G, options.get("gentle", False) or options.get("small_counts", False),
               options.get("outliers", True))
    feeds = {
        "X": x, "T": x, "learning_rate": 1e-3,
        "warm_up_weight": options.get("warm_up_weight", 1.0),
        "is_training": is_training, "use_deterministic_z": deterministic,
        "number_of_iw_samples": R, "number_of_mc_samples": S, "sample_size": 2,
        "batch_indices": rng.randint(0, kwargs.get("number_of_batches") or 1, size=(B, 1)),
        "count_sum_feature": rng.uniform(0.2, 1.0, size=(B, 1)),
        "count_sum": x.sum(axis=1, keepdims=True),
    }
    cls = classes[model]
    layout = discover_variables(cls, kwargs, feeds, G)
    variables = randomised_variables(rng, layout)
    if options.get("gentle"):
        variables = {k: (0.3 * v if k.endswith("/weights") else v) for k, v in variables.items()}
    for prefix, factor in options.get("scale", {}).items():
        variables = {k: (factor * v if k.startswith(prefix) and k.endswith("/weights") else v)
                     for k, v in variables.items()}
    K = kwargs.get("number_of_latent_clusters", 1) if model == "GMVAE" else 1
    n_noise = 0 if deterministic else K
    noise = [rng.standard_normal(size=(R * S, B, kwargs["latent_size"])) for _ in range(n_noise)]
    if not is_training and kwargs.get("minibatch_normalisation", True):
        # moving statistics of a trained model track the activations: take them from the batch
        # statistics of a training-mode pass (slightly perturbed), not from thin air
        tf1_standin.STATE.reset(feeds=dict(feeds, is_training=True), initial=variables,
                                noise=[n.copy() for n in noise], seed=seed)
        cls(feature_size=G, log_directory="log", **kwargs)
        for path, stats in tf1_standin.STATE.batch_statistics.items():
            mean = sum(m for m, _ in stats).numpy() / len(stats)
            variance = sum(v for _, v in stats).numpy() / len(stats)
            variables[path + "/moving_mean"] = mean + 0.05 * rng.standard_normal(mean.shape)
            variables[path + "/moving_variance"] = variance * rng.uniform(0.9, 1.1, mean.shape)
    initial = dict(variables)
    adam_step = options.get("adam_step", 0)
    slots_m, slots_v = {}, {}
    if adam_step:      # a later optimiser step: non-zero Adam slots and step counter
        for vname, shape, trainable in layout:
            if trainable:
                slots_m[vname] = 0.01 * rng.standard_normal(size=shape)
                slots_v[vname] = 1e-4 * rng.uniform(0.1, 1.0, size=shape)
        initial.update({"__adam_m__": slots_m, "__adam_v__": slots_v,
                        "__adam_step__": adam_step})
    # dropout: discover the sites (scopes) with generated masks, then re-run with them injected
    tf1_standin.STATE.reset(feeds=feeds, initial=initial, noise=[n.copy() for n in noise],
                            seed=seed)
    net = cls(feature_size=G, log_directory="log", **kwargs)
    state = tf1_standin.STATE
    created = [n for n in state.variables if n != "global_step"]
    assert sorted(created) == sorted(variables), (sorted(created), sorted(variables))
    assert not state.noise, "unused injected noise"
    masks = {site: mask.numpy().copy() for site, mask in state.dropout_masks.items()}

    record = {}
    for key, value in variables.items():
        record["in/var/" + key] = value
    for key, value in feeds.items():
        record["in/feed/" + key] = numpy.asarray(value)
    for k, eps in enumerate(noise):
        record["in/eps/{}".format(k)] = eps
    for site, mask in masks.items():
        record["in/dropout/" + site] = mask
    for key, value in slots_m.items():
        record["in/adam_m/" + key] = value
        record["in/adam_v/" + key] = slots_v[key]
    for fetch in (VAE_FETCHES if model == "VAE" else GMVAE_FETCHES):
        value = getattr(net, fetch)
        record["out/" + fetch] = value.detach().numpy() if torch.is_tensor(value) \
            else numpy.asarray(value)
    if is_training:
        for key, value in state.gradients.items():
            record["grad/" + key] = value.numpy()
        for key, value in state.updates.items():
            if not key.startswith("__") and key != "global_step":
                record["new/" + key] = value.numpy()
    meta = {"model": model, "kwargs": kwargs, "G": G, "B": B, "R": R, "S": S, "is_training": is_training,
            "use_deterministic_z": deterministic, "adam_step": adam_step,
            "variables": [[n, list(s), bool(t)] for n, s, t in layout if n != "global_step"],
            "sample_calls": [[n, list(s)] for n, s in state.sample_calls],
            "bn_update_order": list(state.collections[tf1_standin.GraphKeys.UPDATE_OPS]),
            "dropout_sites": list(masks)}
    record["meta"] = numpy.array(json.dumps(meta, sort_keys=True))
    numpy.savez_compressed(os.path.join(OUT, "reference_graph_{}.npz".format(name)), **record)
    print("{:34s} ELBO {: .6f}  vars {:2d}  samples {}".format(
        name, float(record["out/lower_bound"]), len(variables),
        [c[0] for c in meta["sample_calls"]]))


def main():
    os.makedirs(OUT, exist_ok=True)
    classes = import_reference_classes()
    import zlib
    for name, model, kwargs, options in CASES:
        if sys.argv[1:] and name not in sys.argv[1:]:
            continue
        # seeded by the case name: adding a case leaves the others' fixtures unchanged
        run_case(classes, name, model, kwargs, options, seed=zlib.crc32(name.encode()) % 2 ** 31)


if __name__ == "__main__":
    main()
