"""The reference-facing model class on the GPU: train / resume / evaluate / sample, on-disk
contract (checkpoint naming, event tags) and epoch-level parity with the oracle."""
import os

import numpy
import pytest
import scipy.sparse
import torch

from oracle import scvae_oracle as O

pytestmark = pytest.mark.gpu


def _data(n=300, g=60, seed=3):
    from scvae_b200.data_set import DataSet
    x, labels = O.synthetic_counts(n, g, n_types=3, seed=seed, target_zero_fraction=0.8)
    x = numpy.minimum(x, 50.0)
    return DataSet("toy", values=scipy.sparse.csr_matrix(x), labels=labels.astype(str))


def test_train_resume_evaluate_sample(tmp_path):
    from scvae_b200.variational_autoencoder import VariationalAutoencoder
    from scvae_b200 import model_utilities as MU
    full = _data()
    training, validation, test = full.split()
    model = VariationalAutoencoder(
        feature_size=60, latent_size=4, hidden_sizes=[32], reconstruction_distribution="negative binomial",
        number_of_warm_up_epochs=2, log_directory=str(tmp_path), seed=1)
    assert model.train(training, validation, number_of_epochs=3, minibatch_size=50,
                       learning_rate=1e-2, shuffle_seed=0) == 0
    directory = model.log_directory()
    assert open(os.path.join(directory, "checkpoint")).readline().strip() == 'model_checkpoint_path: "model.ckpt-3"'
    assert os.path.exists(os.path.join(directory, "model.ckpt-3.pt"))
    assert not os.path.exists(os.path.join(directory, "model.ckpt-2.pt"))      # max_to_keep=1
    assert os.path.isdir(os.path.join(directory, "best"))
    assert model.has_been_trained()
    curves = MU.load_learning_curves(model, ["training", "validation"])
    assert len(curves["training"]["lower_bound"]) == 3
    assert curves["training"]["lower_bound"][-1] > curves["training"]["lower_bound"][0]
    assert MU.load_number_of_epochs_trained(model) == 3
    kl = MU.load_kl_divergences(model, "training")
    assert kl.shape == (3, 4)
    assert os.path.exists(os.path.join(directory, "metadata_log-0-3.log"))
    # resume: two more epochs, learning curve continues at step 4
    assert model.train(training, validation, number_of_epochs=5, minibatch_size=50,
                       learning_rate=1e-2, shuffle_seed=1) == 0
    assert MU.load_number_of_epochs_trained(model) == 5
    assert model.train(training, validation, number_of_epochs=5, minibatch_size=50) == 0  # no-op

    transformed, reconstructed, latent = model.evaluate(
        test, minibatch_size=64, evaluation_subset_indices={0, 5, 7}, output_versions="all")
    assert transformed is test
    assert reconstructed.values.shape == (test.number_of_examples, 60)
    assert reconstructed.total_standard_deviations[5].toarray().max() > 0
    assert reconstructed.total_standard_deviations[1].toarray().max() == 0
    assert latent["z"].values.shape == (test.number_of_examples, 4)
    assert numpy.isfinite(reconstructed.values).all()
    ev = MU.load_learning_curves(model, "evaluation")
    assert ev["lower_bound"] is not None
    only_latent = model.evaluate(test, output_versions="latent", use_best_model=True)
    assert set(only_latent) == {"z"}
    sample_set, sample_latent = model.sample(sample_size=37, minibatch_size=16)
    assert sample_set.values.shape == (37, 60) and sample_latent["z"].values.shape == (37, 4)
    untrained = VariationalAutoencoder(feature_size=60, latent_size=3, log_directory=str(tmp_path))
    with pytest.raises(Exception, match="not been trained"):
        untrained.evaluate(test)


def test_one_epoch_matches_oracle(tmp_path):
    """train() for one epoch (2 steps) == oracle with the same permutation and noise."""
    from scvae_b200.variational_autoencoder import VariationalAutoencoder
    from scvae_b200 import kernels as K
    ds = _data(n=200, g=48, seed=9)
    model = VariationalAutoencoder(
        feature_size=48, latent_size=3, hidden_sizes=[16], reconstruction_distribution="poisson",
        log_directory=str(tmp_path), seed=4, tensor_cores=False)
    model.train(ds, number_of_epochs=1, minibatch_size=100, learning_rate=1e-3, shuffle_seed=5,
                noise_seed=21, use_cuda_graph=False)
    cfg = O.VAEConfig(48, 3, [16], "poisson")
    params = O.vae_init_params(cfg, seed=4, dtype=torch.float64)
    state = O.AdamState(params)
    perm = numpy.random.RandomState(5).permutation(200)
    x = torch.tensor(ds.values.toarray(), dtype=torch.float64)
    for step in range(2):
        eps = torch.zeros(100, 3, device="cuda")
        K.fill_normal(eps, 21, 0, torch.tensor([step], dtype=torch.int64, device="cuda"))
        xb = x[perm[step * 100:(step + 1) * 100]]
        O.train_step(cfg, params, state, xb, xb, eps.cpu().double().reshape(1, 100, 3), 1e-3)
    got = model._get_engine().export_parameters()
    for k, v in params.items():
        if k.endswith("DENSE/biases") and k.replace("DENSE/biases", "BATCH_NORM/beta") in params:
            continue   # zero-gradient bias in front of a batch norm: Adam amplifies fp32 noise
        assert (got[k].double() - v).abs().max().item() <= 2e-5 * max(v.abs().max().item(), 1.0), k
    # the logged training ELBO is the eval-mode pass with the reference's N/B divisor (A.7)
    from scvae_b200 import model_utilities as MU
    logged = MU.load_learning_curves(model, "training")["lower_bound"][0]
    assert numpy.isfinite(logged)


def test_gmvae_train_evaluate_sample(tmp_path):
    from scvae_b200.gaussian_mixture_variational_autoencoder import (
        GaussianMixtureVariationalAutoencoder)
    from scvae_b200 import model_utilities as MU
    full = _data(n=240, g=40, seed=4)
    training, validation, test = full.split()
    model = GaussianMixtureVariationalAutoencoder(
        feature_size=40, latent_size=3, hidden_sizes=[24], number_of_latent_clusters=3,
        reconstruction_distribution="zero-inflated negative binomial", log_directory=str(tmp_path),
        seed=2)
    assert model.type == "GMVAE" and "c_3" in model.name
    assert model.train(training, validation, number_of_epochs=2, minibatch_size=48,
                       learning_rate=5e-3, shuffle_seed=0) == 0
    curves = MU.load_learning_curves(model, ["training", "validation"])
    assert len(curves["training"]["kl_divergence_y"]) == 2
    assert MU.load_accuracies(model, "training") is not None
    centroids = MU.load_centroids(model, "validation")
    assert centroids["posterior"]["means"].shape == (2, 3, 3)
    transformed, reconstructed, latent = model.evaluate(test, minibatch_size=32)
    assert set(latent) == {"z", "y"}
    assert latent["y"].values.shape == (test.number_of_examples, 3)
    assert numpy.allclose(latent["y"].values.sum(1), 1, atol=1e-5)
    assert reconstructed.values.shape == (test.number_of_examples, 40)
    assert latent["z"].predicted_cluster_ids.shape == (test.number_of_examples,)
    sample_set, sample_latent = model.sample(sample_size=20, minibatch_size=8)
    assert sample_set.values.shape == (20, 40) and sample_latent["y"].values.shape == (20, 3)


def test_cli_train_and_evaluate_on_tsv(tmp_path):
    """BASELINE config C1 in miniature: TSV count matrix -> `scvae train` -> `scvae evaluate`."""
    import pandas
    from scvae_b200 import cli
    ds = _data(n=120, g=30, seed=6)
    frame = pandas.DataFrame(ds.values.toarray().astype(int),
                             index=["cell {}".format(i) for i in range(120)],
                             columns=["gene {}".format(j) for j in range(30)])
    path = tmp_path / "counts.tsv"
    frame.to_csv(path, sep="\t")
    models = str(tmp_path / "models")
    assert cli.main(["train", str(path), "-r", "poisson", "-l", "3", "-H", "16", "-e", "2",
                     "-B", "40", "-M", models, "--split-data-set"]) == 0
    assert cli.main(["evaluate", str(path), "-r", "poisson", "-l", "3", "-H", "16", "-B", "40",
                     "-M", models, "--split-data-set"]) == 0
    found = [f for _, _, files in os.walk(models) for f in files]
    assert "checkpoint" in found and any(f.startswith("model.ckpt-2") for f in found)


def test_peer_exchange_matches_nccl_on_two_gpus():
    """Fused peer-memory exchange + optimiser vs NCCL all-reduce + replicated Adam
    (tools/dp_check.py under torchrun; needs two GPUs on the box)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
                        "29541", os.path.join(root, "tools", "dp_check.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "dp_check OK" in r.stdout, r.stdout[-3000:]


def test_train_evaluate_with_batch_correction_count_sum_and_lfm(tmp_path):
    """The `--batch-correction`, `--count-sum` and `--generative-architecture LFM` options of the
    reference CLI through the model class (VAE:2400-2462)."""
    from scvae_b200.data_set import DataSet
    from scvae_b200.variational_autoencoder import VariationalAutoencoder
    x, labels = O.synthetic_counts(240, 64, n_types=3, seed=8, target_zero_fraction=0.8)
    x = numpy.minimum(x, 50.0)
    batches = (numpy.arange(240) % 3).reshape(-1, 1)
    full = DataSet("toy", values=scipy.sparse.csr_matrix(x), labels=labels.astype(str),
                   batch_indices=batches, batch_names=["a", "b", "c"])
    training, validation, test = full.split()
    assert training.batch_indices is not None
    for kwargs in (dict(batch_correction=True, number_of_batches=3, count_sum=True),
                   dict(generative_architecture="LFM", batch_correction=True, number_of_batches=3)):
        model = VariationalAutoencoder(
            feature_size=64, latent_size=4, hidden_sizes=[32],
            reconstruction_distribution="negative binomial", log_directory=str(tmp_path), seed=1,
            **kwargs)
        assert model.train(training, validation, number_of_epochs=3, minibatch_size=48,
                           learning_rate=1e-2, shuffle_seed=0) == 0
        from scvae_b200 import model_utilities as MU
        curve = MU.load_learning_curves(model, "training")["lower_bound"]
        assert len(curve) == 3 and numpy.isfinite(curve).all()
        transformed, reconstructed, latent = model.evaluate(test, minibatch_size=64,
                                                            output_versions="all")
        assert reconstructed.values.shape == (test.number_of_examples, 64)
        assert numpy.isfinite(reconstructed.values).all()
        with pytest.raises(NotImplementedError):
            model.sample(sample_size=5)


def test_train_evaluate_constrained_poisson(tmp_path):
    """`-r "constrained poisson"`: softmax over genes, the cell's count sum as parameter N."""
    from scvae_b200.variational_autoencoder import VariationalAutoencoder
    from scvae_b200 import model_utilities as MU
    full = _data(n=240, g=64, seed=9)
    training, validation, test = full.split()
    model = VariationalAutoencoder(
        feature_size=64, latent_size=4, hidden_sizes=[32],
        reconstruction_distribution="constrained poisson", log_directory=str(tmp_path), seed=1)
    assert model.use_count_sum_as_parameter
    assert model.train(training, validation, number_of_epochs=3, minibatch_size=48,
                       learning_rate=1e-2, shuffle_seed=0) == 0
    curve = MU.load_learning_curves(model, "training")["lower_bound"]
    assert len(curve) == 3 and numpy.isfinite(curve).all()
    transformed, reconstructed, latent = model.evaluate(test, minibatch_size=64, output_versions="all")
    # the reconstruction of a constrained Poisson sums to the cell's count sum
    sums = reconstructed.values.sum(axis=1)
    assert numpy.allclose(sums, numpy.asarray(test.count_sum).reshape(-1), rtol=1e-3)


def test_train_evaluate_piecewise_categorical(tmp_path):
    """`-k 2`: piecewise-categorical negative binomial through the model class."""
    from scvae_b200.variational_autoencoder import VariationalAutoencoder
    from scvae_b200 import model_utilities as MU
    full = _data(n=240, g=64, seed=10)
    training, validation, test = full.split()
    model = VariationalAutoencoder(
        feature_size=64, latent_size=4, hidden_sizes=[32], reconstruction_distribution="negative binomial",
        number_of_reconstruction_classes=2, log_directory=str(tmp_path), seed=1)
    assert "k_2" in model.name
    assert model.train(training, validation, number_of_epochs=3, minibatch_size=48,
                       learning_rate=1e-2, shuffle_seed=0) == 0
    curve = MU.load_learning_curves(model, "training")["lower_bound"]
    assert len(curve) == 3 and numpy.isfinite(curve).all()
    transformed, reconstructed, latent = model.evaluate(test, minibatch_size=64, output_versions="all")
    assert numpy.isfinite(reconstructed.values).all() and (reconstructed.values >= 0).all()


def test_gmvae_train_evaluate_with_batch_correction_and_count_sum(tmp_path):
    from scvae_b200.data_set import DataSet
    from scvae_b200.gaussian_mixture_variational_autoencoder import (
        GaussianMixtureVariationalAutoencoder)
    from scvae_b200 import model_utilities as MU
    x, labels = O.synthetic_counts(240, 48, n_types=3, seed=12, target_zero_fraction=0.8)
    x = numpy.minimum(x, 50.0)
    full = DataSet("toy", values=scipy.sparse.csr_matrix(x), labels=labels.astype(str),
                   batch_indices=(numpy.arange(240) % 2).reshape(-1, 1), batch_names=["a", "b"])
    training, validation, test = full.split()
    model = GaussianMixtureVariationalAutoencoder(
        feature_size=48, latent_size=3, hidden_sizes=[24], number_of_latent_clusters=3,
        reconstruction_distribution="negative binomial", batch_correction=True,
        number_of_batches=2, count_sum=True, log_directory=str(tmp_path), seed=2)
    assert model.train(training, validation, number_of_epochs=2, minibatch_size=48,
                       learning_rate=5e-3, shuffle_seed=0) == 0
    curve = MU.load_learning_curves(model, "training")["lower_bound"]
    assert len(curve) == 2 and numpy.isfinite(curve).all()
    transformed, reconstructed, latent = model.evaluate(test, minibatch_size=64, output_versions="all")
    assert numpy.isfinite(reconstructed.values).all()


@pytest.mark.parametrize("prior", ["uniform", "learn"])
def test_gmvae_trains_with_free_nats_under_graph_capture(tmp_path, prior):
    """Free nats for KL_y (GMVAE:3258-3261, :3391-3398) with the training step replayed from a
    CUDA graph: the threshold proportion * H[p(y)] is formed on the device (no host read inside
    the captured step) and follows a learnt prior from step to step."""
    from scvae_b200.gaussian_mixture_variational_autoencoder import (
        GaussianMixtureVariationalAutoencoder)
    from scvae_b200 import model_utilities as MU
    full = _data(n=240, g=40, seed=4)
    training, validation, _ = full.split()
    model = GaussianMixtureVariationalAutoencoder(
        feature_size=40, latent_size=3, hidden_sizes=[24], number_of_latent_clusters=4,
        reconstruction_distribution="negative binomial", prior_probabilities_method=prior,
        proportion_of_free_nats_for_y_kl_divergence=0.8, log_directory=str(tmp_path), seed=2)
    assert model.train(training, validation, number_of_epochs=3, minibatch_size=48,
                       learning_rate=5e-3, shuffle_seed=0, use_cuda_graph=True) == 0
    curves = MU.load_learning_curves(model, ["training", "validation"])
    assert len(curves["training"]["lower_bound"]) == 3
    assert numpy.isfinite(curves["training"]["lower_bound"]).all()
    assert curves["training"]["lower_bound"][-1] > curves["training"]["lower_bound"][0]


@pytest.mark.parametrize("likelihood", ["gaussian", "softplus gaussian", "log-normal", "gamma", "bernoulli",
                                        "lomax", "exponentially_modified_gaussian"])
def test_train_evaluate_with_continuous_likelihoods(tmp_path, likelihood):
    """`-r` choices outside the count family: train, evaluate, reconstruct, with the
    Bernoulli likelihood reading the binarised values as its targets (VAE:854-857)."""
    from scvae_b200.data_set import DataSet
    from scvae_b200.variational_autoencoder import VariationalAutoencoder
    from scvae_b200 import model_utilities as MU
    x, labels = O.synthetic_counts(240, 40, n_types=3, seed=4, target_zero_fraction=0.5)
    x = numpy.minimum(x, 30.0)
    if likelihood in ("log-normal", "gamma"):
        x = x + 1.0                      # strictly positive support
    binarised = scipy.sparse.csr_matrix((x > 0).astype(numpy.float32)) if likelihood == "bernoulli" else None
    full = DataSet("toy", values=scipy.sparse.csr_matrix(x), labels=labels.astype(str),
                   binarised_values=binarised)
    training, validation, test = full.split()
    model = VariationalAutoencoder(
        feature_size=40, latent_size=3, hidden_sizes=[24], reconstruction_distribution=likelihood,
        log_directory=str(tmp_path), seed=2)
    assert model.train(training, validation, number_of_epochs=3, minibatch_size=48,
                       learning_rate=2e-3, shuffle_seed=0) == 0
    curves = MU.load_learning_curves(model, ["training", "validation"])
    lb = curves["training"]["lower_bound"]
    assert len(lb) == 3 and numpy.isfinite(lb).all() and lb[-1] > lb[0]
    transformed, reconstructed, latent = model.evaluate(test, minibatch_size=32)
    assert reconstructed.values.shape == (test.number_of_examples, 40)
    if likelihood != "lomax":            # (the Lomax mean does not exist for concentration <= 1: nan)
        assert numpy.isfinite(reconstructed.values).all()


def test_training_from_host_resident_data_matches_device_resident(tmp_path):
    """train(..., data_residency="host"): the matrix stays in host memory and streams as packed
    slabs (hotloop.PackedStream); same permutation, noise and updates as the resident path."""
    from scvae_b200.variational_autoencoder import VariationalAutoencoder
    from scvae_b200 import model_utilities as MU
    full = _data(n=600, g=256, seed=6)
    training, validation, _ = full.split()
    curves = {}
    for residency in ("device", "host"):
        model = VariationalAutoencoder(
            feature_size=256, latent_size=5, hidden_sizes=[32], reconstruction_distribution="negative binomial",
            log_directory=str(tmp_path / residency), seed=3)
        assert model.train(training, None, number_of_epochs=3, minibatch_size=128, learning_rate=3e-3,
                           shuffle_seed=5, data_residency=residency) == 0
        curves[residency] = MU.load_learning_curves(model, ["training", "validation"])["training"]
    for key in ("lower_bound", "reconstruction_error", "kl_divergence"):
        a, b = numpy.asarray(curves["device"][key]), numpy.asarray(curves["host"][key])
        assert numpy.allclose(a, b, rtol=2e-4), (key, a, b)
    assert curves["host"]["lower_bound"][-1] > curves["host"]["lower_bound"][0]


@pytest.mark.parametrize("dtype", ["float32", "float16"])
def test_evaluate_streams_the_reconstruction_to_the_host(tmp_path, dtype):
    """f2: p_x_mean of every minibatch goes through hotloop.ReconstructionSink (device staging
    buffers, copy stream, pinned (N, G) result) -- several minibatches, a ragged last one, subset
    rows with deviations -- and equals the oracle's moments of the restored variables."""
    from scvae_b200.variational_autoencoder import VariationalAutoencoder
    ds = _data(n=210, g=52, seed=12)           # 52 genes: not a multiple of 8 (fp16 row pitch)
    model = VariationalAutoencoder(
        feature_size=52, latent_size=3, hidden_sizes=[16],
        reconstruction_distribution="negative binomial", log_directory=str(tmp_path), seed=2)
    model.train(ds, number_of_epochs=2, minibatch_size=70, learning_rate=1e-2)
    subset = {1, 40, 209}
    _, reconstructed, latent = model.evaluate(
        ds, minibatch_size=32, evaluation_subset_indices=subset, use_deterministic_z=True,
        reconstruction_dtype=dtype, log_results=False)
    values = reconstructed.values
    assert values.shape == (210, 52)
    params = {k: v.double() for k, v in model._get_engine().export_parameters().items()}
    cfg = O.VAEConfig(52, 3, [16], "negative binomial")
    x = torch.tensor(ds.values.toarray(), dtype=torch.float64)
    out = O.vae_forward(cfg, params, x, x, torch.zeros(1, 210, 3, dtype=torch.float64),
                        is_training=False, use_deterministic_z=True, moments=True)
    want = out["p_x_mean"].numpy()
    if dtype == "float16":
        # (fp16 carries 11 bits and saturates beyond 65504: an untrained NB head can exceed that)
        ok = want < 6e4
        assert ok.mean() > 0.99
        assert (numpy.abs(values.astype(numpy.float64) - want)[ok] <= 2e-3 * numpy.maximum(want[ok], 1.0)).all()
    else:
        assert numpy.abs(values.astype(numpy.float64) - want).max() <= 2e-4 * max(1.0, numpy.abs(want).max())
    total = reconstructed.total_standard_deviations
    for i in subset:
        got = total[i].toarray().reshape(-1)
        ref = out["p_x_stddev"][i].numpy()
        assert numpy.abs(got - ref).max() <= 2e-4 * max(1.0, numpy.abs(ref).max())
    assert total[2].toarray().max() == 0



def test_gmvae_full_covariance_mixture_train_evaluate_sample(tmp_path):
    """`-q "full-covariance gaussian mixture"` end to end (f4): train, per-epoch centroid tags with
    covariances, evaluate, sample."""
    from scvae_b200.gaussian_mixture_variational_autoencoder import (
        GaussianMixtureVariationalAutoencoder)
    from scvae_b200 import model_utilities as MU
    full = _data(n=240, g=40, seed=6)
    training, validation, test = full.split()
    model = GaussianMixtureVariationalAutoencoder(
        feature_size=40, latent_size=3, hidden_sizes=[16], number_of_latent_clusters=3,
        latent_distribution="full-covariance gaussian mixture",
        reconstruction_distribution="negative binomial", log_directory=str(tmp_path), seed=1)
    assert "full_covariance_gaussian_mixture" in model.name.replace("-", "_")
    assert model.train(training, validation, number_of_epochs=3, minibatch_size=48,
                       learning_rate=1e-2, shuffle_seed=0) == 0
    curve = MU.load_learning_curves(model, "training")["lower_bound"]
    assert len(curve) == 3 and numpy.isfinite(curve).all()
    transformed, reconstructed, latent = model.evaluate(test, minibatch_size=32)
    assert numpy.isfinite(reconstructed.values).all()
    assert latent["z"].values.shape == (test.number_of_examples, 3)
    result = model.last_evaluation
    cov = result["q_z_covariances"]
    assert cov.shape == (3, 3, 3) and numpy.allclose(cov, numpy.swapaxes(cov, 1, 2), atol=1e-6)
    assert numpy.allclose(numpy.diagonal(cov, axis1=1, axis2=2), result["q_z_variances"], rtol=1e-5)
    assert (numpy.linalg.eigvalsh(result["p_z_covariances"]) > 0).all()
    centroids = MU.load_centroids(model, "evaluation") if hasattr(MU, "load_centroids") else None
    sample_set, sample_latent = model.sample(sample_size=21, minibatch_size=8)
    assert sample_set.values.shape == (21, 40) and numpy.isfinite(sample_set.values).all()
