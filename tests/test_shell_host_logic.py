"""The reference-facing model classes end to end on the CPU: ``train`` (epochs, per-epoch
evaluation passes, warm-up, checkpoints, summaries, resume), ``evaluate`` and ``sample`` of
``VariationalAutoencoder`` / ``GaussianMixtureVariationalAutoencoder`` with every kernel wrapper
replaced by its CPU stand-in (``tests/kernel_standins.py``).  The test bodies are the GPU ones of
``tests/test_gpu_model.py``, run unchanged: what they assert about the on-disk contract
(checkpoint naming, event tags, learning curves, resume semantics) and the returned data sets is
host logic and holds without a device.  The product has no CPU path; the stand-ins are injected
by monkeypatching inside this module only.
"""
import pytest
import torch

import kernel_standins
import test_gpu_model as G


@pytest.fixture
def shell_on_cpu(monkeypatch):
    import scvae_b200
    import scvae_b200.engine as E
    import scvae_b200.gmvae_engine as GE
    import scvae_b200.hotloop as H
    import scvae_b200.variational_autoencoder as V
    monkeypatch.setattr(scvae_b200, "kernels", kernel_standins)      # function-local imports
    for module in (E, GE, H):
        monkeypatch.setattr(module, "K", kernel_standins)

    def on_cpu(cls, **forced):
        original = cls.__init__

        def init(self, *args, **kwargs):
            kwargs.update(forced)
            original(self, *args, **kwargs)
            if hasattr(self, "overlap_streams"):
                self.overlap_streams = False
        monkeypatch.setattr(cls, "__init__", init)

    on_cpu(E.VAEEngine, device="cpu", tensor_cores=False)
    on_cpu(GE.GMVAEEngine, device="cpu", tensor_cores=False)
    on_cpu(V.VariationalAutoencoder, device="cpu")
    on_cpu(H.TrainLoop, use_graph=False)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)

    def on_host(factory):
        def build(*args, **kwargs):
            if str(kwargs.get("device", "")).startswith("cuda"):
                kwargs["device"] = "cpu"
            return factory(*args, **kwargs)
        return build

    monkeypatch.setattr(torch, "zeros", on_host(torch.zeros))       # scratch tensors of the tests
    monkeypatch.setattr(torch, "tensor", on_host(torch.tensor))
    del kernel_standins.launches[:]
    return kernel_standins.launches


def test_vae_train_resume_evaluate_sample_on_cpu(shell_on_cpu, tmp_path):
    G.test_train_resume_evaluate_sample(tmp_path)
    assert "adam_clip_step" in shell_on_cpu and "csr_densify" in shell_on_cpu


def test_gmvae_train_evaluate_sample_on_cpu(shell_on_cpu, tmp_path):
    G.test_gmvae_train_evaluate_sample(tmp_path)
    assert "gmvae_bound" in shell_on_cpu


def test_vae_options_on_cpu(shell_on_cpu, tmp_path):
    G.test_train_evaluate_with_batch_correction_count_sum_and_lfm(tmp_path)


def test_vae_constrained_poisson_on_cpu(shell_on_cpu, tmp_path):
    G.test_train_evaluate_constrained_poisson(tmp_path)


def test_vae_piecewise_categorical_on_cpu(shell_on_cpu, tmp_path):
    G.test_train_evaluate_piecewise_categorical(tmp_path)


def test_gmvae_options_on_cpu(shell_on_cpu, tmp_path):
    G.test_gmvae_train_evaluate_with_batch_correction_and_count_sum(tmp_path)


def test_unit_variance_gaussian_with_sampled_kl_on_cpu(shell_on_cpu, tmp_path):
    import test_zz_gpu_reference_graph as Z
    Z.test_train_evaluate_unit_variance_gaussian_with_its_default_sampled_kl(tmp_path)
    assert "gaussian_sampled_kl_bwd" in shell_on_cpu and "vae_bound" not in shell_on_cpu


def test_dropout_through_the_model_class_on_cpu(shell_on_cpu, tmp_path):
    import test_zz_gpu_reference_graph as Z
    Z.test_train_evaluate_with_dropout(tmp_path)
    assert "dropout_fwd" in shell_on_cpu and "dropout_bwd" in shell_on_cpu


def test_gmvae_dropout_through_the_model_class_on_cpu(shell_on_cpu, tmp_path):
    import test_zz_gpu_reference_graph as Z
    Z.test_gmvae_train_evaluate_with_dropout(tmp_path)
    assert "dropout_fwd" in shell_on_cpu and "dropout_bwd" in shell_on_cpu


def test_gmvae_constrained_poisson_through_the_model_class_on_cpu(shell_on_cpu, tmp_path):
    import test_zz_gpu_reference_graph as Z
    Z.test_gmvae_train_evaluate_constrained_poisson(tmp_path)
    assert "constrained_poisson_mixture_moments" in shell_on_cpu


def test_cli_train_and_evaluate_on_tsv_on_cpu(shell_on_cpu, tmp_path):
    """BASELINE configs[0] in miniature (TSV -> `scvae train` -> `scvae evaluate`) through the
    command-line front end."""
    G.test_cli_train_and_evaluate_on_tsv(tmp_path)
    assert "adam_clip_step" in shell_on_cpu


def test_one_epoch_matches_oracle_on_cpu(shell_on_cpu, tmp_path):
    """train() for one epoch == the oracle's steps on the same permutation and noise: minibatch
    order, learning rate, warm-up weight and step counting of the training loop."""
    G.test_one_epoch_matches_oracle(tmp_path)
    assert shell_on_cpu.count("adam_clip_step") == 2


@pytest.mark.parametrize("prior", ["uniform", "learn"])
def test_gmvae_free_nats_on_cpu(shell_on_cpu, tmp_path, prior):
    G.test_gmvae_trains_with_free_nats_under_graph_capture(tmp_path, prior)
    assert "gmvae_bound" in shell_on_cpu


@pytest.mark.parametrize("likelihood", ["gaussian", "log-normal", "gamma", "bernoulli", "lomax",
                                        "exponentially_modified_gaussian"])
def test_continuous_likelihoods_through_the_model_class_on_cpu(shell_on_cpu, tmp_path, likelihood):
    G.test_train_evaluate_with_continuous_likelihoods(tmp_path, likelihood)
    assert "continuous_likelihood" in shell_on_cpu and "continuous_moments" in shell_on_cpu



def test_gmvae_full_covariance_mixture_on_cpu(shell_on_cpu, tmp_path):
    G.test_gmvae_full_covariance_mixture_train_evaluate_sample(tmp_path)
    assert "gmvae_latent_full_fwd" in shell_on_cpu and "gmvae_latent_full_bwd" in shell_on_cpu
    assert "gmvae_full_covariance_mean" in shell_on_cpu
