"""Host-side logic that needs no GPU: CLI surface, option validation, model naming."""
import pytest


def test_cli_accepts_the_reference_flags_of_the_hot_path():
    from scvae_b200 import cli
    a = cli._parser().parse_args(
        ["train", "development", "-m", "GMVAE", "-K", "4", "-r", "negative binomial", "-k", "2",
         "--bc", "--count-sum", "--generative-architecture", "LFM", "-l", "8", "-H", "64", "32",
         "-e", "3", "-B", "128"])
    assert a.model_type == "GMVAE" and a.number_of_classes == 4
    assert a.number_of_reconstruction_classes == 2 and a.batch_correction and a.count_sum
    assert a.generative_architecture == "LFM" and a.hidden_sizes == [64, 32]
    e = cli._parser().parse_args(["evaluate", "development", "--model-versions", "end_of_training", "best_model"])
    assert e.func is cli.evaluate


def test_option_validation_mirrors_the_scope():
    from scvae_b200.variational_autoencoder import VariationalAutoencoder
    from scvae_b200.gaussian_mixture_variational_autoencoder import (
        GaussianMixtureVariationalAutoencoder)
    # built this round: -k, batch correction, count sum, LFM, constrained Poisson (VAE)
    model = VariationalAutoencoder(feature_size=50, latent_size=4, hidden_sizes=[16],
                                   reconstruction_distribution="negative binomial",
                                   number_of_reconstruction_classes=3, batch_correction=True,
                                   number_of_batches=2, count_sum=True,
                                   generative_architecture="LFM")
    assert "k_3" in model.name and "bc" in model.name and "ga_LFM" in model.name
    VariationalAutoencoder(feature_size=50, latent_size=4, hidden_sizes=[16],
                           reconstruction_distribution="constrained poisson")
    GaussianMixtureVariationalAutoencoder(feature_size=50, latent_size=4, hidden_sizes=[16],
                                          number_of_latent_clusters=3,
                                          reconstruction_distribution="negative binomial",
                                          number_of_reconstruction_classes=2, batch_correction=True,
                                          number_of_batches=2)
    with pytest.raises(ValueError):      # as the reference (MU:883-897): no -k around zero inflation
        VariationalAutoencoder(feature_size=50, latent_size=4, hidden_sizes=[16],
                               reconstruction_distribution="zero-inflated poisson",
                               number_of_reconstruction_classes=2)
    # still outside: loud, never silently degraded
    VariationalAutoencoder(feature_size=50, latent_size=4, hidden_sizes=[16],
                           dropout_keep_probabilities=[0.9])          # VAE: built
    with pytest.raises(NotImplementedError):
        GaussianMixtureVariationalAutoencoder(feature_size=50, latent_size=4, hidden_sizes=[16],
                                              number_of_latent_clusters=3,
                                              dropout_keep_probabilities=[0.9])
    GaussianMixtureVariationalAutoencoder(feature_size=50, latent_size=4, hidden_sizes=[16],
                                          number_of_latent_clusters=3,
                                          reconstruction_distribution="constrained poisson")
    # the continuous / binary reconstruction distributions are built (csrc/continuous.cu)
    for name in ("gamma", "gaussian", "modified gaussian", "log-normal", "bernoulli", "lomax",
                 "exponentially_modified_gaussian"):
        VariationalAutoencoder(feature_size=50, latent_size=4, hidden_sizes=[16],
                               reconstruction_distribution=name)
    with pytest.raises(NotImplementedError):
        VariationalAutoencoder(feature_size=50, latent_size=4, hidden_sizes=[16],
                               reconstruction_distribution="multinomial")
    with pytest.raises(TypeError):
        VariationalAutoencoder(feature_size=50, latent_size=4, hidden_sizes=[16],
                               batch_correction=True)          # number of batches missing
