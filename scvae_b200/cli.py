"""Thin ``scvae train | evaluate`` front-end over the B200 model classes.

Caller side of the drop-in boundary: the two sub-commands of the reference that
enter the hot path (scvae/cli.py:111 ``train``, :267 ``evaluate``) with the reference's flag
names, short options and defaults for everything the model classes take.  Data input is limited
to what feeds the hot path here: a TSV count matrix or the built-in ``development`` data set;
analyses / plots / cross-analysis stay with the reference.
"""

import argparse
import os
import sys

from .data_set import DataSet, development_data_set, load_matrix_tsv
from .defaults import defaults
from .model_utilities import parse_model_versions


def _load(data_set_file_or_name, data_format):
    if data_set_file_or_name == "development":
        return development_data_set()
    if not os.path.exists(data_set_file_or_name):
        raise FileNotFoundError("Data set `{}` not found.".format(data_set_file_or_name))
    orientation = "fbe" if data_format == "matrix_fbe" else "ebf"
    return load_matrix_tsv(data_set_file_or_name, orientation=orientation)


def _setup_model(data_set, arguments):
    from .gaussian_mixture_variational_autoencoder import GaussianMixtureVariationalAutoencoder
    from .variational_autoencoder import VariationalAutoencoder
    common = dict(
        feature_size=data_set.number_of_features, latent_size=arguments.latent_size,
        hidden_sizes=arguments.hidden_sizes,
        number_of_monte_carlo_samples=arguments.number_of_monte_carlo_samples,
        number_of_importance_samples=arguments.number_of_importance_samples,
        reconstruction_distribution=arguments.reconstruction_distribution,
        number_of_reconstruction_classes=arguments.number_of_reconstruction_classes,
        minibatch_normalisation=arguments.minibatch_normalisation,
        inference_architecture=arguments.inference_architecture,
        generative_architecture=arguments.generative_architecture,
        batch_correction=arguments.batch_correction,
        number_of_batches=(data_set.number_of_batches if arguments.batch_correction else None),
        count_sum=arguments.count_sum,
        dropout_keep_probabilities=arguments.dropout_keep_probabilities,
        number_of_warm_up_epochs=arguments.number_of_warm_up_epochs,
        kl_weight=arguments.kl_weight, log_directory=arguments.models_directory)
    model_type = arguments.model_type.upper()
    if model_type == "VAE":
        return VariationalAutoencoder(latent_distribution=arguments.latent_distribution, **common)
    if model_type == "GMVAE":
        return GaussianMixtureVariationalAutoencoder(
            latent_distribution=arguments.latent_distribution,
            number_of_latent_clusters=arguments.number_of_classes,
            prior_probabilities_method=arguments.prior_probabilities_method,
            proportion_of_free_nats_for_y_kl_divergence=(
                arguments.proportion_of_free_nats_for_y_kl_divergence), **common)
    raise ValueError("Model type `{}` not found.".format(arguments.model_type))


def _sets(arguments):
    data_set = _load(arguments.data_set_file_or_name, arguments.format)
    if arguments.split_data_set:
        return data_set.split(arguments.splitting_method, arguments.splitting_fraction)
    return data_set, None, data_set


def train(arguments):
    training_set, validation_set, _ = _sets(arguments)
    model = _setup_model(training_set, arguments)
    print(model.description)
    return model.train(
        training_set, validation_set, number_of_epochs=arguments.number_of_epochs,
        minibatch_size=arguments.minibatch_size, learning_rate=arguments.learning_rate,
        run_id=arguments.run_id, new_run=arguments.new_run,
        reset_training=arguments.reset_training,
        temporary_log_directory=arguments.caches_directory)


def evaluate(arguments):
    training_set, validation_set, test_set = _sets(arguments)
    kind = arguments.evaluation_set_kind
    evaluation_set = {"training": training_set, "validation": validation_set,
                      "test": test_set, "full": test_set}.get(kind) or test_set
    model = _setup_model(evaluation_set, arguments)
    for version in parse_model_versions(arguments.model_versions):
        use_best = version == "best_model"
        use_early = version == "early_stopping"
        directory = model.log_directory(run_id=arguments.run_id, best_model=use_best,
                                        early_stopping=use_early)
        if not os.path.exists(os.path.join(directory, "checkpoint")):
            continue
        print("Evaluating {} version.".format(version.replace("_", " ")))
        model.evaluate(evaluation_set, minibatch_size=arguments.minibatch_size,
                       run_id=arguments.run_id, use_best_model=use_best,
                       use_early_stopping_model=use_early, output_versions="all")
        if arguments.sample_size:
            model.sample(sample_size=arguments.sample_size,
                         minibatch_size=arguments.minibatch_size, run_id=arguments.run_id,
                         use_best_model=use_best, use_early_stopping_model=use_early)
    return 0


def _parser():
    d, m = defaults["data"], defaults["models"]
    parser = argparse.ArgumentParser(prog="scvae", description="scVAE on the B200 hot path.")
    subparsers = parser.add_subparsers(dest="command")
    subparsers.required = True

    def common(sub):
        sub.add_argument("data_set_file_or_name")
        sub.add_argument("--format", "-f", default=d["format"])
        sub.add_argument("--split-data-set", action="store_true", default=d["split_data_set"])
        sub.add_argument("--splitting-method", default=d["splitting_method"])
        sub.add_argument("--splitting-fraction", type=float, default=d["splitting_fraction"])
        sub.add_argument("--model-type", "-m", default=m["type"])
        sub.add_argument("--latent-size", "-l", type=int, default=m["latent_size"])
        sub.add_argument("--hidden-sizes", "-H", type=int, nargs="+", default=m["hidden_sizes"])
        sub.add_argument("--number-of-importance-samples", type=int, nargs="+", default=None)
        sub.add_argument("--number-of-monte-carlo-samples", type=int, nargs="+", default=None)
        sub.add_argument("--latent-distribution", "-q", default=None)
        sub.add_argument("--number-of-classes", "-K", type=int, default=None)
        sub.add_argument("--reconstruction-distribution", "-r",
                         default=m["reconstruction_distribution"])
        sub.add_argument("--number-of-reconstruction-classes", "-k", type=int,
                         default=m["number_of_reconstruction_classes"])
        sub.add_argument("--prior-probabilities-method", default=m["prior_probabilities_method"])
        sub.add_argument("--number-of-warm-up-epochs", "-w", type=int,
                         default=m["number_of_warm_up_epochs"])
        sub.add_argument("--kl-weight", type=float, default=m["kl_weight"])
        sub.add_argument("--proportion-of-free-nats-for-y-kl-divergence", type=float,
                         default=m["proportion_of_free_nats_for_y_kl_divergence"])
        sub.add_argument("--minibatch-normalisation", "-b", action="store_true",
                         default=m["minibatch_normalisation"])
        sub.add_argument("--no-minibatch-normalisation", dest="minibatch_normalisation",
                         action="store_false")
        sub.add_argument("--inference-architecture", default=m["inference_architecture"])
        sub.add_argument("--generative-architecture", default=m["generative_architecture"])
        sub.add_argument("--batch-correction", "--bc", action="store_true",
                         default=m["batch_correction"])
        sub.add_argument("--count-sum", action="store_true", default=m["count_sum"])
        sub.add_argument("--dropout-keep-probabilities", type=float, nargs="+",
                         default=m["dropout_keep_probabilities"])
        sub.add_argument("--minibatch-size", "-B", type=int, default=m["minibatch_size"])
        sub.add_argument("--run-id", default=m["run_id"])
        sub.add_argument("--models-directory", "-M", default=m["directory"])

    train_parser = subparsers.add_parser("train")
    common(train_parser)
    train_parser.add_argument("--number-of-epochs", "-e", type=int, default=m["number_of_epochs"])
    train_parser.add_argument("--learning-rate", type=float, default=m["learning_rate"])
    train_parser.add_argument("--new-run", action="store_true", default=m["new_run"])
    train_parser.add_argument("--reset-training", action="store_true", default=m["reset_training"])
    train_parser.add_argument("--caches-directory", "-C", default=None)
    train_parser.set_defaults(func=train)

    evaluate_parser = subparsers.add_parser("evaluate")
    common(evaluate_parser)
    evaluate_parser.add_argument("--evaluation-set-kind", default=defaults["evaluation"]["data_set_kind"])
    evaluate_parser.add_argument("--sample-size", type=int, default=m["sample_size"])
    evaluate_parser.add_argument("--model-versions", nargs="+",
                                 default=defaults["evaluation"]["model_versions"])
    evaluate_parser.set_defaults(func=evaluate)
    return parser


def main(argv=None):
    arguments = _parser().parse_args(argv)
    return arguments.func(arguments)


if __name__ == "__main__":
    sys.exit(main())
