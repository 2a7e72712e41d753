"""Generate golden fixtures by importing the REFERENCE's own code (test infrastructure).

Runs only in the build container (needs /root/reference); its outputs are committed under
``tests/golden/`` so that nothing at test/bench time reads the reference tree.

The reference cannot be imported as a package here (TensorFlow 1.15, PyTables, loompy are not
installable: SURVEY §8c), but ``scvae/data/loaders.py`` and ``scvae/data/processing.py`` only
*use* those modules inside format-specific loaders, so they are loaded by file path with empty
stand-in modules for the missing imports.  Fixtures:

  development_data_set.npz   the reference's deterministic synthetic data set
                             (``_create_development_data_set``, loaders.py:942-1022,
                             RandomState(60)): values (10000 x 25), labels.
  normalise_string.npz       ``scvae.utilities.normalise_string`` on distribution / data names.
  model_utilities.json       the pure-Python helpers of ``scvae/models/utilities.py`` that the CLI
                             and the model classes call (MU:140-176, :591-613, :647-658,
                             :755-898, :973-990) and ``format_duration`` of ``scvae/utilities.py``:
                             inputs -> outputs (or the exception type raised).
  model_names.json           ``name`` / ``description`` / ``log_directory()`` of the reference's
                             ``VariationalAutoencoder`` (VAE:412-608) and
                             ``GaussianMixtureVariationalAutoencoder`` (GMVAE:441-640) classes for
                             a sweep of constructor options: the on-disk naming contract.  The
                             classes are imported with mock TensorFlow / TFP packages and their
                             three graph-building methods replaced by no-ops (run as
                             ``make_golden.py names`` in a fresh interpreter).
"""

import hashlib
import importlib.util
import os
import sys
import types

import numpy

REFERENCE = os.environ.get("SCVAE_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    module = importlib.util.module_from_spec(spec)
    sys.modules[name] = module
    spec.loader.exec_module(module)
    return module


def import_reference_modules():
    for missing in ("loompy", "tables", "tensorflow", "importlib_resources"):
        if missing not in sys.modules:
            sys.modules[missing] = types.ModuleType(missing)
    pkg = types.ModuleType("scvae")
    pkg.__path__ = [os.path.join(REFERENCE, "scvae")]
    sys.modules["scvae"] = pkg
    utilities = _load_by_path("scvae.utilities", os.path.join(REFERENCE, "scvae", "utilities.py"))
    data_pkg = types.ModuleType("scvae.data")
    data_pkg.__path__ = [os.path.join(REFERENCE, "scvae", "data")]
    sys.modules["scvae.data"] = data_pkg
    loaders = _load_by_path("scvae.data.loaders",
                            os.path.join(REFERENCE, "scvae", "data", "loaders.py"))
    return utilities, loaders


def import_reference_model_utilities():
    """scvae/models/utilities.py with stand-ins for TensorFlow (only its graph helpers use it)."""
    tf = types.ModuleType("tensorflow")
    tf.reduce_mean = None
    contrib = types.ModuleType("tensorflow.contrib")
    layers = types.ModuleType("tensorflow.contrib.layers")
    layers.fully_connected = layers.batch_norm = layers.dropout = None
    sys.modules.update({"tensorflow": tf, "tensorflow.contrib": contrib,
                        "tensorflow.contrib.layers": layers})
    models_pkg = types.ModuleType("scvae.models")
    models_pkg.__path__ = [os.path.join(REFERENCE, "scvae", "models")]
    sys.modules["scvae.models"] = models_pkg
    return _load_by_path("scvae.models.utilities",
                         os.path.join(REFERENCE, "scvae", "models", "utilities.py"))


def _record(function, *args, **kwargs):
    try:
        value = function(*args, **kwargs)
        if isinstance(value, tuple):
            value = list(value)
        return {"args": list(args), "kwargs": kwargs, "result": value}
    except Exception as exc:            # the type is part of the contract (SURVEY 8b "Errors")
        return {"args": list(args), "kwargs": kwargs, "raises": type(exc).__name__}


def model_utilities_golden(utilities):
    import json
    MU = import_reference_model_utilities()
    nan = float("nan")
    records = {
        "parse_numbers_of_samples": [_record(MU.parse_numbers_of_samples, a) for a in
                                     (1, [5], [5, 2], {"training": 3, "evaluation": 7}, [1, 2, 3])],
        "parse_model_versions": [_record(MU.parse_model_versions, a) for a in
                                 ("all", ["all"], "best_model", ["end_of_training", "early_stopping"],
                                  ["e", "b"], "nonsense")],
        "early_stopping_status": [_record(MU.early_stopping_status, a, b) for a, b in
                                  (([-10, -9, -9.5, -9.6, -9.7], 2), ([-10, -9, -8], 10),
                                   ([-5, -6, -7, -8], 3), ([-5.0], 10), ([-3, -2, -2.5], 1))],
        "check_run_id": [_record(MU.check_run_id, a) for a in
                         ("Run_1", "2019-run", "bad id!", "a/b", "")],
        "validate_model_parameters": [
            _record(MU.validate_model_parameters, **kw) for kw in (
                dict(reconstruction_distribution="negative binomial", number_of_reconstruction_classes=3,
                     model_type="VAE", latent_distribution="gaussian", parameterise_latent_posterior=False),
                dict(reconstruction_distribution="zero-inflated poisson", number_of_reconstruction_classes=2,
                     model_type="VAE", latent_distribution="gaussian", parameterise_latent_posterior=False),
                dict(reconstruction_distribution="constrained poisson", number_of_reconstruction_classes=2,
                     model_type="VAE", latent_distribution="gaussian", parameterise_latent_posterior=False),
                dict(reconstruction_distribution="bernoulli", number_of_reconstruction_classes=1,
                     model_type="VAE", latent_distribution="gaussian", parameterise_latent_posterior=False),
                dict(reconstruction_distribution="poisson", number_of_reconstruction_classes=0,
                     model_type="GMVAE", latent_distribution="gaussian mixture",
                     parameterise_latent_posterior=False),
                dict(reconstruction_distribution="poisson", number_of_reconstruction_classes=0,
                     model_type="VAE", latent_distribution="gaussian mixture",
                     parameterise_latent_posterior=True),
            )],
        "build_training_string": [_record(MU.build_training_string, *a) for a in
                                  (("model", 0, 10, "training set"), ("model for run r", 3, 10, "training set"),
                                   ("model", 10, 10, "training set"), ("model", 12, 10, "full set"))],
        "format_duration": [_record(utilities.format_duration, a) for a in
                            (0.0004, 0.5, 3.2, 75, 3700, 90000)],
    }

    def clean(o):
        if isinstance(o, float) and o != o:
            return "nan"
        if isinstance(o, dict):
            return {k: clean(v) for k, v in o.items()}
        if isinstance(o, (list, tuple)):
            return [clean(v) for v in o]
        if isinstance(o, (numpy.bool_, bool)):
            return bool(o)
        if isinstance(o, numpy.generic):
            return o.item()
        return o
    with open(os.path.join(OUT, "model_utilities.json"), "w") as handle:
        json.dump(clean(records), handle, indent=1, sort_keys=True)
    print("model utilities:", {k: len(v) for k, v in records.items()})


class _MockModule(types.ModuleType):
    """Package whose attributes are mocks; ``Distribution`` is a real class (it is subclassed)."""
    __path__ = []

    def __getattr__(self, key):
        from unittest import mock
        if key.startswith("__"):
            raise AttributeError(key)
        value = (type(key, (object,), {"__init__": lambda self, *a, **kw: None})
                 if key == "Distribution" else mock.MagicMock(name=self.__name__ + "." + key))
        setattr(self, key, value)
        return value


class _MockFinder:
    """Serves any (sub)module of the packages the reference imports but this image lacks."""
    ROOTS = ("tensorflow", "tensorflow_probability", "loompy", "tables", "matplotlib", "seaborn",
             "mpl_toolkits")

    def find_spec(self, name, path=None, target=None):
        import importlib.machinery
        if name.split(".")[0] in self.ROOTS or name.startswith("scvae.analyses"):
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        return _MockModule(spec.name)

    def exec_module(self, module):
        pass


MODEL_NAME_CASES = [
    ("VAE", dict()),
    ("VAE", dict(latent_size=50, hidden_sizes=[100], reconstruction_distribution="negative binomial")),
    ("VAE", dict(latent_size=8, hidden_sizes=[64, 32], reconstruction_distribution="poisson",
                 number_of_reconstruction_classes=3, batch_correction=True, number_of_batches=2,
                 count_sum=True)),
    ("VAE", dict(reconstruction_distribution="zero-inflated negative binomial",
                 minibatch_normalisation=False, number_of_warm_up_epochs=20)),
    ("VAE", dict(reconstruction_distribution="constrained poisson", kl_weight=0.5)),
    ("VAE", dict(reconstruction_distribution="zero-inflated poisson",
                 number_of_monte_carlo_samples={"training": 3, "evaluation": 7},
                 number_of_importance_samples={"training": 2, "evaluation": 5})),
    ("VAE", dict(reconstruction_distribution="negative binomial", dropout_keep_probabilities=[0.8, 0.9])),
    ("VAE", dict(reconstruction_distribution="negative binomial", inference_architecture="LFM",
                 latent_size=2, hidden_sizes=[])),
    ("VAE", dict(reconstruction_distribution="negative binomial", inference_architecture="MLP",
                 generative_architecture="LFM")),
    ("VAE", dict(latent_distribution="gaussian mixture", number_of_latent_clusters=4,
                 analytical_kl_term=False)),
    ("VAE", dict(latent_distribution="gaussian", analytical_kl_term=False, count_sum_feature=True)),
    ("VAE", dict(reconstruction_distribution="poisson", count_sum_feature=True, count_sum=True,
                 batch_correction=True, number_of_batches=3)),
    ("GMVAE", dict()),
    ("GMVAE", dict(latent_size=50, hidden_sizes=[100, 100], number_of_latent_clusters=9,
                   reconstruction_distribution="zero-inflated negative binomial")),
    ("GMVAE", dict(latent_distribution="full-covariance gaussian mixture", number_of_latent_clusters=3)),
    ("GMVAE", dict(number_of_latent_clusters=5, prior_probabilities_method="uniform",
                   reconstruction_distribution="negative binomial", number_of_reconstruction_classes=4,
                   number_of_warm_up_epochs=10, kl_weight=2)),
    ("GMVAE", dict(number_of_latent_clusters=5, proportion_of_free_nats_for_y_kl_divergence=0.8,
                   minibatch_normalisation=False, batch_correction=True, number_of_batches=2)),
    # the reference's GMVAE has no LFM form: it takes these keywords and ignores them
    ("GMVAE", dict(number_of_latent_clusters=3, inference_architecture="LFM",
                   generative_architecture="LFM")),
    ("GMVAE", dict(number_of_latent_clusters=2, number_of_monte_carlo_samples=[4, 2],
                   number_of_importance_samples=3, dropout_keep_probabilities=[0.5],
                   count_sum_feature=True)),
]

LOG_DIRECTORY_CASES = [
    dict(), dict(run_id="r1"), dict(run_id="r1", early_stopping=True),
    dict(run_id="r1", best_model=True), dict(base="other", early_stopping=True),
]


def model_names_golden():
    """Names, descriptions and log directories from the reference's own model classes."""
    import importlib
    import json
    sys.meta_path.insert(0, _MockFinder())
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
    classes = {
        "VAE": importlib.import_module(
            "scvae.models.variational_autoencoder").VariationalAutoencoder,
        "GMVAE": importlib.import_module(
            "scvae.models.gaussian_mixture_variational_autoencoder"
        ).GaussianMixtureVariationalAutoencoder,
    }
    for cls in classes.values():
        for method in ("_setup_model_graph", "_setup_loss_function", "_setup_optimiser"):
            setattr(cls, method, lambda self: None)
    records = []
    for kind, kwargs in MODEL_NAME_CASES:
        record = {"model": kind, "kwargs": kwargs}
        try:
            model = classes[kind](feature_size=100, log_directory="log", **kwargs)
            record["name"] = model.name
            record["description"] = model.description
            record["log_directories"] = [
                {"kwargs": kw, "result": model.log_directory(**kw)} for kw in LOG_DIRECTORY_CASES]
        except Exception as exc:
            record["raises"] = type(exc).__name__
        records.append(record)
    with open(os.path.join(OUT, "model_names.json"), "w") as handle:
        json.dump(records, handle, indent=1, sort_keys=True)
    print("model names:", len(records), "records,",
          sum("raises" in r for r in records), "raising")


def main():
    os.makedirs(OUT, exist_ok=True)
    if sys.argv[1:] == ["names"]:
        return model_names_golden()
    utilities, loaders = import_reference_modules()
    dd = loaders._create_development_data_set()
    values = dd["values"]
    assert values.shape == (10000, 25) and values.dtype == numpy.float32
    assert values.max() < 65536 and numpy.array_equal(values, numpy.round(values))
    digest = hashlib.sha256(values.tobytes()).hexdigest()
    numpy.savez_compressed(
        os.path.join(OUT, "development_data_set.npz"),
        values=values.astype(numpy.uint16), labels=dd["labels"].astype("U8"),
        sha256=numpy.array(digest))
    print("development data set:", values.shape, "sum", values.sum(), "sha256", digest[:16])

    # normalise_string golden (utilities.py) -- used for distribution / model-name parsing
    samples = ["Zero-Inflated Negative Binomial", "negative binomial", "gaussian mixture",
               "10x-PBMC PP", "unit-variance gaussian", "constrained poisson"]
    numpy.savez(os.path.join(OUT, "normalise_string.npz"),
                inputs=numpy.array(samples), outputs=numpy.array(
                    [utilities.normalise_string(s) for s in samples]))
    print("normalise_string:", [utilities.normalise_string(s) for s in samples])
    model_utilities_golden(utilities)
    import subprocess
    subprocess.run([sys.executable, os.path.abspath(__file__), "names"], check=True)


if __name__ == "__main__":
    main()
