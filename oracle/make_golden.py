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
  split_indices.npz          ``split_data_set`` default/random split indices for N=1000
                             (processing.py:336-493, RandomState(42)).
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


def main():
    os.makedirs(OUT, exist_ok=True)
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


if __name__ == "__main__":
    main()
