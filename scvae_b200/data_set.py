"""Minimal ``DataSet`` for the caller side of the drop-in boundary.

Only what the model shells read is implemented (SURVEY §1, L3 -> L1): CSR ``values``,
``preprocessed_values`` / ``binarised_values`` (optional), ``count_sum``,
``normalised_count_sum``, ``number_of_examples``, ``number_of_features``, labels, names,
``kind``, ``noisy_preprocess``, ``batch_indices``; plus the three input routes needed to feed
the hot path: in-memory arrays, a TSV matrix (the reference's ``matrix_ebf`` / ``matrix_fbe``
loaders, scvae/data/loaders.py:399-404, :801-871) and the deterministic ``development``
data set (loaders.py:942-1022).  Formats, feature mapping, HDF5 caches, preprocessing
pipelines and plots stay with the reference (out of scope: SURVEY §2 rows 7-9).
"""

import numpy
import scipy.sparse


class DataSet:
    def __init__(self, name, values=None, labels=None, example_names=None, feature_names=None,
                 title=None, kind="full", version="original", preprocessed_values=None,
                 binarised_values=None, batch_indices=None, batch_names=None,
                 total_standard_deviations=None, explained_standard_deviations=None, **extra):
        self.name = name
        self.title = title or name
        self.kind = kind
        self.version = version
        self.values = None
        self.count_sum = None
        self.normalised_count_sum = None
        self.preprocessed_values = preprocessed_values
        self.binarised_values = binarised_values
        self.labels = None if labels is None else numpy.asarray(labels)
        self.example_names = example_names
        self.feature_names = feature_names
        self.batch_indices = batch_indices
        self.batch_names = batch_names
        self.total_standard_deviations = total_standard_deviations
        self.explained_standard_deviations = explained_standard_deviations
        self.noisy_preprocess = None
        self.preprocessing_methods = extra.get("preprocessing_methods", [])
        self.specifications = extra.get("specifications", {})
        self.features_mapped = extra.get("features_mapped", False)
        self.feature_selection = extra.get("feature_selection", [])
        self.example_filter = extra.get("example_filter", [])
        self.label_superset = None
        self.predicted_cluster_ids = None
        if values is not None:
            self.update(values=values)

    @property
    def number_of_batches(self):
        """Number of distinct batches (DataSet.number_of_batches of the reference, DS:309-316)."""
        if self.batch_names is not None:
            return len(self.batch_names)
        if self.batch_indices is None:
            return None
        return int(numpy.asarray(self.batch_indices).max()) + 1

    def update(self, values=None, **_):
        if values is not None:
            if not scipy.sparse.issparse(values):
                values = numpy.asarray(values, dtype=numpy.float32)
                if values.ndim != 2:
                    raise ValueError("`values` must be a 2-D (examples x features) matrix.")
            self.values = values
            count_sum = numpy.asarray(values.sum(axis=1)).reshape(-1, 1)
            self.count_sum = count_sum
            self.normalised_count_sum = count_sum / max(count_sum.max(), 1e-30)
            m, n = values.shape
            if self.example_names is None:
                self.example_names = numpy.array(["example {}".format(i + 1) for i in range(m)])
            if self.feature_names is None:
                self.feature_names = numpy.array(["feature {}".format(j + 1) for j in range(n)])
            if self.labels is not None and len(self.labels) != m:
                raise ValueError("The number of labels does not match the number of examples.")

    # --- attributes read by the models --------------------------------------------------------
    @property
    def number_of_examples(self):
        return self.values.shape[0]

    @property
    def number_of_features(self):
        return self.values.shape[1]

    @property
    def has_values(self):
        return self.values is not None

    @property
    def has_preprocessed_values(self):
        return self.preprocessed_values is not None

    @property
    def has_binarised_values(self):
        return self.binarised_values is not None

    @property
    def has_labels(self):
        return self.labels is not None

    @property
    def class_names(self):
        return None if self.labels is None else numpy.unique(self.labels).tolist()

    @property
    def number_of_classes(self):
        return None if self.labels is None else len(self.class_names)

    def csr(self, which="values"):
        """The requested matrix as float32 CSR (what the hot loop uploads)."""
        matrix = getattr(self, which)
        if matrix is None:
            raise ValueError("Data set has no `{}`.".format(which))
        return scipy.sparse.csr_matrix(matrix, dtype=numpy.float32)

    def subset(self, indices, kind):
        def take(a):
            return None if a is None else a[indices]
        return DataSet(self.name, values=self.values[indices], labels=take(self.labels),
                       example_names=take(self.example_names), feature_names=self.feature_names,
                       title=self.title, kind=kind, version=self.version,
                       preprocessed_values=take(self.preprocessed_values),
                       binarised_values=take(self.binarised_values),
                       batch_indices=take(self.batch_indices), batch_names=self.batch_names)

    def split(self, method=None, fraction=None):
        """training / validation / test sets with the reference's rule (processing.py:336-493):
        RandomState(42) permutation, ``fraction`` of all examples for training+validation and
        ``fraction`` of those for training."""
        method = (method or "default").lower()
        fraction = 0.9 if fraction is None else fraction
        if method == "default":
            method = "random"
        n = self.number_of_examples
        if method == "random":
            indices = numpy.random.RandomState(42).permutation(n)
        elif method == "sequential":
            indices = numpy.arange(n)
        else:
            raise ValueError("Splitting method `{}` not found.".format(method))
        n_training_validation = int(fraction * n)
        n_training = int(fraction * n_training_validation)
        return (self.subset(indices[:n_training], "training"),
                self.subset(indices[n_training:n_training_validation], "validation"),
                self.subset(indices[n_training_validation:], "test"))


# ---------------------------------------------------------------------------------------------
# input routes
# ---------------------------------------------------------------------------------------------
def load_matrix_tsv(path, name=None, orientation="ebf"):
    """Tab-separated count matrix with a header row and a name column; ``ebf`` =
    examples-by-features, ``fbe`` = features-by-examples (loaders.py:801-871)."""
    import pandas
    frame = pandas.read_csv(path, sep="\t", index_col=0, header=0)
    values = frame.to_numpy(dtype=numpy.float32)
    row_names = numpy.array(frame.index.astype(str))
    column_names = numpy.array(frame.columns.astype(str))
    if orientation == "fbe":
        values, row_names, column_names = values.T, column_names, row_names
    elif orientation != "ebf":
        raise ValueError("`orientation` must be `ebf` or `fbe`.")
    import os
    name = name or os.path.splitext(os.path.basename(path))[0]
    return DataSet(name, values=scipy.sparse.csr_matrix(values), example_names=row_names,
                   feature_names=column_names)


def development_data_set(n_examples=10000, n_features=25, scale=10, update_probability=0.0001):
    """The reference's built-in deterministic test data: zero-inflated NB counts per cell type,
    ``numpy.random.RandomState(60)`` (loaders.py:942-1022).  The sequence of draws from the
    random stream is part of the specification (pinned by tests/golden/development_data_set.npz):
    type parameters (r, p, dropout) -> one uniform per example with occasional re-draws ->
    shuffle -> 10 % unlabelled -> per entry one NB draw then one Bernoulli draw."""
    rs = numpy.random.RandomState(60)

    def draw_type():
        return scale * rs.rand(n_features), rs.rand(n_features), rs.rand(n_features)

    r = numpy.empty((n_examples, n_features))
    p = numpy.empty((n_examples, n_features))
    keep = numpy.empty((n_examples, n_features))
    labels = numpy.empty(n_examples, numpy.int32)
    current, label = draw_type(), 1
    for i in range(n_examples):
        if rs.rand() > 1 - update_probability:
            current, label = draw_type(), label + 1
        r[i], p[i], keep[i] = current
        labels[i] = label
    order = rs.permutation(n_examples)
    r, p, keep, labels = r[order], p[order], keep[order], labels[order]
    labels[rs.permutation(n_examples)[:int(0.1 * n_examples)]] = 0
    values = numpy.empty((n_examples, n_features), numpy.float32)
    for i in range(n_examples):
        for j in range(n_features):
            count = rs.negative_binomial(r[i, j], p[i, j])
            values[i, j] = rs.binomial(1, keep[i, j]) * count
    return DataSet(
        "development", values=scipy.sparse.csr_matrix(values), labels=labels.astype(str),
        example_names=numpy.array(["example {}".format(i + 1) for i in range(n_examples)]),
        feature_names=numpy.array(["feature {}".format(j + 1) for j in range(n_features)]))
