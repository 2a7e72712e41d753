"""``GaussianMixtureVariationalAutoencoder``: the reference's GMVAE class
(scvae/models/gaussian_mixture_variational_autoencoder.py:51) on the B200 step engine.

Shares the train / evaluate / sample shell with ``VariationalAutoencoder`` and specialises what
the reference specialises: constructor keywords (``prior_probabilities_method``,
``prior_probabilities``, ``number_of_latent_clusters``,
``proportion_of_free_nats_for_y_kl_divergence``), the model name, the fetched quantities
(KL_z / KL_y, q(y|x) logits, cluster statistics), the cluster -> label accuracy logged per
epoch (GMVAE:1299-1333) and the ``{"z", "y"}`` latent outputs of ``evaluate``.
"""

import os

import numpy

from .data_set import DataSet
from .defaults import defaults
from .model_utilities import normalise_string
from .variational_autoencoder import VariationalAutoencoder

GMVAE_LATENT_DISTRIBUTIONS = ["gaussian mixture", "full-covariance gaussian mixture",
                              "legacy gaussian mixture"]


def map_cluster_ids_to_label_ids(label_ids, cluster_ids, excluded_class_ids=()):
    """Each cluster predicts the most frequent label among its members
    (scvae/analyses/prediction.py:134-146)."""
    predicted = numpy.zeros_like(cluster_ids)
    for cluster in numpy.unique(cluster_ids):
        members = cluster_ids == cluster
        labels = label_ids[members]
        for excluded in excluded_class_ids:
            labels = labels[labels != excluded]
        if labels.size:
            values, counts = numpy.unique(labels, return_counts=True)
            predicted[members] = values[counts.argmax()]     # smallest label wins ties (mode)
    return predicted


def accuracy(label_ids, predicted_label_ids, excluded_class_ids=()):
    """Fraction of correctly mapped labels, ignoring excluded classes
    (scvae/analyses/metrics/clustering.py:145-148)."""
    keep = numpy.ones(len(label_ids), dtype=bool)
    for excluded in excluded_class_ids:
        keep &= label_ids != excluded
    return float(numpy.mean(predicted_label_ids[keep] == label_ids[keep])) if keep.any() else None


class GaussianMixtureVariationalAutoencoder(VariationalAutoencoder):
    _model_type = "GMVAE"
    _loss_keys = ("lower_bound", "reconstruction_error", "kl_divergence_z", "kl_divergence_y")
    _latent_key = "z_mean"

    def __init__(self, feature_size, latent_size=None, hidden_sizes=None,
                 reconstruction_distribution=None, number_of_reconstruction_classes=None,
                 prior_probabilities_method=None, prior_probabilities=None,
                 number_of_latent_clusters=None, minibatch_normalisation=None,
                 batch_correction=None, number_of_batches=None,
                 proportion_of_free_nats_for_y_kl_divergence=None, number_of_warm_up_epochs=None,
                 log_directory=None, **kwargs):
        d = defaults["models"]
        if prior_probabilities_method is None:
            prior_probabilities_method = d["prior_probabilities_method"]
        self.prior_probabilities_method = prior_probabilities_method
        self.prior_probabilities = prior_probabilities
        if prior_probabilities_method == "custom" and prior_probabilities is None:
            raise TypeError("No prior probabilities supplied for custom method.")
        if proportion_of_free_nats_for_y_kl_divergence is None:
            proportion_of_free_nats_for_y_kl_divergence = d[
                "proportion_of_free_nats_for_y_kl_divergence"]
        self.proportion_of_free_nats_for_y_kl_divergence = (
            proportion_of_free_nats_for_y_kl_divergence)
        if number_of_latent_clusters is None and prior_probabilities is not None:
            number_of_latent_clusters = len(prior_probabilities)
        kwargs["number_of_latent_clusters"] = number_of_latent_clusters
        latent_distribution = kwargs.pop("latent_distribution", None)
        super().__init__(
            feature_size, latent_size=latent_size, hidden_sizes=hidden_sizes,
            reconstruction_distribution=reconstruction_distribution,
            number_of_reconstruction_classes=number_of_reconstruction_classes,
            latent_distribution=latent_distribution,
            minibatch_normalisation=minibatch_normalisation, batch_correction=batch_correction,
            number_of_batches=number_of_batches,
            number_of_warm_up_epochs=number_of_warm_up_epochs, log_directory=log_directory,
            **kwargs)
        self.n_clusters = self.number_of_latent_clusters
        # the flag only affects the model name for the GMVAE (quirk Q8)
        self.analytical_kl_term = bool(kwargs.get("analytical_kl_term") or False)
        if self.latent_distribution_name not in ("gaussian mixture", "full-covariance gaussian mixture"):
            raise NotImplementedError(
                "Not on the B200 hot path yet (SURVEY §8 f4): latent distribution `{}`.".format(
                    self.latent_distribution_name))

    def _latent_choices(self):
        return GMVAE_LATENT_DISTRIBUTIONS

    @property
    def name(self):
        """Same scheme as GMVAE:441-502."""
        latent = [normalise_string(self.latent_distribution_name),
                  "c_{}".format(self.number_of_latent_clusters)]
        if self.prior_probabilities_method != "uniform":
            latent.append("p_" + self.prior_probabilities_method)
        minor = [normalise_string(self.reconstruction_distribution_name)]
        if self.k_max:
            minor.append("k_{}".format(self.k_max))
        if self.use_count_sum_as_feature:
            minor.append("sum")
        minor.append("l_{}".format(self.latent_size))
        minor.append("h_" + "_".join(str(h) for h in self.hidden_sizes))
        minor.append("mc_{}".format(self.number_of_monte_carlo_samples["training"]))
        minor.append("iw_{}".format(self.number_of_importance_samples["training"]))
        if self.analytical_kl_term:
            minor.append("kl")
        if self.minibatch_normalisation:
            minor.append("bn")
        if self.batch_correction:
            minor.append("bc")
        if self.dropout_parts:
            minor.append("dropout_" + "_".join(self.dropout_parts))
        if self.kl_weight_value != 1:
            minor.append("klw_{}".format(self.kl_weight_value))
        if self.number_of_warm_up_epochs:
            minor.append("wu_{}".format(self.number_of_warm_up_epochs))
        if self.proportion_of_free_nats_for_y_kl_divergence:
            minor.append("fn_{}".format(self.proportion_of_free_nats_for_y_kl_divergence))
        return os.path.join(self.type, "-".join(latent), "-".join(minor))

    _description_hides_single_samples = True
    _description_kl_weight_label = "KL weight"
    _description_mentions_analytical_kl = False

    def _description_latent(self):
        lines = []
        if "mixture" in self.latent_distribution_name:
            lines.append("latent clusters: {}".format(self.number_of_latent_clusters))
            lines.append("prior probabilities: " + self.prior_probabilities_method)
        return lines

    def _description_training_terms(self):
        if self.proportion_of_free_nats_for_y_kl_divergence:
            return ["proportion of free nats for y KL divergence: {}".format(
                self.proportion_of_free_nats_for_y_kl_divergence)]
        return []

    # ------------------------------------------------------------------------------------------
    def _build_engine(self):
        from .gmvae_engine import GMVAEEngine
        return GMVAEEngine(
            self.feature_size, self.latent_size, self.number_of_latent_clusters, self.hidden_sizes,
            self.reconstruction_distribution_name, self.minibatch_normalisation,
            self.kl_weight_value, self.prior_probabilities_method, self.prior_probabilities,
            self.proportion_of_free_nats_for_y_kl_divergence, device=self._device,
            seed=self._seed, tensor_cores=self._tensor_cores,
            number_of_batches=self.number_of_batches if self.batch_correction else 0,
            count_sum_feature=bool(self.use_count_sum_as_feature),
            number_of_reconstruction_classes=self.k_max,
            dropout_keep_probabilities=self.dropout_keep_probabilities,
            latent_distribution=self.latent_distribution_name)

    def _evaluate_pass(self, engine, x_csr, t_csr, minibatch_size, R, S, deterministic=False,
                       seed=0, on_batch=None):
        """Forward-only pass; fetches what GMVAE:1229-1240 / 2438-2452 fetch."""
        import torch
        from .hotloop import ResidentCSR
        from . import kernels as K
        if t_csr is not None and t_csr is not x_csr:
            raise NotImplementedError("Separate preprocessed inputs are not supported by the "
                                      "GMVAE engine yet.")
        n = x_csr.shape[0]
        Kc, L = self.number_of_latent_clusters, self.latent_size
        dev = engine.device
        data = x_csr if isinstance(x_csr, ResidentCSR) else ResidentCSR(x_csr, dev)
        minibatch_size = min(minibatch_size, engine.max_single_chunk_minibatch(R * S))
        n_batches = -(-n // minibatch_size)
        log = torch.zeros(n_batches, 6, dtype=torch.float32, device=dev)
        stats = torch.zeros(n_batches, 2, Kc, L, dtype=torch.float32, device=dev)
        full_cov = bool(getattr(engine, "full_cov", False))
        q_cov = torch.zeros(n_batches, Kc, L, L, dtype=torch.float32, device=dev) if full_cov else None
        q_y_probabilities = torch.zeros(n_batches, Kc, dtype=torch.float32, device=dev)
        z_mean = torch.zeros(n, L, dtype=torch.float32, device=dev)
        y_mean = torch.zeros(n, Kc, dtype=torch.float32, device=dev)
        logits = torch.zeros(n, Kc, dtype=torch.float32, device=dev)
        for b, i in enumerate(range(0, n, minibatch_size)):
            rows = min(minibatch_size, n - i)
            plan = engine._plan(rows, R * S)
            idx = torch.arange(i, i + rows, dtype=torch.int64, device=dev)
            engine.set_batch_csr(plan, data.indptr, data.indices, data.values, idx)
            if engine.number_of_batches:
                K.gather_f32(data.batch_index, idx, plan.batch_index)
            if engine.count_sum_feature:
                K.gather_f32(data.count_sum_feature, idx, plan.count_sum)
            if engine.constrained:
                K.gather_f32(data.count_sum_parameter, idx, plan.count_sum_parameter)
            K.fill_normal(plan.eps, seed, b)
            engine.forward(plan, False, R, S, 1.0)
            log[b].copy_(plan.bound)
            z_mean[i:i + rows].copy_(engine.z_mean(plan))
            y_mean[i:i + rows].copy_(plan.y)
            logits[i:i + rows].copy_(plan.logits[:, :Kc])
            K.col_mean(plan.y, rows, Kc, q_y_probabilities[b])
            # q(z|x,y=k) means / variances averaged over the batch (GMVAE:2881-2885)
            qh = plan.QH.view(Kc, rows, -1)
            for k in range(Kc):
                K.col_mean(qh[k], rows, L, stats[b, 0, k])
            if full_cov:      # mean over the cells of S S^T; its diagonal = stddev^2 (GMVAE:2883-2893)
                K.gmvae_full_covariance_mean(plan.QH, Kc, rows, L, q_cov[b])
                stats[b, 1].copy_(torch.diagonal(q_cov[b], dim1=-2, dim2=-1))
            else:
                stats[b, 1].copy_(self._posterior_variances(plan, rows, L))
            if on_batch is not None:
                on_batch(plan, i, rows)
        log = log.cpu().numpy().astype(numpy.float64)
        divisor = n / minibatch_size
        stats = stats.cpu().numpy().astype(numpy.float64).sum(axis=0) / divisor
        p_mean, p_var, p_cov, _ = self._prior_moments(engine)
        result = {
            "lower_bound": log[:, 0].sum() / divisor,
            "reconstruction_error": log[:, 2].sum() / divisor,
            "kl_divergence_z": log[:, 3].sum() / divisor,
            "kl_divergence_y": log[:, 4].sum() / divisor,
            "z_mean": z_mean.cpu().numpy(), "y_mean": y_mean.cpu().numpy(),
            "q_y_logits": logits.cpu().numpy(),
            "q_y_probabilities": q_y_probabilities.cpu().numpy().sum(axis=0) / divisor,
            "q_z_means": stats[0], "q_z_variances": stats[1],
            "p_y_probabilities": numpy.exp(engine.log_py.cpu().numpy()),
            "p_z_means": p_mean, "p_z_variances": p_var,
        }
        if full_cov:
            result["q_z_covariances"] = q_cov.cpu().numpy().astype(numpy.float64).sum(axis=0) / divisor
            result["p_z_covariances"] = p_cov
        result["kl_divergence"] = result["kl_divergence_z"] + result["kl_divergence_y"]
        result["kl_divergence_neurons"] = numpy.array([result["kl_divergence"]])  # GMVAE:3401
        return result

    @staticmethod
    def _fill_triangular(scales):
        """tfp.distributions.fill_triangular (lower) on a (K, L (L + 1) / 2) numpy array."""
        m = scales.shape[-1]
        n = int(round((numpy.sqrt(8 * m + 1) - 1) / 2))
        full = numpy.concatenate([scales[..., n:], scales[..., ::-1]], axis=-1)
        return numpy.tril(full.reshape(scales.shape[:-1] + (n, n)))

    def _prior_moments(self, engine):
        """(means (K, L), variances (K, L), covariances (K, L, L) or None, scale matrices or
        None) of p(z|y=k) from the variables (GMVAE:2877-2893; K * L-sized host glue)."""
        prior = engine.export_parameters()
        softplus = lambda a: numpy.log1p(numpy.exp(-numpy.abs(a))) + numpy.maximum(a, 0)
        base = "Z/P/{}/".format(engine.z_scope)

        def head(name):
            return (prior[base + name + "/DENSE/weights"] + prior[base + name + "/DENSE/biases"]).numpy()
        means = head(engine.z_heads[0])
        if engine.full_cov:
            tril = self._fill_triangular(numpy.maximum(softplus(head("SCALES").astype(numpy.float64)),
                                                       numpy.finfo(numpy.float32).tiny))
            cov = tril @ numpy.swapaxes(tril, -1, -2)
            return means, numpy.diagonal(cov, axis1=-2, axis2=-1).copy(), cov, tril
        return means, softplus(head("SOFTPLUS_SCALE")), None, None

    @staticmethod
    def _posterior_variances(plan, rows, L):
        """mean_b softplus(s_k[b]) per cluster (K, L): a K*L-sized statistic that only feeds the
        TensorBoard centroid tags (GMVAE:2883-2885), computed with torch on the device."""
        import torch
        s = plan.QH.view(-1, rows, plan.QH.shape[1])
        return torch.nn.functional.softplus(s[:, :, L:2 * L]).mean(dim=1)

    # ------------------------------------------------------------------------------------------
    def _result_string(self, result):
        text = "ELBO: {:.5g}, ENRE: {:.5g}, KL_z: {:.5g}, KL_y: {:.5g}".format(
            result["lower_bound"], result["reconstruction_error"], result["kl_divergence_z"],
            result["kl_divergence_y"])
        if result.get("accuracy") is not None:
            text += ", Acc: {:.5g}".format(result["accuracy"])
        return text

    def _summary_scalars(self, result, data_set, with_centroids):
        if data_set is not None and data_set.has_labels and "accuracy" not in result:
            names = numpy.unique(data_set.labels)
            label_ids = numpy.searchsorted(names, data_set.labels)
            cluster_ids = result["q_y_logits"].argmax(axis=1)
            excluded = [i for i, name in enumerate(names)
                        if name in getattr(data_set, "excluded_classes", [])]
            predicted = map_cluster_ids_to_label_ids(label_ids, cluster_ids, excluded)
            result["accuracy"] = accuracy(label_ids, predicted, excluded)
            result["cluster_ids"] = cluster_ids
            result["predicted_labels"] = names[predicted]
        scalars = {"losses/" + key: result[key] for key in self._loss_keys}
        if result.get("accuracy") is not None:
            scalars["accuracy"] = result["accuracy"]
        scalars["kl_divergence_neurons/0"] = result["kl_divergence"]
        if with_centroids:
            scalars.update(self._centroid_scalars(result))
        return scalars

    def _centroid_scalars(self, result):
        scalars = {}
        for k in range(self.number_of_latent_clusters):
            scalars["prior/cluster_{}/probability".format(k)] = result["p_y_probabilities"][k]
            scalars["posterior/cluster_{}/probability".format(k)] = result["q_y_probabilities"][k]
            for l in range(self.latent_size):
                for dist, key in (("prior", "p_z"), ("posterior", "q_z")):
                    scalars["{}/cluster_{}/mean/dimension_{}".format(dist, k, l)] = \
                        result[key + "_means"][k, l]
                    scalars["{}/cluster_{}/variance/dimension_{}".format(dist, k, l)] = \
                        result[key + "_variances"][k, l]
                    if key + "_covariances" in result:          # GMVAE:1406-1417
                        for l_ in range(self.latent_size):
                            scalars["{}/cluster_{}/covariance/dimension_{}_{}".format(dist, k, l, l_)] = \
                                result[key + "_covariances"][k, l, l_]
        return scalars

    def _centroids(self, result):
        def pack(probabilities, means, variances, covariances):
            if covariances is None:                             # GMVAE:1837-1859
                covariances = numpy.stack([numpy.diag(v) for v in variances])
            return {"probabilities": probabilities, "means": means,
                    "covariance_matrices": covariances}
        return {"prior": pack(result["p_y_probabilities"], result["p_z_means"],
                              result["p_z_variances"], result.get("p_z_covariances")),
                "posterior": pack(result["q_y_probabilities"], result["q_z_means"],
                                  result["q_z_variances"], result.get("q_z_covariances"))}

    def _latent_sets(self, evaluation_set, result, common):
        L, Kc = self.latent_size, self.number_of_latent_clusters
        z = DataSet(evaluation_set.name, values=result["z_mean"], version="z", feature_names=(
            numpy.array(["z variable {}".format(i + 1) for i in range(L)])), **common)
        y = DataSet(evaluation_set.name, values=result["y_mean"], version="y", feature_names=(
            numpy.array(["y variable {}".format(i + 1) for i in range(Kc)])), **common)
        for subset in (z, y):
            subset.predicted_cluster_ids = result.get("cluster_ids",
                                                      result["q_y_logits"].argmax(axis=1))
            subset.predicted_labels = result.get("predicted_labels")
        return {"z": z, "y": y}

    # ------------------------------------------------------------------------------------------
    def sample(self, sample_size=None, minibatch_size=None, run_id=None,
               use_early_stopping_model=False, use_best_model=False, **kwargs):
        """y ~ p(y), z ~ p(z|y), x = E[x|z] (GMVAE:1949-2160): returns
        ``(sample DataSet, {"z": ..., "y": ...})``."""
        import torch
        from . import kernels as K
        if sample_size is None:
            sample_size = defaults["models"]["sample_size"]
        if minibatch_size is None:
            minibatch_size = defaults["models"]["minibatch_size"]
        if self.batch_correction or self.use_count_sum_as_parameter or self.use_count_sum_as_feature:
            # a free sample has no batch index / count sum to feed (as for the VAE, VAE:1639-1649)
            raise NotImplementedError("Sampling with batch correction or count sums is not "
                                      "implemented (as in the reference).")
        engine, _, _ = self._load_for_inference(run_id, use_early_stopping_model, use_best_model,
                                                "sample from")
        L, G, Kc = self.latent_size, self.feature_size, self.number_of_latent_clusters
        rng = numpy.random.RandomState(kwargs.get("noise_seed", 11))
        p_y = numpy.exp(engine.log_py.cpu().numpy().astype(numpy.float64))
        clusters = rng.choice(Kc, size=sample_size, p=p_y / p_y.sum())
        means, variances, _, tril = self._prior_moments(engine)
        noise = rng.standard_normal((sample_size, L))
        if tril is not None:          # z = loc + S eps
            z = (means[clusters] + numpy.einsum("nij,nj->ni", tril[clusters], noise)).astype(numpy.float32)
        else:
            z = (means[clusters] + numpy.sqrt(variances)[clusters] * noise).astype(numpy.float32)
        x = numpy.empty((sample_size, G), numpy.float32)
        from .engine import VAEEngine
        for i in range(0, sample_size, minibatch_size):
            rows = min(minibatch_size, sample_size - i)
            plan = engine._plan(rows, 1)
            plan.Z[:rows, :L].copy_(torch.from_numpy(z[i:i + rows]))
            VAEEngine.decode(engine, plan, rows)       # decoder only, moving statistics, 1 group
            outs = [torch.empty(rows, engine.Gn, dtype=torch.float32, device=engine.device)
                    for _ in range(3)]
            if engine.k_max:       # mean of the Categorised distribution (P_K class heads, CAT:210-247)
                K.piecewise_moments(engine.kind, engine.k_max, plan.A[:rows], engine.Gn, rows, G, 1,
                                    *outs)
            elif engine.continuous:
                K.continuous_moments(engine.kind, plan.A[:rows], engine.Gn, rows, G, 1, 1, None,
                                     *outs)
            else:
                K.likelihood_moments(engine.kind, plan.A[:rows], engine.Gn, rows, G, 1, 1, None,
                                     *outs)
            x[i:i + rows] = outs[0][:, :G].cpu().numpy()
        names = numpy.array(["example {}".format(i + 1) for i in range(sample_size)])
        y = numpy.eye(Kc, dtype=numpy.float32)[clusters]
        sample_set = DataSet("sample", values=x, example_names=names, kind="sample")
        latent = {
            "z": DataSet("sample", values=z, example_names=names, kind="sample", version="z",
                         feature_names=numpy.array(["z variable {}".format(i + 1) for i in range(L)])),
            "y": DataSet("sample", values=y, example_names=names, kind="sample", version="y",
                         feature_names=numpy.array(["y variable {}".format(i + 1) for i in range(Kc)])),
        }
        return sample_set, latent
