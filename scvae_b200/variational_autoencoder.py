"""``VariationalAutoencoder``: the reference's model class
(scvae/models/variational_autoencoder.py:47) re-hosted on the B200 step engine.

Same constructor keywords, ``train`` / ``evaluate`` / ``sample`` signatures, return values,
log-directory layout, checkpoint naming and TensorBoard tags; the TensorFlow graph, session
and saver underneath are replaced by ``scvae_b200.engine.VAEEngine`` (hand-written sm_100a
kernels behind ``libscvae_b200.so``).  Options of the reference that are outside the hot-path
scope of this round (the continuous likelihoods, dropout for the GMVAE) raise
``NotImplementedError`` instead of silently degrading.
"""

import copy
import os
import shutil
from time import time

import numpy
import scipy.sparse

from .data_set import DataSet
from .defaults import defaults
from .model_utilities import (
    SummaryWriter, build_training_string, check_run_id, checkpoint_epoch, clear_log_directory,
    copy_model_directory, early_stopping_status, format_duration, format_time,
    generate_unique_run_id_for_model, get_checkpoint_state, load_learning_curves,
    normalise_string, parse_numbers_of_samples, read_checkpoint, remove_old_checkpoints,
    validate_model_parameters, write_checkpoint)

RECONSTRUCTION_DISTRIBUTIONS = [
    "poisson", "negative binomial", "zero-inflated poisson", "zero-inflated negative binomial",
    "constrained poisson",
    # continuous / binary distributions (DU:31-73, :125-245; csrc/continuous.cu)
    "bernoulli", "lomax", "gaussian", "softplus gaussian", "modified gaussian", "log-normal", "gamma",
    "exponentially_modified_gaussian",
    # known to the reference, not on the B200 hot path (a categorical / multinomial over genes):
    "categorical", "multinomial",
]
VAE_LATENT_DISTRIBUTIONS = ["gaussian", "unit-variance gaussian"]


def parse_distribution(name, choices, kind):
    key = normalise_string(name)
    for choice in choices:
        if normalise_string(choice) == key:
            return choice
    raise ValueError("{} distribution `{}` not supported.".format(kind.capitalize(), name))


class VariationalAutoencoder:
    """Variational auto-encoder for count data (see module docstring)."""

    def __init__(self, feature_size, latent_size=None, hidden_sizes=None,
                 reconstruction_distribution=None, number_of_reconstruction_classes=None,
                 latent_distribution=None, minibatch_normalisation=None, batch_correction=None,
                 number_of_batches=None, number_of_warm_up_epochs=None, log_directory=None,
                 **kwargs):
        d = defaults["models"]
        self.type = self._model_type
        self.feature_size = feature_size
        self.latent_size = d["latent_size"] if latent_size is None else latent_size
        self.hidden_sizes = list(d["hidden_sizes"] if hidden_sizes is None else hidden_sizes)
        if reconstruction_distribution is None:
            reconstruction_distribution = d["reconstruction_distribution"]
        self.reconstruction_distribution_name = parse_distribution(
            reconstruction_distribution, RECONSTRUCTION_DISTRIBUTIONS, "reconstruction")
        if number_of_reconstruction_classes is None:
            number_of_reconstruction_classes = d["number_of_reconstruction_classes"]
        self.number_of_reconstruction_classes = number_of_reconstruction_classes + 1
        self.k_max = number_of_reconstruction_classes
        if latent_distribution is None:
            latent_distribution = d["latent_distribution"][self.type]
        self.latent_distribution_name = parse_distribution(
            latent_distribution, self._latent_choices(), "latent")
        self._parse_common_keywords(kwargs, minibatch_normalisation, batch_correction,
                                    number_of_batches, number_of_warm_up_epochs, log_directory)
        if kwargs.get("analytical_kl_term") is None:
            self.analytical_kl_term = self.latent_distribution_name == "gaussian"
        else:
            self.analytical_kl_term = bool(kwargs["analytical_kl_term"])
        self.early_stopping_rounds = 10
        self.stopped_early = None
        validate_model_parameters(
            reconstruction_distribution=self.reconstruction_distribution_name,
            number_of_reconstruction_classes=self.k_max, model_type=self.type,
            latent_distribution=self.latent_distribution_name,
            parameterise_latent_posterior=self.parameterise_latent_posterior)
        self._check_supported()
        self._engine = None
        self._device = kwargs.get("device", "cuda")
        self._tensor_cores = kwargs.get("tensor_cores", True)
        self._seed = kwargs.get("seed", 0)

    # ------------------------------------------------------------------------------------------
    def _latent_choices(self):
        return VAE_LATENT_DISTRIBUTIONS

    def _parse_common_keywords(self, kwargs, minibatch_normalisation, batch_correction,
                               number_of_batches, number_of_warm_up_epochs, log_directory):
        d = defaults["models"]

        def option(name, value=None):
            value = kwargs.get(name) if value is None else value
            return d[name] if value is None else value

        self.parameterise_latent_posterior = option("parameterise_latent_posterior")
        clusters = kwargs.get("number_of_latent_clusters")
        if clusters is None:
            clusters = d["number_of_classes"] if "mixture" in self.latent_distribution_name else 1
        self.number_of_latent_clusters = clusters
        for key in ("number_of_monte_carlo_samples", "number_of_importance_samples"):
            value = kwargs.get(key)
            value = copy.deepcopy(d["number_of_samples"]) if value is None \
                else parse_numbers_of_samples(value)
            setattr(self, key, value)
        self.inference_architecture = option("inference_architecture").upper()
        self.generative_architecture = option("generative_architecture").upper()
        if self.type != "VAE":
            # the reference's GMVAE has no linear-factor form: it accepts these keywords (the
            # CLI passes them to both model types) and builds the MLP graph regardless
            # (GMVAE:2788-3221; recorded in tests/golden/model_names.json)
            self.inference_architecture = self.generative_architecture = "MLP"
        self.minibatch_normalisation = option("minibatch_normalisation", minibatch_normalisation)
        self.batch_correction = option("batch_correction", batch_correction)
        if self.batch_correction and number_of_batches is None:
            raise TypeError("The number of batches for batch correction was not provided.")
        self.number_of_batches = number_of_batches
        keep = option("dropout_keep_probabilities")
        self.dropout_keep_probabilities = keep
        values = list(keep) if isinstance(keep, (list, tuple)) else [keep]
        self.dropout_parts = [str(p) for p in values if p and p != 1]
        self.use_count_sum_as_feature = option("count_sum")
        self.use_count_sum_as_parameter = (
            "constrained" in self.reconstruction_distribution_name
            or "multinomial" in self.reconstruction_distribution_name)
        self.kl_weight_value = option("kl_weight")
        self.number_of_warm_up_epochs = option("number_of_warm_up_epochs",
                                               number_of_warm_up_epochs)
        self.base_log_directory = d["directory"] if log_directory is None else log_directory

    def _check_supported(self):
        from .kernels import LIKELIHOOD_KINDS
        problems = []
        if self.reconstruction_distribution_name not in LIKELIHOOD_KINDS:
            problems.append("reconstruction distribution `{}`".format(
                self.reconstruction_distribution_name))
        if self.k_max and self.reconstruction_distribution_name == "constrained poisson":
            problems.append("piecewise-categorical likelihoods (number_of_reconstruction_classes) "
                            "around the constrained Poisson")
        if self.inference_architecture not in ("MLP", "LFM") \
                or self.generative_architecture not in ("MLP", "LFM"):
            raise ValueError("The inference and generative architectures can only be a neural "
                             "network (MLP) or a linear factor model (LFM).")
        if self.use_count_sum_as_parameter \
                and self.reconstruction_distribution_name != "constrained poisson":
            problems.append("count-sum-parameterised likelihoods other than the constrained "
                            "Poisson")
        if self.parameterise_latent_posterior:
            problems.append("parameterised latent posteriors")
        if problems:
            raise NotImplementedError(
                "Not on the B200 hot path yet (SURVEY §8 f3): " + "; ".join(problems) + ".")

    # ------------------------------------------------------------------------------------------
    @property
    def name(self):
        """Short name used in directory names (same scheme as VAE:412-469)."""
        major = [normalise_string(self.latent_distribution_name)]
        if "mixture" in self.latent_distribution_name:
            major.append("c_{}".format(self.number_of_latent_clusters))
        if self.parameterise_latent_posterior:
            major.append("parameterised")
        if self.inference_architecture != "MLP":
            major.append("ia_{}".format(self.inference_architecture))
        if self.generative_architecture != "MLP":
            major.append("ga_{}".format(self.generative_architecture))
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
        return os.path.join(self.type, "-".join(major), "-".join(minor))

    @property
    def description(self):
        """Model summary printed by the CLI; wording and order follow VAE:471-553 (and, through
        the ``_description_*`` hooks, GMVAE:505-592), typos included ("KL weigth")."""
        lines = ["Model setup:", "type: {}".format(self.type),
                 "feature size: {}".format(self.feature_size),
                 "latent size: {}".format(self.latent_size),
                 "hidden sizes: {}".format(", ".join(map(str, self.hidden_sizes))),
                 "latent distribution: " + self.latent_distribution_name]
        lines += self._description_latent()
        lines.append("reconstruction distribution: " + self.reconstruction_distribution_name)
        if self.k_max > 0:
            lines.append("reconstruction classes: {} (including 0s)".format(self.k_max))
        for label, numbers in (("Monte Carlo samples", self.number_of_monte_carlo_samples),
                               ("importance samples", self.number_of_importance_samples)):
            if self._description_hides_single_samples and max(numbers.values()) <= 1:
                continue
            text = "{}: {}".format(label, numbers["training"])
            if numbers["evaluation"] != numbers["training"]:
                text += " (training), {} (evaluation)".format(numbers["evaluation"])
            lines.append(text)
        if self.kl_weight_value != 1:
            lines.append("{}: {}".format(self._description_kl_weight_label, self.kl_weight_value))
        if self._description_mentions_analytical_kl and self.analytical_kl_term:
            lines.append("using analytical KL term")
        if self.minibatch_normalisation:
            lines.append("using batch normalisation for minibatches")
        if self.batch_correction:
            lines.append("with batch correction")
        if self.number_of_warm_up_epochs:
            lines.append("using linear warm-up weighting for the first {} epochs".format(
                self.number_of_warm_up_epochs))
        lines += self._description_training_terms()
        if self.dropout_parts:
            lines.append("dropout keep probability: {}".format(", ".join(self.dropout_parts)))
        if self.use_count_sum_as_feature:
            lines.append("using count sums")
        if self.early_stopping_rounds:
            lines.append("early stopping: after {} epoch with no improvements".format(
                self.early_stopping_rounds))
        return "\n    ".join(lines)

    _description_hides_single_samples = False
    _description_kl_weight_label = "KL weigth"
    _description_mentions_analytical_kl = True

    def _description_latent(self):
        lines = []
        if "mixture" in self.latent_distribution_name:
            lines.append("latent clusters: {}".format(self.number_of_latent_clusters))
        if self.parameterise_latent_posterior:
            lines.append("using parameterisation of latent posterior parameters")
        return lines

    def _description_training_terms(self):
        return []

    @property
    def parameters(self):
        engine = self._get_engine()
        rows = [(k, tuple(v.shape)) for k, v in engine.export_parameters().items()
                if not k.endswith(("moving_mean", "moving_variance"))]
        width = max(len(k) for k, _ in rows)
        return "\n    ".join(["Trainable parameters"] + [
            "{:{}}  {}".format(k, width, shape) for k, shape in rows])

    def log_directory(self, base=None, run_id=None, early_stopping=False, best_model=False):
        directory = os.path.join(base or self.base_log_directory, self.name)
        if run_id is None:
            run_id = defaults["models"]["run_id"]
        if run_id:
            directory = os.path.join(directory, "run_{}".format(check_run_id(run_id)))
        if early_stopping and best_model:
            raise ValueError("Early-stopping model and best model are mutually exclusive.")
        if early_stopping:
            directory = os.path.join(directory, "early_stopping")
        elif best_model:
            directory = os.path.join(directory, "best")
        return directory

    def has_been_trained(self, run_id=None):
        return bool(get_checkpoint_state(self.log_directory(run_id=run_id)))

    def early_stopping_status(self, run_id=None):
        es_directory = self.log_directory(run_id=run_id, early_stopping=True)
        directory = os.path.dirname(es_directory)
        if os.path.exists(directory) and os.path.exists(es_directory):
            losses = load_learning_curves(self, "validation", run_id=run_id,
                                          log_directory=directory)["lower_bound"]
            return early_stopping_status(losses, self.early_stopping_rounds)
        return False, 0

    # ------------------------------------------------------------------------------------------
    # engine plumbing
    # ------------------------------------------------------------------------------------------
    def _build_engine(self):
        from .engine import VAEEngine
        return VAEEngine(self.feature_size, self.latent_size, self.hidden_sizes,
                         self.reconstruction_distribution_name, self.latent_distribution_name,
                         self.minibatch_normalisation, self.kl_weight_value,
                         device=self._device, seed=self._seed, tensor_cores=self._tensor_cores,
                         number_of_batches=self.number_of_batches if self.batch_correction else 0,
                         count_sum_feature=bool(self.use_count_sum_as_feature),
                         inference_architecture=self.inference_architecture,
                         generative_architecture=self.generative_architecture,
                         number_of_reconstruction_classes=self.k_max,
                         analytical_kl_term=self.analytical_kl_term,
                         dropout_keep_probabilities=self.dropout_keep_probabilities)

    def _attach_features(self, data, data_set):
        """Per-cell decoder features of a data set (VAE:816-833): batch indices for batch
        correction, the normalised count sum as a feature."""
        batch_indices = count_sum = None
        if self.batch_correction:
            batch_indices = data_set.batch_indices
            if batch_indices is None:
                raise TypeError("No batch indices found in {} set.".format(data_set.kind))
        if self.use_count_sum_as_feature:
            count_sum = data_set.normalised_count_sum
        parameter = data_set.count_sum if self.use_count_sum_as_parameter else None
        if batch_indices is not None or count_sum is not None or parameter is not None:
            data.set_features(batch_indices, count_sum, parameter)
        return data

    def _get_engine(self):
        if self._engine is None:
            self._engine = self._build_engine()
        return self._engine

    def _samples(self, scenario):
        return (self.number_of_importance_samples[scenario],
                self.number_of_monte_carlo_samples[scenario])

    @staticmethod
    def _inputs(data_set, reconstruction_distribution_name):
        """x (network input) and t (likelihood target) matrices as the reference picks them
        (VAE:849-861): preprocessed values if present, raw values as targets."""
        if data_set.noisy_preprocess:
            raise NotImplementedError("Noisy preprocessing is not on the B200 hot path yet.")
        x = data_set.preprocessed_values if data_set.has_preprocessed_values else data_set.values
        t = data_set.binarised_values if reconstruction_distribution_name == "bernoulli" \
            else data_set.values
        return x, t

    def _restore(self, engine, log_directory):
        state = get_checkpoint_state(log_directory)
        if not state:
            return None
        engine.load_state_dict(read_checkpoint(state)["engine"])
        return checkpoint_epoch(state)

    # ------------------------------------------------------------------------------------------
    # forward passes over a whole data set (per-epoch evaluation, evaluate())
    # ------------------------------------------------------------------------------------------
    def _evaluate_pass(self, engine, x_csr, t_csr, minibatch_size, R, S, deterministic=False,
                       seed=0, on_batch=None, total_examples=None):
        """Forward-only pass in data order.  Aggregation follows the reference:
        per-batch means are summed and divided by N / B.  ``total_examples``: the matrix is this
        rank's shard of a data set of that many cells -- the per-batch sums are added over the
        ranks before the division."""
        import torch
        from .hotloop import ResidentCSR
        from . import kernels as K
        from .hotloop import PackedStream
        n = x_csr.shape[0]
        L = self.latent_size
        dev = engine.device
        packed = x_csr if isinstance(x_csr, PackedStream) else None
        if packed is not None:
            # host-resident matrix: the rows stream through in data order as packed slabs
            if on_batch is not None or (t_csr is not None and t_csr is not x_csr):
                raise NotImplementedError("host-resident data feeds the lean evaluation passes only")
            if packed.B != minibatch_size:
                raise ValueError("the packed stream was built for minibatches of {} rows".format(packed.B))
            packed.pack_epoch(None)          # (starts the feeder thread: slabs in data order)
        data = packed if packed is not None else (
            x_csr if isinstance(x_csr, ResidentCSR) else ResidentCSR(x_csr, dev))
        targets = None
        if t_csr is not None and t_csr is not x_csr:
            targets = t_csr if isinstance(t_csr, ResidentCSR) else ResidentCSR(t_csr, dev)
        n_batches = -(-n // minibatch_size)
        log = torch.zeros(n_batches, 4 + L, dtype=torch.float32, device=dev)
        q_z_mean = torch.zeros(n, L, dtype=torch.float32, device=dev)
        for b, i in enumerate(range(0, n, minibatch_size)):
            rows = min(minibatch_size, n - i)
            plan = engine._plan(rows, R * S)
            idx = torch.arange(i, i + rows, dtype=torch.int64, device=dev)
            # no per-batch consumer of the head pre-activations: 16-bit minibatch + forward-only
            # fused heads (when the shapes allow; otherwise this is the fp32 path as before)
            lean = on_batch is None and targets is None
            if packed is not None:
                slot = packed.fetch(b % 2, b)
                torch.cuda.current_stream().wait_event(slot["ready"])
                split = slot.get("split", 0)
                if split:        # hybrid feeder: two slabs per minibatch
                    engine.set_batch_packed(plan, slot["buf"], f16_exact=packed.f16_exact, rows=(0, split))
                    engine.set_batch_packed(plan, slot["buf_host"], f16_exact=packed.f16_exact,
                                            rows=(split, rows))
                else:
                    engine.set_batch_packed(plan, slot["buf"], f16_exact=packed.f16_exact)
                slot["free"].record(torch.cuda.current_stream())
            else:
                engine.set_batch_csr(plan, data.indptr, data.indices, data.values, idx,
                                     u16_ok=data.u16_ok, f16_exact=data.f16_exact, train16=lean,
                                     row_const_all=data.row_const)
            if targets is not None:
                if plan.T is None:
                    plan.T = torch.zeros(rows, engine.Gp, dtype=torch.float32, device=dev)
                K.csr_densify(targets.indptr, targets.indices, targets.values, idx, engine.G,
                              plan.T, plan.row_const)
                plan.use_T = True
            if getattr(engine, "constrained", False):
                K.gather_f32(data.count_sum_parameter, idx, plan.count_sum_parameter)
            if getattr(engine, "n_extra", 0):
                if engine.number_of_batches:
                    K.gather_f32(data.batch_index, idx, plan.batch_index)
                if engine.count_sum_feature:
                    K.gather_f32(data.count_sum_feature, idx, plan.count_sum)
            if not deterministic:
                K.fill_normal(plan.eps, seed, b)
            engine.forward(plan, False, R, S, 1.0, deterministic=deterministic,
                           keep_heads=not lean)
            log[b, :4].copy_(plan.bound)
            log[b, 4:].copy_(engine.kl_neurons(plan))
            q_z_mean[i:i + rows].copy_(plan.PH[:rows, :L])
            if on_batch is not None:
                on_batch(plan, i, rows)
        sums = log.sum(dim=0, dtype=torch.float64)
        divisor = n / minibatch_size
        if total_examples is not None:
            from . import distributed as D
            D.all_reduce_sum_(sums)
            divisor = total_examples / minibatch_size
        sums = sums.cpu().numpy()
        return {
            "lower_bound": sums[0] / divisor,
            "reconstruction_error": sums[2] / divisor,
            "kl_divergence": sums[3] / divisor,
            "kl_divergence_neurons": sums[4:] / divisor,
            "q_z_mean": q_z_mean.cpu().numpy(),
        }

    # ------------------------------------------------------------------------------------------
    # train
    # ------------------------------------------------------------------------------------------
    def train(self, training_set, validation_set=None, number_of_epochs=None,
              minibatch_size=None, learning_rate=None, run_id=None, new_run=None,
              reset_training=None, **kwargs):
        """Train for ``number_of_epochs`` epochs (resuming from the newest checkpoint); returns 0
        like the reference (VAE:640-1599)."""
        import torch
        from .hotloop import ResidentCSR, TrainLoop

        d = defaults["models"]
        number_of_epochs = d["number_of_epochs"] if number_of_epochs is None else number_of_epochs
        minibatch_size = d["minibatch_size"] if minibatch_size is None else minibatch_size
        learning_rate = d["learning_rate"] if learning_rate is None else learning_rate
        run_id = d["run_id"] if run_id is None else run_id
        new_run = d["new_run"] if new_run is None else new_run
        reset_training = d["reset_training"] if reset_training is None else reset_training
        start_time = time()
        if run_id:
            run_id = check_run_id(run_id)
        elif new_run:
            run_id = generate_unique_run_id_for_model(self, timestamp=start_time)
        model_string = "model for run {}".format(run_id) if run_id else "model"

        permanent_log_directory = self.log_directory(run_id=run_id)
        if reset_training and os.path.exists(permanent_log_directory):
            clear_log_directory(permanent_log_directory)
        metadata_log = {"epochs trained": None, "start time": format_time(start_time),
                        "training duration": None, "last epoch duration": None,
                        "learning rate": learning_rate, "minibatch size": minibatch_size}
        state = get_checkpoint_state(permanent_log_directory)
        epoch_start = checkpoint_epoch(state) if state else 0

        temporary_log_directory = kwargs.get("temporary_log_directory")
        base = temporary_log_directory or None
        log_directory = self.log_directory(base=base, run_id=run_id)
        early_stopping_log_directory = self.log_directory(base=base, run_id=run_id,
                                                          early_stopping=True)
        best_model_log_directory = self.log_directory(base=base, run_id=run_id, best_model=True)
        if temporary_log_directory:
            temporary_state = get_checkpoint_state(log_directory)
            temporary_epoch = checkpoint_epoch(temporary_state) if temporary_state else 0
            replace_temporary = temporary_epoch <= epoch_start
            epoch_start = max(epoch_start, temporary_epoch)

        data_string = "{} set".format(training_set.kind)
        print(build_training_string(model_string, epoch_start, number_of_epochs, data_string))
        if epoch_start >= number_of_epochs:
            return 0
        if (temporary_log_directory and os.path.exists(permanent_log_directory)
                and replace_temporary):
            if os.path.exists(log_directory):
                shutil.rmtree(log_directory)
            shutil.copytree(permanent_log_directory, log_directory)

        R, S = self._samples("training")
        if self.type == "VAE":   # the GMVAE does not rescale
            minibatch_size = int(numpy.ceil(minibatch_size / (R * S)))

        engine = self._get_engine()
        # data-parallel: active iff a torch.distributed process group exists
        from . import distributed as D
        rank, world = D.rank(), D.world_size()
        is_main = rank == 0
        D.attach(engine)
        x_train, t_train = self._inputs(training_set, self.reconstruction_distribution_name)
        n_train = training_set.number_of_examples
        minibatch_size = min(minibatch_size, n_train)
        # data_residency="host": the count matrix stays in (pinned) host memory and every step's rows
        # travel as one packed slab (hotloop.PackedStream) -- for matrices beyond HBM
        host_resident = kwargs.get("data_residency", "device") == "host"
        sharding = kwargs.get("data_sharding", "replicated")
        if sharding not in ("replicated", "rank"):
            raise ValueError("`data_sharding` is `replicated` or `rank`.")
        rank_sharded = world > 1 and sharding == "rank" and not host_resident
        stream = eval_stream = None
        if host_resident:
            from .hotloop import PackedStream
            if t_train is not x_train or getattr(engine, "n_extra", 0) or getattr(engine, "constrained", False):
                raise NotImplementedError("host-resident training data: counts only (no separate "
                                          "targets, batch correction or count-sum inputs)")
            csr_train = scipy.sparse.csr_matrix(x_train, dtype=numpy.float32)
            per_rank = minibatch_size // world if world > 1 else minibatch_size
            stream = PackedStream(csr_train, engine.device, per_rank)
            eval_stream = PackedStream(csr_train, engine.device, minibatch_size)
            data = eval_stream
        elif rank_sharded:
            # data_sharding="rank" (SURVEY §8e): rank r keeps the cells r::W in HBM and nothing
            # else; every step it takes its share of the minibatch from a shuffle of ITS cells
            if self.type != "VAE" or t_train is not x_train or getattr(engine, "n_extra", 0) \
                    or getattr(engine, "constrained", False):
                raise NotImplementedError("rank-sharded training data: VAE on counts only (no "
                                          "separate targets, batch correction or count-sum inputs)")
            local = scipy.sparse.csr_matrix(x_train, dtype=numpy.float32)[rank::world]
            data = ResidentCSR(local, engine.device)
            n_local = local.shape[0]
            if n_local < -(-n_train // minibatch_size):
                raise ValueError("rank-sharded training data: fewer cells per rank than steps")
            local_rng = numpy.random.RandomState(
                (int(kwargs.get("shuffle_seed", 0)) + 1) * 7919 + rank)
        else:
            data = ResidentCSR(scipy.sparse.csr_matrix(x_train, dtype=numpy.float32), engine.device)
            self._attach_features(data, training_set)
        if t_train is not x_train:
            # targets other than the network input (binarised values for the Bernoulli likelihood,
            # raw counts behind preprocessed inputs; VAE:849-861): a second resident matrix
            data.targets = ResidentCSR(scipy.sparse.csr_matrix(t_train, dtype=numpy.float32),
                                       engine.device)
            data.u16_ok = False             # (the 16-bit fused path reads x as its own target)
        if validation_set:
            x_valid, t_valid = self._inputs(validation_set, self.reconstruction_distribution_name)
            valid_data = ResidentCSR(scipy.sparse.csr_matrix(x_valid, dtype=numpy.float32),
                                     engine.device)
            self._attach_features(valid_data, validation_set)
            if t_valid is not x_valid:
                valid_data.targets = ResidentCSR(scipy.sparse.csr_matrix(t_valid, dtype=numpy.float32),
                                                 engine.device)
                valid_data.u16_ok = False

        training_writer = SummaryWriter(os.path.join(log_directory, "training")) \
            if is_main else None
        validation_writer = SummaryWriter(os.path.join(log_directory, "validation")) \
            if (validation_set and is_main) else None

        restored = self._restore(engine, log_directory)
        if restored is not None:
            epoch_start = restored
            if validation_set:
                curve = load_learning_curves(self, "validation", run_id=run_id,
                                             log_directory=log_directory)["lower_bound"]
                lower_bound_valid_maximum = curve.max()
                self.stopped_early, epochs_with_no_improvement = self.early_stopping_status(
                    run_id=run_id)
                if numpy.isnan(epochs_with_no_improvement):
                    lower_bound_valid_early_stopping = curve[-1]
                else:
                    lower_bound_valid_early_stopping = curve[-1 - epochs_with_no_improvement]
        else:
            engine.initialise(self._seed)
            epoch_start = 0
            if validation_set:
                lower_bound_valid_maximum = -numpy.inf
                epochs_with_no_improvement = 0
                lower_bound_valid_early_stopping = -numpy.inf
                self.stopped_early = False
        metadata_log["epochs trained"] = (epoch_start, number_of_epochs)
        D.broadcast_parameters(engine)

        rng = numpy.random.RandomState(kwargs["shuffle_seed"]) if "shuffle_seed" in kwargs \
            else numpy.random
        loops = {}
        noise_seed = kwargs.get("noise_seed", 1)
        use_graph = kwargs.get("use_cuda_graph", True)
        learning_curves = {"training": {k: [] for k in self._loss_keys}}
        if validation_set:
            learning_curves["validation"] = copy.deepcopy(learning_curves["training"])
        training_time_start = time()
        epoch_duration = 0.0

        for epoch in range(epoch_start, number_of_epochs):
            epoch_time_start = time()
            if self.number_of_warm_up_epochs:
                warm_up_weight = float(min(epoch / self.number_of_warm_up_epochs, 1.0))
            else:
                warm_up_weight = 1.0
            if world > 1:
                shuffled = D.broadcast_permutation(n_train, rng, engine.device)
            else:
                shuffled = torch.from_numpy(numpy.asarray(rng.permutation(n_train))).to(
                    engine.device)
            n_steps = -(-n_train // minibatch_size)
            step_bounds = torch.zeros(n_steps, 8, dtype=torch.float32, device=engine.device)
            if stream is not None:
                # this rank's rows of every global minibatch, in step order, as packed slabs
                order_host = shuffled.cpu().numpy()
                pieces = []
                for i in range(0, n_train, minibatch_size):
                    rows = min(minibatch_size, n_train - i)
                    lo, hi = D.shard_bounds(i, rows, rank, world) if world > 1 else (i, i + rows)
                    pieces.append(order_host[lo:hi])
                order_rank = numpy.concatenate(pieces)
                stream.pack_epoch(order_rank)       # (uneven last steps only without ranks)
                compute = torch.cuda.current_stream()
                pending = stream.fetch(0, 0)
                for s in range(len(stream.slabs)):
                    slot = pending
                    if s + 1 < len(stream.slabs):
                        pending = stream.fetch((s + 1) % 2, s + 1)
                    rows = slot["rows"]
                    if rows not in loops:
                        loops[rows] = TrainLoop(engine, rows, R, S, seed=noise_seed + 7919 * rank,
                                                use_graph=use_graph)
                    compute.wait_event(slot["ready"])
                    bound = loops[rows].step(slot, learning_rate, warm_up_weight)
                    slot["free"].record(compute)
                    step_bounds[s, :bound.numel()].copy_(bound)
            if rank_sharded:
                # a shuffle of this rank's own cells, cut into one share per global step
                shuffled = torch.from_numpy(local_rng.permutation(n_local)).to(engine.device)
                cuts = numpy.linspace(0, n_local, n_steps + 1).astype(numpy.int64)
            for s, i in enumerate(range(0, n_train, minibatch_size) if stream is None else []):
                rows = min(minibatch_size, n_train - i)
                if rank_sharded:
                    i, rows = int(cuts[s]), int(cuts[s + 1] - cuts[s])
                elif world > 1:   # this rank's slice of the global minibatch
                    i, hi = D.shard_bounds(i, rows, rank, world)
                    rows = hi - i
                    if rows == 0:
                        continue
                if rows not in loops:
                    # (independent reparameterisation noise per rank: the shards hold different cells)
                    loops[rows] = TrainLoop(engine, rows, R, S, seed=noise_seed + 7919 * rank,
                                            use_graph=use_graph)
                loop = loops[rows]
                loop.rows.copy_(shuffled[i:i + rows])
                bound = loop.step(data, learning_rate, warm_up_weight)
                step_bounds[s, :bound.numel()].copy_(bound)
            bounds = step_bounds.cpu().numpy()
            # a peer-memory exchange whose flag wait ran out (a rank stalled for seconds) has
            # summed stale gradients: the replicas are no longer identical and nothing computed
            # since may reach a checkpoint
            D.raise_if_exchange_failed(engine)
            if numpy.isnan(bounds[:, 0]).any():
                raise ArithmeticError("Aborting. The ELBO became indefinite during training.")
            epoch_duration = time() - epoch_time_start
            print("Epoch {} ({}): {:.3g} cells/s".format(
                epoch + 1, format_duration(epoch_duration), n_train / max(epoch_duration, 1e-9)))

            # evaluation passes with the *training* sample counts (VAE:1103-1106, 1262-1265)
            results = {"training": self._evaluate_pass(engine, data, getattr(data, "targets", None),
                                                       minibatch_size, R, S,
                                                       seed=noise_seed + 1000 + epoch,
                                                       **({"total_examples": n_train}
                                                          if rank_sharded else {}))}
            if validation_set:
                results["validation"] = self._evaluate_pass(
                    engine, valid_data, valid_data.targets,
                    min(minibatch_size, valid_data.shape[0]), R, S, seed=noise_seed + 2000 + epoch)
            for kind, result in results.items():
                if numpy.isnan(result["lower_bound"]):
                    raise ArithmeticError("Aborting. The ELBO for the {} set became indefinite."
                                          .format(kind))
                for key in learning_curves[kind]:
                    learning_curves[kind][key].append(result[key])
                subset = training_set if kind == "training" else validation_set
                scalars = self._summary_scalars(
                    result, subset, with_centroids=(kind == "validation" or not validation_set))
                writer = training_writer if kind == "training" else validation_writer
                if writer is not None:
                    writer.add_scalars(scalars, global_step=epoch + 1)
                    writer.flush()
                if not is_main:
                    continue
                print("    {} set: {}.".format(subset.kind.capitalize(),
                                               self._result_string(result)))

            # peer-exchange mode keeps one slice of the Adam slots per rank: rebuilding the full
            # slots for the checkpoint is a collective, so every rank takes part
            engine_state = engine.state_dict() \
                if (is_main or getattr(engine, "_peer", None) is not None) else None
            if not is_main:
                continue
            # early stopping (VAE:1385-1441)
            if validation_set and not self.stopped_early:
                lower_bound_valid = results["validation"]["lower_bound"]
                if lower_bound_valid < lower_bound_valid_early_stopping:
                    if epochs_with_no_improvement == 0:
                        lower_bound_valid_early_stopping = lower_bound_valid
                        current = get_checkpoint_state(log_directory)
                        if current:
                            copy_model_directory(current, early_stopping_log_directory)
                    epochs_with_no_improvement += 1
                else:
                    epochs_with_no_improvement = 0
                    lower_bound_valid_early_stopping = lower_bound_valid
                    if os.path.exists(early_stopping_log_directory):
                        shutil.rmtree(early_stopping_log_directory)
                if epochs_with_no_improvement >= self.early_stopping_rounds:
                    print("    Early stopping in effect.")
                    self.stopped_early = True
                    epochs_with_no_improvement = numpy.nan

            # checkpoint: model.ckpt-<epoch> (VAE:1446-1450) and best model (VAE:1456-1474)
            write_checkpoint(log_directory, epoch + 1,
                             {"engine": engine_state, "model": self.name,
                              "epoch": epoch + 1})
            if validation_set and results["validation"]["lower_bound"] > lower_bound_valid_maximum:
                lower_bound_valid_maximum = results["validation"]["lower_bound"]
                current = get_checkpoint_state(log_directory)
                if current:
                    copy_model_directory(current, best_model_log_directory)
                remove_old_checkpoints(best_model_log_directory)

            analyser = kwargs.get("intermediate_analyser")
            if analyser:
                last = results["validation" if validation_set else "training"]
                analyser(epoch=epoch, learning_curves=learning_curves, epoch_start=epoch_start,
                         model_type=self.type, latent_values=last[self._latent_key],
                         data_set=validation_set or training_set,
                         centroids=self._centroids(last), model_name=self.name, run_id=run_id,
                         analyses_directory=kwargs.get("analyses_directory",
                                                       defaults["analyses"]["directory"]))

        if not is_main:
            return 0
        training_writer.close()
        if validation_writer:
            validation_writer.close()
        metadata_log["training duration"] = format_duration(time() - training_time_start)
        metadata_log["last epoch duration"] = format_duration(epoch_duration)
        if temporary_log_directory:
            if os.path.exists(permanent_log_directory):
                shutil.rmtree(permanent_log_directory)
            shutil.move(log_directory, permanent_log_directory)
        with open(os.path.join(permanent_log_directory, "metadata_log-{}-{}.log".format(
                *metadata_log["epochs trained"])), "w") as handle:
            for key, value in metadata_log.items():
                handle.write("{}: {}\n".format(key, value))
        return 0

    # --- hooks specialised by the GMVAE -------------------------------------------------------
    _model_type = "VAE"
    _loss_keys = ("lower_bound", "reconstruction_error", "kl_divergence")
    _latent_key = "q_z_mean"

    def _result_string(self, result):
        return "ELBO: {:.5g}, ENRE: {:.5g}, KL: {:.5g}".format(
            result["lower_bound"], result["reconstruction_error"], result["kl_divergence"])

    def _summary_scalars(self, result, data_set, with_centroids):
        scalars = {"losses/" + key: result[key] for key in self._loss_keys}
        for j, value in enumerate(result["kl_divergence_neurons"]):
            scalars["kl_divergence_neurons/{}".format(j)] = value
        if with_centroids:   # VAE:1184, 1334
            scalars.update(self._centroid_scalars(result))
        return scalars

    def _centroid_scalars(self, result):
        """prior/cluster_0 tags of the single standard-normal prior (VAE:1184-1214; the
        'variance' is the prior's stddev, quirk Q3 -- both are 1)."""
        scalars = {"prior/cluster_0/probability": 1.0}
        for l in range(self.latent_size):
            scalars["prior/cluster_0/mean/dimension_{}".format(l)] = 0.0
            scalars["prior/cluster_0/variance/dimension_{}".format(l)] = 1.0
        return scalars

    def _centroids(self, result):
        L = self.latent_size
        return {"prior": {"probabilities": numpy.ones(1), "means": numpy.zeros((1, L)),
                          "covariance_matrices": numpy.eye(L)[None]},
                "posterior": None}

    # ------------------------------------------------------------------------------------------
    # evaluate
    # ------------------------------------------------------------------------------------------
    def _load_for_inference(self, run_id, use_early_stopping_model, use_best_model, what):
        if run_id is None:
            run_id = defaults["models"]["run_id"]
        if run_id:
            run_id = check_run_id(run_id)
        if use_early_stopping_model and use_best_model:
            raise ValueError("Early-stopping model and best model cannot be evaluated at the "
                             "same time.")
        directory = self.log_directory(run_id=run_id, early_stopping=use_early_stopping_model,
                                       best_model=use_best_model)
        engine = self._get_engine()
        epoch = self._restore(engine, directory)
        if epoch is None:
            raise Exception("Cannot {} model when it has not been trained.".format(what))
        return engine, directory, epoch

    def _reconstruction_collector(self, engine, n, minibatch_size, R, S, deterministic, subset,
                                  dtype):
        """The per-minibatch consumer of an evaluation pass that assembles the reconstruction
        (VAE:1939-2052): ``p_x_mean`` streams to pinned host memory behind the next minibatch
        (hotloop.ReconstructionSink; ``dtype`` None = no reconstruction wanted); the two deviation
        matrices are computed only for the minibatches that hold rows of ``subset``."""
        import torch
        from .hotloop import ReconstructionSink
        G = self.feature_size
        p_x_stddev = scipy.sparse.lil_matrix((n, G), dtype=numpy.float32)
        stddev_of_p_x_mean = scipy.sparse.lil_matrix((n, G), dtype=numpy.float32)
        sink = ReconstructionSink(n, G, minibatch_size, engine.device, dtype) if dtype else None

        def collect(plan, start, rows):
            if sink is None:
                return
            wanted = [i for i in subset if start <= i < start + rows]
            out = sink.mean_buffer()
            _, stddev, stddev_of_mean = engine.moments(
                plan, R, S, deterministic, mean_out=out[:plan.B], want_stddev=bool(wanted))
            sink.push(start, rows)
            if wanted:
                local = torch.tensor([i - start for i in wanted], device=engine.device)
                p_x_stddev[wanted] = stddev[local].cpu().numpy()
                stddev_of_p_x_mean[wanted] = stddev_of_mean[local].cpu().numpy()

        return collect, sink, p_x_stddev, stddev_of_p_x_mean

    def evaluate(self, evaluation_set, minibatch_size=None, run_id=None,
                 use_early_stopping_model=False, use_best_model=False, **kwargs):
        """Evaluate a trained model (VAE:1781-2217).  Returns, in ``output_versions`` order, the
        transformed DataSet, the reconstructed DataSet (``p_x_mean`` with total / explained
        standard deviations for the evaluation subset) and ``{"z": latent DataSet}``."""
        import torch
        from . import kernels as K
        if minibatch_size is None:
            minibatch_size = defaults["models"]["minibatch_size"]
        output_versions = kwargs.get("output_versions", "all")
        if output_versions == "all":
            output_versions = ["transformed", "reconstructed", "latent"]
        elif not isinstance(output_versions, list):
            output_versions = [output_versions]
        for version in output_versions:
            if version not in ("transformed", "reconstructed", "latent"):
                raise ValueError("`output_versions` can only be `all`, `transformed`, "
                                 "`reconstructed`, and/or `latent`.")
        subset = kwargs.get("evaluation_subset_indices") or set()
        log_results = kwargs.get("log_results", True)
        deterministic = bool(kwargs.get("use_deterministic_z", False))
        engine, directory, epoch = self._load_for_inference(
            run_id, use_early_stopping_model, use_best_model, "evaluate")
        R, S = self._samples("evaluation")
        if self.type == "VAE":
            minibatch_size = int(numpy.ceil(minibatch_size / (R * S)))
        x_eval, t_eval = self._inputs(evaluation_set, self.reconstruction_distribution_name)
        n, G = evaluation_set.number_of_examples, self.feature_size
        minibatch_size = min(minibatch_size, n)
        x_csr = scipy.sparse.csr_matrix(x_eval, dtype=numpy.float32)
        t_csr = x_csr if t_eval is x_eval else scipy.sparse.csr_matrix(t_eval, dtype=numpy.float32)

        want_reconstruction = "reconstructed" in output_versions
        subset = sorted(int(i) for i in subset)
        collect, sink, p_x_stddev, stddev_of_p_x_mean = self._reconstruction_collector(
            engine, n, minibatch_size, R, S, deterministic, subset,
            kwargs.get("reconstruction_dtype", "float32") if want_reconstruction else None)

        evaluating_time_start = time()
        from .hotloop import ResidentCSR
        x_data = self._attach_features(ResidentCSR(x_csr, engine.device), evaluation_set)
        result = self._evaluate_pass(engine, x_data, x_data if t_csr is x_csr else t_csr,
                                     minibatch_size, R, S, deterministic=deterministic,
                                     seed=kwargs.get("noise_seed", 7), on_batch=collect)
        p_x_mean = sink.finish() if sink is not None else None
        if numpy.isnan(result["lower_bound"]):
            raise ArithmeticError("Aborting. The ELBO for the evaluation set became indefinite.")
        if log_results:
            writer = SummaryWriter(os.path.join(directory, "evaluation"))
            writer.add_scalars(self._summary_scalars(result, evaluation_set, True),
                               global_step=epoch)
            writer.close()
        print("    {} set ({}): {}.".format(
            evaluation_set.kind.capitalize(), format_duration(time() - evaluating_time_start),
            self._result_string(result)))
        self.last_evaluation = result

        outputs = []
        common = dict(title=evaluation_set.title, labels=evaluation_set.labels,
                      example_names=evaluation_set.example_names,
                      batch_indices=evaluation_set.batch_indices,
                      batch_names=evaluation_set.batch_names, kind=evaluation_set.kind)
        for version in output_versions:
            if version == "transformed":
                outputs.append(evaluation_set)
            elif version == "reconstructed":
                outputs.append(DataSet(evaluation_set.name, values=p_x_mean,
                                       total_standard_deviations=p_x_stddev,
                                       explained_standard_deviations=stddev_of_p_x_mean,
                                       feature_names=evaluation_set.feature_names,
                                       version="reconstructed", **common))
            else:
                outputs.append(self._latent_sets(evaluation_set, result, common))
        return outputs[0] if len(outputs) == 1 else outputs

    def _latent_sets(self, evaluation_set, result, common):
        names = numpy.array(["latent variable {}".format(i + 1) for i in range(self.latent_size)])
        return {"z": DataSet(evaluation_set.name, values=result["q_z_mean"], feature_names=names,
                             version="z", **common)}

    # ------------------------------------------------------------------------------------------
    # sample
    # ------------------------------------------------------------------------------------------
    def sample(self, sample_size=None, minibatch_size=None, run_id=None,
               use_early_stopping_model=False, use_best_model=False, **kwargs):
        """Draw z ~ p(z) and decode to E[x|z] (VAE:1601-1779): returns
        ``(sample DataSet, {"z": latent DataSet})``."""
        import torch
        from . import kernels as K
        if sample_size is None:
            sample_size = defaults["models"]["sample_size"]
        if minibatch_size is None:
            minibatch_size = defaults["models"]["minibatch_size"]
        if self.batch_correction or self.use_count_sum_as_parameter or self.use_count_sum_as_feature:
            raise NotImplementedError("Sampling with batch correction or count sums is not "
                                      "implemented (as in the reference, VAE:1639-1649).")
        engine, _, _ = self._load_for_inference(run_id, use_early_stopping_model, use_best_model,
                                                "sample from")
        L, G = self.latent_size, self.feature_size
        z = numpy.empty((sample_size, L), numpy.float32)
        x = numpy.empty((sample_size, G), numpy.float32)
        seed = kwargs.get("noise_seed", 11)
        for b, i in enumerate(range(0, sample_size, minibatch_size)):
            rows = min(minibatch_size, sample_size - i)
            plan = engine._plan(rows, 1)
            K.fill_normal(plan.eps, seed, b)           # p(z) = N(0, 1)
            plan.Z[:rows, :L].copy_(plan.eps[:rows])
            engine.decode(plan, rows)
            mean, _, _ = engine.moments(plan, 1, 1)
            z[i:i + rows] = plan.eps[:rows].cpu().numpy()
            x[i:i + rows] = mean[:rows].cpu().numpy()
        names = numpy.array(["example {}".format(i + 1) for i in range(sample_size)])
        sample_set = DataSet("sample", values=x, example_names=names, kind="sample",
                             version="original")
        latent = {"z": DataSet("sample", values=z, example_names=names, kind="sample",
                               feature_names=numpy.array(
                                   ["latent variable {}".format(i + 1) for i in range(L)]),
                               version="z")}
        return sample_set, latent
