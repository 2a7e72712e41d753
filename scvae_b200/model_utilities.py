"""Host-side helpers of the model shells: naming, run ids, checkpoint directories, event
files, early stopping.  Mirrors the behaviour (not the code) of the reference's
``scvae/models/utilities.py`` and ``scvae/utilities.py`` so that the CLI / analyses side of the
drop-in boundary sees the same directory layout, file names and TensorBoard tags
(SURVEY Appendix C).  The TensorFlow pieces are replaced as follows:

  * ``tf.train.Saver``              -> ``torch.save`` of the engine state as
                                       ``model.ckpt-<epoch>.pt`` + the usual ``checkpoint`` text
                                       file (``model_checkpoint_path: "model.ckpt-<epoch>"``);
  * ``tf.summary.FileWriter``       -> ``torch.utils.tensorboard.SummaryWriter`` (simple_value
                                       scalars: readable by ``tf.train.summary_iterator`` too);
  * ``tf.train.summary_iterator``   -> tensorboard's ``EventFileLoader``.
"""

import os
import random
import re
import shutil
import time
from collections import namedtuple
from datetime import datetime, timezone
from string import ascii_uppercase

import numpy

from .defaults import defaults

ScalarEvent = namedtuple("ScalarEvent", ["wall_time", "step", "value"])
CheckpointState = namedtuple("CheckpointState", ["model_checkpoint_path"])
CHECKPOINT_SUFFIX = ".pt"

_TO_UNDERSCORE = re.compile(r"[ \-/]")
_TO_NOTHING = re.compile(r"[(),$<>:\"/\\|?*]")


def normalise_string(s):
    """Lower-case; spaces, dashes and slashes become underscores; punctuation is dropped
    (reference: scvae/utilities.py:63-77; pinned by tests/golden/normalise_string.npz)."""
    return _TO_NOTHING.sub("", _TO_UNDERSCORE.sub("_", s.lower()))


def format_duration(seconds):
    if seconds < 1e-3:
        return "<1 ms"
    if seconds < 1:
        return "{:.0f} ms".format(1e3 * seconds)
    if seconds < 60:
        return "{:.3g} s".format(seconds)
    total = int(round(seconds))
    hours, rest = divmod(total, 3600)
    minutes, secs = divmod(rest, 60)
    if hours:
        return "{:d}h {:d}m {:d}s".format(hours, minutes, secs)
    return "{:d}m {:d}s".format(minutes, secs)


def format_time(t):
    return time.strftime("%Y-%m-%d %H:%M:%S %Z", time.localtime(t))


# ---------------------------------------------------------------------------------------------
# argument parsing / validation
# ---------------------------------------------------------------------------------------------
def _as_sample_count(number):
    if isinstance(number, bool) or not isinstance(number, (int, float)):
        raise TypeError("Number of samples should be an integer.")
    if isinstance(number, float):
        if not number.is_integer():
            raise TypeError("Number of samples should be an integer.")
        number = int(number)
    return number


def parse_numbers_of_samples(proposed):
    """int | [n] | [n_train, n_eval] | {"training": .., "evaluation": ..} -> dict."""
    scenarios = ("training", "evaluation")
    if isinstance(proposed, (int, float)):
        proposed = [proposed]
    if isinstance(proposed, (list, tuple)):
        numbers = [_as_sample_count(n) for n in proposed]
        if len(numbers) == 1:
            numbers = numbers * 2
        if len(numbers) != 2:
            raise ValueError("List of number of samples can only contain one or two numbers.")
        return dict(zip(scenarios, numbers))
    if isinstance(proposed, dict):
        try:
            return {s: _as_sample_count(proposed[s]) for s in scenarios}
        except (KeyError, TypeError):
            raise ValueError("A dictionary of numbers of samples must contain the keys "
                             "`training` and `evaluation` with integer values.")
    raise TypeError("Expected an `int`, `list`, or `dict`; got `{}`.".format(type(proposed)))


_VERSION_ALIASES = {
    "end_of_training": ("eot", "end", "finish", "finished"),
    "best_model": ("bm", "best", "optimal_parameters", "op", "optimal"),
    "early_stopping": ("es", "early", "stop", "stopped"),
}


def parse_model_versions(proposed_versions):
    if not isinstance(proposed_versions, list):
        proposed_versions = [proposed_versions]
    if proposed_versions == ["all"]:
        return list(_VERSION_ALIASES)
    parsed = []
    for proposed in proposed_versions:
        key = normalise_string(proposed)
        match = [v for v, aliases in _VERSION_ALIASES.items() if key == v or key in aliases]
        if not match:
            raise ValueError("`{}` is not a model version.".format(proposed))
        parsed.append(match[0])
    return parsed


def validate_model_parameters(reconstruction_distribution=None,
                              number_of_reconstruction_classes=None, model_type=None,
                              latent_distribution=None, parameterise_latent_posterior=None):
    if reconstruction_distribution and number_of_reconstruction_classes:
        if number_of_reconstruction_classes > 0:
            offenders = []
            if reconstruction_distribution == "bernoulli":
                offenders.append("the Bernoulli distribution")
            if "zero-inflated" in reconstruction_distribution:
                offenders.append("zero-inflated distributions")
            if "constrained" in reconstruction_distribution:
                offenders.append("constrained distributions")
            if offenders:
                text = " or ".join(offenders)
                raise ValueError(text[0].upper() + text[1:] + " cannot be piecewise categorical.")
    if model_type and latent_distribution and parameterise_latent_posterior:
        if "VAE" in model_type and not (model_type == "VAE"
                                        and latent_distribution == "gaussian mixture"):
            raise ValueError("Cannot parameterise latent posterior parameters for {} or {} "
                             "distribution.".format(model_type, latent_distribution))


def check_run_id(run_id):
    if run_id is None:
        raise TypeError("The run ID has not been set.")
    run_id = str(run_id)
    if not re.fullmatch(r"\w+", run_id):
        raise ValueError("`run_id` can only contain letters, numbers, and underscores ('_').")
    return run_id


def generate_unique_run_id_for_model(model, timestamp=None):
    directory = model.log_directory()
    taken = set()
    if os.path.isdir(directory):
        taken = {d[4:] for d in os.listdir(directory) if d.startswith("run_")}
    stamp = datetime.fromtimestamp(timestamp if timestamp is not None else time.time(),
                                   tz=timezone.utc).strftime("%Y%m%dT%H%M%SZ")
    while True:
        run_id = stamp + "_" + "".join(random.choices(ascii_uppercase, k=2))
        if run_id not in taken:
            return run_id


# ---------------------------------------------------------------------------------------------
# checkpoint directories (TF Saver on-disk contract, without TF)
# ---------------------------------------------------------------------------------------------
def get_checkpoint_state(directory):
    """Stand-in for ``tf.train.get_checkpoint_state``: parses the ``checkpoint`` text file."""
    path = os.path.join(directory, "checkpoint")
    if not os.path.isfile(path):
        return None
    with open(path) as handle:
        first = handle.readline()
    match = re.match(r'model_checkpoint_path:\s*"(.*)"', first.strip())
    if not match:
        return None
    prefix = match.group(1)
    if not os.path.isabs(prefix):
        prefix = os.path.join(directory, prefix)
    if not os.path.exists(prefix + CHECKPOINT_SUFFIX):
        return None
    return CheckpointState(model_checkpoint_path=prefix)


def checkpoint_epoch(state):
    return int(os.path.basename(state.model_checkpoint_path).split("-")[-1])


def write_checkpoint(directory, epoch, payload, keep=1):
    """``saver.save(session, <dir>/model.ckpt, global_step=epoch)`` with max_to_keep=1."""
    import torch
    os.makedirs(directory, exist_ok=True)
    name = "model.ckpt-{}".format(epoch)
    torch.save(payload, os.path.join(directory, name + CHECKPOINT_SUFFIX))
    with open(os.path.join(directory, "checkpoint"), "w") as handle:
        handle.write('model_checkpoint_path: "{0}"\nall_model_checkpoint_paths: "{0}"\n'
                     .format(name))
    if keep == 1:
        for f in os.listdir(directory):
            if f.startswith("model.ckpt-") and not f.startswith(name + "."):
                os.remove(os.path.join(directory, f))
    return os.path.join(directory, name)


def read_checkpoint(state):
    import torch
    return torch.load(state.model_checkpoint_path + CHECKPOINT_SUFFIX, map_location="cpu",
                      weights_only=True)      # plain tensors / str / int only: nothing executable


def copy_model_directory(checkpoint_state, destination):
    """Copy the current checkpoint, the top-level event files and the training/validation
    event directories (what ``best/`` and ``early_stopping/`` hold in the reference)."""
    source, prefix = os.path.split(checkpoint_state.model_checkpoint_path)
    os.makedirs(destination, exist_ok=True)
    with open(os.path.join(destination, "checkpoint"), "w") as handle:
        handle.write('model_checkpoint_path: "{0}"\nall_model_checkpoint_paths: "{0}"\n'
                     .format(prefix))
    for entry in os.listdir(source):
        path = os.path.join(source, entry)
        if os.path.isfile(path) and (entry.startswith(prefix) or "events" in entry):
            shutil.copy(path, destination)
        elif os.path.isdir(path) and entry in ("training", "validation"):
            target = os.path.join(destination, entry)
            os.makedirs(target, exist_ok=True)
            for sub in os.listdir(path):
                shutil.copy(os.path.join(path, sub), target)


def remove_old_checkpoints(directory):
    state = get_checkpoint_state(directory)
    if not state:
        return
    current = os.path.basename(state.model_checkpoint_path)
    for entry in os.listdir(directory):
        path = os.path.join(directory, entry)
        if os.path.isfile(path) and "model" in entry and not entry.startswith(current + "."):
            os.remove(path)


def clear_log_directory(log_directory):
    """Remove a model's logs but keep sibling ``run_*`` directories."""
    if not os.path.exists(log_directory):
        return
    keep_parent = False
    for entry in os.listdir(log_directory):
        path = os.path.join(log_directory, entry)
        if os.path.isdir(path):
            if entry.startswith("run_"):
                keep_parent = True
            else:
                shutil.rmtree(path)
        else:
            os.remove(path)
    if not keep_parent:
        shutil.rmtree(log_directory)


# ---------------------------------------------------------------------------------------------
# event files
# ---------------------------------------------------------------------------------------------
class SummaryWriter:
    """Scalar-only event writer with the reference's ``add_summary(..., global_step)`` rhythm."""

    def __init__(self, directory):
        from torch.utils.tensorboard import SummaryWriter as _Writer
        os.makedirs(directory, exist_ok=True)
        self._writer = _Writer(log_dir=directory)

    def add_scalars(self, scalars, global_step):
        for tag, value in scalars.items():
            self._writer.add_scalar(tag, float(value), global_step=global_step)

    def flush(self):
        self._writer.flush()

    def close(self):
        self._writer.close()


def _read_scalars(log_directory, data_set_kinds, tag_searches):
    from tensorboard.backend.event_processing.event_file_loader import EventFileLoader
    from tensorboard.util import tensor_util
    if not isinstance(data_set_kinds, list):
        data_set_kinds = [data_set_kinds]
    if not isinstance(tag_searches, list):
        tag_searches = [tag_searches]
    if not os.path.exists(log_directory):
        return None
    result = {}
    for kind in data_set_kinds:
        directory = os.path.join(log_directory, kind)
        if not os.path.exists(directory):
            result[kind] = None
            continue
        found = {}
        for filename in sorted(os.listdir(directory)):
            if not filename.startswith("event"):
                continue
            for event in EventFileLoader(os.path.join(directory, filename)).Load():
                if not event.HasField("summary"):
                    continue
                for value in event.summary.value:
                    if not any(search in value.tag for search in tag_searches):
                        continue
                    if value.HasField("tensor"):
                        number = float(tensor_util.make_ndarray(value.tensor).reshape(-1)[0])
                    else:
                        number = value.simple_value
                    found.setdefault(value.tag, []).append(
                        ScalarEvent(event.wall_time, event.step, number))
        result[kind] = found
    return result


def _curve(scalars):
    if not scalars:
        return None
    curve = numpy.empty(max(len(scalars), max(s.step for s in scalars)))
    curve[:] = numpy.nan
    if len(scalars) == 1:
        curve = numpy.array([scalars[0].value])
    else:
        for s in scalars:
            curve[s.step - 1] = s.value
    return curve


def _loss_names(model):
    if model.type == "GMVAE":
        return ["lower_bound", "reconstruction_error", "kl_divergence_z", "kl_divergence_y"]
    if model.type == "VAE":
        return ["lower_bound", "reconstruction_error", "kl_divergence"]
    return ["log_likelihood"]


def load_learning_curves(model, data_set_kinds="all", run_id=None, early_stopping=False,
                         best_model=False, log_directory=None):
    if data_set_kinds == "all":
        data_set_kinds = ["training", "validation", "evaluation"]
    elif not isinstance(data_set_kinds, list):
        data_set_kinds = [data_set_kinds]
    if not log_directory:
        log_directory = model.log_directory(run_id=run_id, early_stopping=early_stopping,
                                            best_model=best_model)
    losses = _loss_names(model)
    scalar_sets = _read_scalars(log_directory, data_set_kinds, ["losses/" + l for l in losses])
    curves = {}
    for kind in data_set_kinds:
        per_kind = (scalar_sets or {}).get(kind) or {}
        curves[kind] = {l: _curve(per_kind.get("losses/" + l)) for l in losses}
    return curves[data_set_kinds[0]] if len(data_set_kinds) == 1 else curves


def load_number_of_epochs_trained(model, run_id=None, early_stopping=False, best_model=False):
    tag = "losses/" + ("lower_bound" if "VAE" in model.type else "log_likelihood")
    directory = model.log_directory(run_id=run_id, early_stopping=early_stopping,
                                    best_model=best_model)
    sets = _read_scalars(directory, "training", tag)
    scalars = ((sets or {}).get("training") or {}).get(tag)
    return max(s.step for s in scalars) if scalars else None


def load_accuracies(model, data_set_kinds="all", superset=False, run_id=None,
                    early_stopping=False, best_model=False):
    if data_set_kinds == "all":
        data_set_kinds = ["training", "validation", "evaluation"]
    elif not isinstance(data_set_kinds, list):
        data_set_kinds = [data_set_kinds]
    tag = "superset_accuracy" if superset else "accuracy"
    directory = model.log_directory(run_id=run_id, early_stopping=early_stopping,
                                    best_model=best_model)
    sets = _read_scalars(directory, data_set_kinds, tag)
    out, empty = {}, 0
    for kind in data_set_kinds:
        per_kind = (sets or {}).get(kind) or {}
        scalars = [s for t, ss in per_kind.items() if t == tag or t.endswith("/" + tag) for s in ss]
        out[kind] = _curve(scalars)
        empty += out[kind] is None
    if empty == len(data_set_kinds):
        return None
    return out[data_set_kinds[0]] if len(data_set_kinds) == 1 else out


def load_kl_divergences(model, data_set_kind=None, run_id=None, early_stopping=False,
                        best_model=False):
    """(epochs, latent) matrix of per-neuron KL divergences (tags kl_divergence_neurons/<i>)."""
    data_set_kind = data_set_kind or "training"
    directory = model.log_directory(run_id=run_id, early_stopping=early_stopping,
                                    best_model=best_model)
    sets = _read_scalars(directory, data_set_kind, "kl_divergence_neurons")
    per_kind = (sets or {}).get(data_set_kind) or {}
    if not per_kind:
        return None
    columns = sorted(per_kind, key=lambda t: int(t.split("/")[-1]))
    curves = [_curve(per_kind[t]) for t in columns]
    return numpy.stack(curves, axis=1)


def load_centroids(model, data_set_kinds="all", run_id=None, early_stopping=False,
                   best_model=False):
    """{kind: {"prior"|"posterior": {"probabilities", "means", "covariance_matrices"}}} with
    arrays indexed (epoch, cluster[, dim[, dim]]) from the cluster_<k> scalar tags."""
    if data_set_kinds == "all":
        data_set_kinds = ["training", "validation", "evaluation"]
    elif not isinstance(data_set_kinds, list):
        data_set_kinds = [data_set_kinds]
    directory = model.log_directory(run_id=run_id, early_stopping=early_stopping,
                                    best_model=best_model)
    sets = _read_scalars(directory, data_set_kinds, "cluster")
    K = max(getattr(model, "number_of_latent_clusters", 1), 1)
    L = model.latent_size
    result = {}
    for kind in data_set_kinds:
        per_kind = (sets or {}).get(kind) or {}
        if not per_kind:
            result[kind] = None
            continue
        n_epochs = max(len(v) for v in per_kind.values())
        kind_result = {}
        for dist in ("prior", "posterior"):
            probs = numpy.full((n_epochs, K), numpy.nan)
            means = numpy.full((n_epochs, K, L), numpy.nan)
            covs = numpy.zeros((n_epochs, K, L, L))
            seen = False
            for tag, scalars in per_kind.items():
                parts = tag.split("/")
                if parts[0] != dist:
                    continue
                seen = True
                k = int(parts[1].split("_")[-1])
                for e, s in enumerate(sorted(scalars, key=lambda s: s.step)):
                    if parts[2] == "probability":
                        probs[e, k] = s.value
                    elif parts[2] == "mean":
                        means[e, k, int(parts[3].split("_")[-1])] = s.value
                    elif parts[2] == "variance":
                        l = int(parts[3].split("_")[-1])
                        covs[e, k, l, l] = s.value
            kind_result[dist] = {"probabilities": probs, "means": means,
                                 "covariance_matrices": covs} if seen else None
        result[kind] = kind_result
    return result[data_set_kinds[0]] if len(data_set_kinds) == 1 else result


# ---------------------------------------------------------------------------------------------
# early stopping
# ---------------------------------------------------------------------------------------------
def early_stopping_status(losses, early_stopping_rounds):
    """Count consecutive epochs whose validation bound fell below the previous epoch's."""
    without_improvement, stopped = 0, False
    if losses is not None:
        for previous, current in zip(losses[:-1], losses[1:]):
            without_improvement = without_improvement + 1 if current < previous else 0
            if without_improvement >= early_stopping_rounds:
                return True, numpy.nan
    return stopped, without_improvement


def better_model_exists(model, run_id=None):
    current = load_number_of_epochs_trained(model, run_id=run_id)
    best = load_number_of_epochs_trained(model, run_id=run_id, best_model=True)
    return bool(best) and best < current


def model_stopped_early(model, run_id=None):
    stopped, _ = model.early_stopping_status(run_id=run_id)
    return stopped


def build_training_string(model_string, epoch_start, number_of_epochs, data_string):
    if epoch_start == 0:
        return "Training {} for {} epochs on {}.".format(model_string, number_of_epochs,
                                                          data_string)
    if epoch_start < number_of_epochs:
        return "Continue training {} for {} additionally epochs (up to {} epochs) on {}.".format(
            model_string, number_of_epochs - epoch_start, number_of_epochs, data_string)
    subject = model_string[0].upper() + model_string[1:]
    if epoch_start == number_of_epochs:
        return "{} has already been trained for {} epochs on {}.".format(
            subject, number_of_epochs, data_string)
    # trained beyond the requested number of epochs (MU:163-171)
    return ("{} has already been trained for more than {} epochs on {}. "
            "Loading model trained for {} epochs.").format(
                subject, number_of_epochs, data_string, epoch_start)
