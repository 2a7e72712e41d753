"""Eager stand-in for the TensorFlow-1.x / TensorFlow-Probability-0.7 calls of the reference's
model graph -- TEST INFRASTRUCTURE (build container only), NOT PRODUCT CODE.

Why: TensorFlow 1.15 / TFP 0.7 cannot be installed here, so the reference cannot
run as shipped.  Its *graph-building code*, however, is plain Python that composes TF
primitives: ``VariationalAutoencoder._setup_model_graph / _setup_loss_function /
_setup_optimiser`` (VAE:2219-2770), the GMVAE counterparts (GMVAE:2788-3470), ``dense_layer(s)`` /
``log_reduce_exp`` (MU:38-137), the distribution table (DU:30-306) and the reference's own
``ZeroInflated`` (ZI:59-199) and ``Categorised`` (CAT:44-274) classes.  This module provides the
primitives those files call -- as eager PyTorch-CPU functions (fp64) under the TF names -- so that
``oracle/make_golden.py graph`` can EXECUTE THE REFERENCE'S OWN FILES, unmodified, from
``/root/reference`` and record {weights, x, eps} -> {ELBO, q_z_mean, log p, moments, gradients,
updated weights} as golden vectors (``tests/golden/reference_graph_*.npz``).

What that pins: everything the reference's own code decides -- variable scopes / names, layer
order, which layers get batch norm / dropout, clip bounds, head order, the tiling / reshape
conventions over (R, S, B), the zero-inflated and piecewise-categorical formulas, the KL and
ELBO aggregation, free nats, the moments, gradient clipping before Adam and the ordering of the
batch-norm updates.  What it does NOT pin: the arithmetic inside TF / TFP themselves
(``fully_connected``, fused ``batch_norm``, ``AdamOptimizer``, ``Normal`` / ``Poisson`` /
``NegativeBinomial`` / ``Categorical``), which is restated here from their published closed
forms (and, for the count distributions, checked against scipy in ``tests/test_oracle.py``).

How eager evaluation fits a graph-mode program: ``tf.placeholder(name=...)`` returns the value
registered for that name *before* the model class is constructed (``feeds``), variables take
their injected values (``initial``), ``Distribution.sample`` consumes injected noise
(``noise``, in call order), dropout consumes injected masks (by variable scope), and
``AdamOptimizer.apply_gradients`` / the batch-norm update ops write the post-step values into
``updates`` instead of mutating the variables.  One model construction = one ``session.run``.
"""

from __future__ import annotations

import contextlib
import importlib.machinery
import math
import sys
import types
from collections import OrderedDict, defaultdict
from unittest import mock

import numpy
import torch


class TT(torch.Tensor):
    """Tensor with the few graph-tensor attributes the reference reads."""
    name = None

    def get_shape(self):
        return _Shape(self.shape)


class _Shape(tuple):
    @property
    def ndims(self):
        return len(self)

    def as_list(self):
        return list(self)


class _State:
    def reset(self, feeds=None, initial=None, noise=None, dropout_masks=None, seed=0):
        self.dtype = torch.float64
        self.feeds = dict(feeds or {})
        self.initial = dict(initial or {})
        self.noise = list(noise or [])
        self.dropout_masks = dict(dropout_masks or {})
        self.variables = OrderedDict()
        self.trainable = []
        self.collections = defaultdict(list)
        self.scope = []
        self.updates = OrderedDict()
        self.gradients = OrderedDict()
        self.generator = torch.Generator().manual_seed(seed)
        self.sample_calls = []
        self.batch_statistics = defaultdict(list)
        self.dropout_visits = {}


STATE = _State()
STATE.reset()


def _t(value, dtype=None):
    """Anything -> TT of the working float dtype (ints / bools keep their type)."""
    if isinstance(value, torch.Tensor):
        out = value
    else:
        out = torch.as_tensor(numpy.asarray(value))
    if dtype is not None:
        out = out.to(dtype)
    elif out.dtype in (torch.float32, torch.float16):
        out = out.to(STATE.dtype)
    return out.as_subclass(TT)


def _named(tensor, name):
    if name is not None and isinstance(tensor, torch.Tensor):
        tensor = tensor.as_subclass(TT)
        tensor.name = "/".join(STATE.scope + [name]) + ":0"
    return tensor


def _scope_path(*parts):
    return "/".join(STATE.scope + [p for p in parts if p])


# --------------------------------------------------------------------------------------------
# tensorflow
# --------------------------------------------------------------------------------------------

class _DType:
    def __init__(self, name, kind):
        self.name, self.kind = name, kind

    def __repr__(self):
        return "tf." + self.name

    def torch(self):
        return {"float": STATE.dtype, "int": torch.int64, "bool": torch.bool}[self.kind]


float32 = _DType("float32", "float")
float64 = _DType("float64", "float")
int32 = _DType("int32", "int")
int64 = _DType("int64", "int")
bool_ = _DType("bool", "bool")


def placeholder(dtype, shape=None, name=None):
    if name not in STATE.feeds:
        raise KeyError("no value fed for placeholder {!r}".format(name))
    value = STATE.feeds[name]
    if shape is not None and len(shape) == 0 and dtype.kind in ("int", "bool"):
        return int(value) if dtype.kind == "int" else bool(value)   # scalar control inputs
    return _named(_t(value, dtype.torch()), name)


def constant(value, dtype=None, shape=None, name=None):
    out = _t(value, dtype.torch() if dtype is not None else None)
    if shape is not None:
        out = out.reshape(shape)
    return _named(out, name)


def convert_to_tensor(value, dtype=None, name=None, **_):
    return _t(value, dtype.torch() if dtype is not None else None)


def identity(x, name=None):
    return _t(x)


def cast(x, dtype, name=None):
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    return _named(x.to(dtype.torch()), name or "Cast")


def expand_dims(x, axis=None, name=None, dim=None):
    return torch.unsqueeze(_t(x), axis if axis is not None else dim)


def squeeze(x, axis=None, name=None):
    x = _t(x)
    if axis is None:
        return torch.squeeze(x)
    return torch.squeeze(x, axis)


def reshape(x, shape, name=None):
    return _named(torch.reshape(_t(x), [int(s) for s in shape]), name)


def tile(x, multiples, name=None):
    return _named(_t(x).repeat(*[int(m) for m in multiples]), name)


def concat(values, axis, name=None):
    return _named(torch.cat([_t(v, STATE.dtype) for v in values], dim=axis), name)


def shape(x, name=None):
    return tuple(x.shape)


def size(x, name=None):
    return len(x) if isinstance(x, tuple) else x.numel()


def broadcast_to(x, shape):
    return _t(x).expand(*[int(s) for s in shape])


def zeros(shape, dtype=None, name=None):
    shape = [shape] if isinstance(shape, int) else list(shape)
    return torch.zeros(shape, dtype=STATE.dtype).as_subclass(TT)


def ones(shape, dtype=None, name=None):
    shape = [shape] if isinstance(shape, int) else list(shape)
    return torch.ones(shape, dtype=STATE.dtype).as_subclass(TT)


def ones_like(x, **_):
    return torch.ones_like(_t(x))


def one_hot(indices, depth, **_):
    return torch.nn.functional.one_hot(_t(indices).long(), int(depth)).to(STATE.dtype).as_subclass(TT)


def _reduce(fn):
    def op(x, axis=None, keepdims=False, name=None, **_):
        x = _t(x)
        if axis is None:
            out = fn(x)
            return _named(out.reshape([1] * x.dim()) if keepdims else out, name)
        axis = tuple(axis) if isinstance(axis, (list, tuple)) else axis
        return _named(fn(x, dim=axis, keepdim=keepdims), name)
    return op


reduce_sum = _reduce(torch.sum)
reduce_mean = _reduce(torch.mean)
reduce_max = _reduce(lambda x, **kw: torch.amax(x, **kw) if kw else torch.max(x))


def add_n(inputs, name=None):
    out = inputs[0]
    for value in inputs[1:]:
        out = out + value
    return out


def exp(x, name=None):
    return torch.exp(_t(x))


def log(x, name=None):
    return torch.log(_t(x))


def sqrt(x, name=None):
    return torch.sqrt(_t(x))


def square(x, name=None):
    x = _t(x)
    return x * x


def sigmoid(x, name=None):
    return torch.sigmoid(_t(x))


def multiply(x, y, name=None):
    return _named(_t(x) * y, name)


def where(condition, x=None, y=None, name=None):
    if isinstance(condition, bool):
        return x if condition else y
    if not isinstance(x, torch.Tensor) and not isinstance(y, torch.Tensor) \
            and condition.dim() == 0:
        return x if bool(condition) else y
    return torch.where(condition, _t(x, STATE.dtype), _t(y, STATE.dtype))


def cond(pred, true_fn=None, false_fn=None, name=None):
    return true_fn() if bool(pred) else false_fn()


def clip_by_value(x, clip_value_min, clip_value_max, name=None):
    return torch.clamp(_t(x), min=float(clip_value_min), max=float(clip_value_max))


def softmax(x, axis=-1, name=None):
    return torch.softmax(_t(x), dim=axis)


def log_softmax(x, axis=-1, name=None):
    return torch.log_softmax(_t(x), dim=axis)


def unstack(x, num=None, axis=0, name=None):
    return [piece.as_subclass(TT) for piece in torch.unbind(x, dim=axis)]


def random_normal(shape, mean=0.0, stddev=1.0, **_):
    return (torch.randn(list(shape), generator=STATE.generator, dtype=STATE.dtype) * stddev
            + mean).as_subclass(TT)


def diag_part(x, name=None):
    return torch.diagonal(x)


@contextlib.contextmanager
def variable_scope(name, reuse=None, **_):
    STATE.scope.append(name)
    try:
        yield
    finally:
        STATE.scope.pop()


@contextlib.contextmanager
def name_scope(name=None, *args, **kwargs):
    yield


@contextlib.contextmanager
def control_dependencies(ops):
    yield


def get_variable(path, initial_value, trainable=True):
    """Variables are shared by full scope path (``reuse=True`` in the reference's K-fold loops)."""
    if path in STATE.variables:
        return STATE.variables[path]
    if path in STATE.initial:
        value = _t(STATE.initial[path], STATE.dtype if torch.as_tensor(
            numpy.asarray(STATE.initial[path])).is_floating_point() else None)
    else:
        value = initial_value() if callable(initial_value) else _t(initial_value)
    value = value.detach().clone().as_subclass(TT)
    if trainable and value.is_floating_point():
        value.requires_grad_(True)
    value.name = path + ":0"
    STATE.variables[path] = value
    if trainable:
        STATE.trainable.append(path)
    return value


def Variable(initial_value=None, name=None, trainable=True, **_):
    return get_variable(_scope_path(name), initial_value, trainable)


def trainable_variables():
    return [STATE.variables[name] for name in STATE.trainable]


def add_to_collection(name, value):
    STATE.collections[name].append(value)


def get_collection(name):
    return list(STATE.collections[name])


def group(*ops, **_):
    return None


def assert_positive(x, **_):
    return None


def global_variables_initializer():
    return None


class Graph:
    @contextlib.contextmanager
    def as_default(self):
        yield self


class GraphKeys:
    UPDATE_OPS = "update_ops"


class AdamOptimizer:
    """tf.train.AdamOptimizer (defaults beta1 0.9, beta2 0.999, epsilon 1e-8), first step from
    zero slots: lr_t = lr sqrt(1 - b2^t) / (1 - b1^t); var -= lr_t m / (sqrt(v) + eps)."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, **_):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon

    def compute_gradients(self, loss, var_list=None):
        variables = trainable_variables()
        grads = torch.autograd.grad(loss, variables, allow_unused=True)
        pairs = []
        for g, v in zip(grads, variables):
            if g is None:       # TF returns (None, v); clip_by_value(None) would raise there too
                g = torch.zeros_like(v)
            STATE.gradients[v.name[:-2]] = g.detach()
            pairs.append((g.detach().as_subclass(TT), v))
        return pairs

    def apply_gradients(self, grads_and_vars, global_step=None, name=None):
        slots_m = STATE.initial.get("__adam_m__", {})
        slots_v = STATE.initial.get("__adam_v__", {})
        step = int(STATE.initial.get("__adam_step__", 0)) + 1
        lr = float(self.lr)
        lr_t = lr * math.sqrt(1.0 - self.b2 ** step) / (1.0 - self.b1 ** step)
        for g, v in grads_and_vars:
            key = v.name[:-2]
            m0 = _t(slots_m[key], STATE.dtype) if key in slots_m else torch.zeros_like(v)
            v0 = _t(slots_v[key], STATE.dtype) if key in slots_v else torch.zeros_like(v)
            m = self.b1 * m0 + (1.0 - self.b1) * g
            s = self.b2 * v0 + (1.0 - self.b2) * g * g
            STATE.updates[key] = (v.detach() - lr_t * m / (torch.sqrt(s) + self.eps)).detach()
            STATE.updates["__adam_m__/" + key] = m.detach()
            STATE.updates["__adam_v__/" + key] = s.detach()
        if global_step is not None:
            STATE.updates["global_step"] = torch.as_tensor(step)
        return "train_op"


# tensorflow.contrib.layers -------------------------------------------------------------------

def fully_connected(inputs, num_outputs, activation_fn=None, scope=None, reuse=None, **_):
    """x W + b; W (in, out) Xavier-uniform, b zeros (tf.contrib.layers defaults)."""
    inputs = _t(inputs, STATE.dtype)
    fan_in = inputs.shape[-1]
    limit = math.sqrt(6.0 / (fan_in + num_outputs))
    path = _scope_path(scope)
    w = get_variable(path + "/weights", lambda: ((torch.rand(
        (fan_in, num_outputs), generator=STATE.generator, dtype=STATE.dtype) * 2 - 1)
        * limit).as_subclass(TT))
    b = get_variable(path + "/biases", lambda: zeros([num_outputs]))
    if tuple(w.shape) != (fan_in, num_outputs):
        raise ValueError("{}: weights {} for a ({}, {}) layer".format(
            path, tuple(w.shape), fan_in, num_outputs))
    out = inputs @ w + b
    return activation_fn(out) if activation_fn is not None else out


def batch_norm(inputs, decay=0.999, center=True, scale=False, epsilon=0.001,
               is_training=True, scope=None, reuse=None, **_):
    """tf.contrib.layers.batch_norm, fused path: batch mean / biased variance to normalise while
    training, moving statistics otherwise; the moving variance is updated with the
    Bessel-corrected batch variance; update ops go to UPDATE_OPS and are run (in creation
    order) before the optimiser (VAE:2760-2768)."""
    path = _scope_path(scope)
    width = inputs.shape[-1]
    beta = get_variable(path + "/beta", lambda: zeros([width])) if center else 0.0
    gamma = get_variable(path + "/gamma", lambda: ones([width])) if scale else 1.0
    moving_mean = get_variable(path + "/moving_mean", lambda: zeros([width]), trainable=False)
    moving_variance = get_variable(path + "/moving_variance", lambda: ones([width]),
                                   trainable=False)
    if bool(is_training):
        n = inputs.shape[0]
        mean = inputs.mean(dim=0)
        variance = ((inputs - mean) ** 2).mean(dim=0)
        for key, old, new in (
                (path + "/moving_mean", moving_mean, mean),
                (path + "/moving_variance", moving_variance, variance * (n / max(n - 1, 1)))):
            current = STATE.updates.get(key, old.detach())
            STATE.updates[key] = (current - (current - new.detach()) * (1.0 - decay)).detach()
        STATE.collections[GraphKeys.UPDATE_OPS].append(path)
        STATE.batch_statistics[path].append((mean.detach(), variance.detach()))
    else:
        mean, variance = moving_mean, moving_variance
    return (inputs - mean) / torch.sqrt(variance + epsilon) * gamma + beta


def dropout(inputs, keep_prob=0.5, is_training=True, **_):
    """Inverted dropout on the layer input while training; the 0/1 mask is injected by scope."""
    if not bool(is_training):
        return inputs
    site = _scope_path()
    # a scope visited again (the GMVAE builds its layers once per cluster with reuse=True) is a
    # NEW dropout op with its own mask in TensorFlow: key the later visits `site#1`, `site#2`, ...
    visit = STATE.dropout_visits.get(site, 0)
    STATE.dropout_visits[site] = visit + 1
    if visit:
        site = "{}#{}".format(site, visit)
    if site in STATE.dropout_masks:
        mask = _t(STATE.dropout_masks[site], STATE.dtype)
    else:
        mask = (torch.rand(inputs.shape, generator=STATE.generator, dtype=torch.float64)
                < keep_prob).to(STATE.dtype)
        STATE.dropout_masks[site] = mask
    return inputs * mask / keep_prob


# --------------------------------------------------------------------------------------------
# tensorflow_probability
# --------------------------------------------------------------------------------------------

class Distribution:
    """tfp Distribution base: public methods forward to the ``_``-prefixed implementations."""

    def __init__(self, dtype=None, reparameterization_type=None, validate_args=False,
                 allow_nan_stats=True, parameters=None, graph_parents=None, name=None):
        self._dtype = dtype
        self._graph_parents = list(graph_parents or [])
        self.name = name

    dtype = property(lambda self: self._dtype)
    event_shape = property(lambda self: self._event_shape())
    batch_shape = property(lambda self: self._batch_shape())

    def _event_shape(self):
        return _Shape(())

    def event_shape_tensor(self):
        return self._event_shape_tensor()

    def batch_shape_tensor(self):
        return self._batch_shape_tensor()

    def log_prob(self, value, name=None):
        return self._log_prob(_t(value, STATE.dtype))

    def prob(self, value, name=None):
        return self._prob(_t(value, STATE.dtype))

    def _prob(self, value):
        return torch.exp(self._log_prob(value))

    def mean(self, name=None):
        return self._mean()

    def variance(self, name=None):
        return self._variance()

    def stddev(self, name=None):
        return self._stddev()

    def _stddev(self):
        return torch.sqrt(self._variance())

    def entropy(self, name=None):
        return self._entropy()

    def sample(self, sample_shape=(), seed=None, name=None):
        if isinstance(sample_shape, (tuple, list)):
            sample_shape = tuple(int(s) for s in sample_shape)
        else:
            sample_shape = (int(sample_shape),)
        STATE.sample_calls.append((type(self).__name__, sample_shape))
        return self._sample(sample_shape)


class Normal(Distribution):
    def __init__(self, loc, scale, **_):
        super().__init__(dtype=float32, graph_parents=[loc, scale], name="Normal")
        self.loc, self.scale = _t(loc, STATE.dtype), _t(scale, STATE.dtype)

    def _batch_shape(self):
        return _Shape(torch.broadcast_shapes(self.loc.shape, self.scale.shape))

    def _batch_shape_tensor(self):
        return tuple(self._batch_shape())

    def _mean(self):
        return self.loc * torch.ones_like(self.scale)

    def _stddev(self):
        return self.scale * torch.ones_like(self.loc)

    def _variance(self):
        return self._stddev() ** 2

    def _log_prob(self, x):
        z = (x - self.loc) / self.scale
        return -0.5 * z * z - 0.5 * math.log(2.0 * math.pi) - torch.log(self.scale)

    def _sample(self, sample_shape):
        shape = tuple(sample_shape) + tuple(self._batch_shape())
        if STATE.noise:
            eps = _t(STATE.noise.pop(0), STATE.dtype).reshape(shape)
        else:
            eps = torch.randn(shape, generator=STATE.generator, dtype=STATE.dtype)
        return self.loc + self.scale * eps


class Poisson(Distribution):
    def __init__(self, rate=None, log_rate=None, **_):
        super().__init__(dtype=float32, graph_parents=[rate], name="Poisson")
        self.rate = _t(rate, STATE.dtype) if rate is not None else torch.exp(log_rate)
        self.log_rate = torch.log(self.rate) if log_rate is None else _t(log_rate, STATE.dtype)

    def _batch_shape(self):
        return _Shape(self.rate.shape)

    def _batch_shape_tensor(self):
        return tuple(self.rate.shape)

    def _mean(self):
        return self.rate

    def _variance(self):
        return self.rate

    def _log_prob(self, x):
        return x * self.log_rate - torch.lgamma(1.0 + x) - self.rate


class NegativeBinomial(Distribution):
    """tfp NegativeBinomial(total_count, probs): probs = success probability of the counted
    events; mean = total_count * probs / (1 - probs)."""

    def __init__(self, total_count, logits=None, probs=None, **_):
        super().__init__(dtype=float32, graph_parents=[total_count, probs], name="NB")
        self.total_count = _t(total_count, STATE.dtype)
        if logits is None:
            probs = _t(probs, STATE.dtype)
            logits = torch.log(probs) - torch.log1p(-probs)
        self.logits = logits

    def _batch_shape(self):
        return _Shape(torch.broadcast_shapes(self.total_count.shape, self.logits.shape))

    def _batch_shape_tensor(self):
        return tuple(self._batch_shape())

    def _mean(self):
        return self.total_count * torch.exp(self.logits)

    def _variance(self):
        return self._mean() / torch.sigmoid(-self.logits)

    def _log_prob(self, x):
        log_unnormalized = (self.total_count * torch.nn.functional.logsigmoid(-self.logits)
                            + x * torch.nn.functional.logsigmoid(self.logits))
        log_normalization = (-torch.lgamma(self.total_count + x) + torch.lgamma(1.0 + x)
                             + torch.lgamma(self.total_count))
        return log_unnormalized - log_normalization


class Categorical(Distribution):
    def __init__(self, logits=None, probs=None, **_):
        super().__init__(dtype=int32, graph_parents=[logits], name="Categorical")
        self.logits = _t(logits, STATE.dtype) if logits is not None else torch.log(_t(probs))
        self.probs = torch.softmax(self.logits, dim=-1)
        self.event_size = int(self.logits.shape[-1])

    def _batch_shape(self):
        return _Shape(self.logits.shape[:-1])

    def _batch_shape_tensor(self):
        return tuple(self.logits.shape[:-1])

    def _log_prob(self, k):
        log_probs = torch.log_softmax(self.logits, dim=-1)
        k = k.long()
        log_probs, k = torch.broadcast_tensors(log_probs, k.unsqueeze(-1))
        return torch.gather(log_probs, -1, k[..., :1]).squeeze(-1)

    def log_prob(self, value, name=None):
        return self._log_prob(_t(value))

    def _entropy(self):
        return -(self.probs * torch.log_softmax(self.logits, dim=-1)).sum(dim=-1)

    def _sample(self, sample_shape):
        n = int(numpy.prod(sample_shape))
        flat = self.probs.reshape(-1, self.event_size).detach()
        draws = torch.multinomial(flat, n, replacement=True, generator=STATE.generator)
        return draws.t().reshape(tuple(sample_shape) + tuple(self._batch_shape()))


def kl_divergence(a, b, **_):
    if isinstance(a, Normal) and isinstance(b, Normal):
        ratio = (a.scale * a.scale) / (b.scale * b.scale)
        return ((a.loc - b.loc) ** 2 / (2.0 * b.scale * b.scale)
                + 0.5 * (ratio - 1.0 - torch.log(ratio)))
    if isinstance(a, Categorical) and isinstance(b, Categorical):
        la, lb = torch.log_softmax(a.logits, -1), torch.log_softmax(b.logits, -1)
        return (torch.softmax(a.logits, -1) * (la - lb)).sum(dim=-1)
    raise NotImplementedError("kl_divergence({}, {})".format(type(a).__name__, type(b).__name__))


class _NotBuilt(Distribution):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(type(self).__name__ + " is outside the stand-in")


class MultivariateNormalDiag(_NotBuilt):
    pass


def fill_triangular(x, upper=False, name=None):
    """tfp.distributions.fill_triangular (TFP 0.7): reshape(concat(x[n:], reverse(x)), (n, n)),
    lower (or upper) triangle kept; [1..6] -> [[4, 0, 0], [6, 5, 0], [3, 2, 1]]."""
    x = _t(x, STATE.dtype)
    m = x.shape[-1]
    n = int(round((math.sqrt(8 * m + 1) - 1) / 2))
    assert n * (n + 1) // 2 == m, "fill_triangular: not a triangular number"
    if upper:
        full = torch.cat([x, torch.flip(x[..., n:], dims=[-1])], dim=-1).reshape(x.shape[:-1] + (n, n))
        return torch.triu(full)
    full = torch.cat([x[..., n:], torch.flip(x, dims=[-1])], dim=-1).reshape(x.shape[:-1] + (n, n))
    return torch.tril(full)


class MultivariateNormalTriL(Distribution):
    """tfd.MultivariateNormalTriL: loc (..., d), scale_tril (..., d, d); batch shape = the broadcast of
    loc.shape[:-1] and scale_tril.shape[:-2], event shape (d,)."""

    def __init__(self, loc=None, scale_tril=None, validate_args=False, allow_nan_stats=True,
                 name="MultivariateNormalTriL"):
        super().__init__(dtype=float32, graph_parents=[loc, scale_tril], name=name)
        self.loc, self.scale_tril = _t(loc, STATE.dtype), _t(scale_tril, STATE.dtype)
        self.d = self.scale_tril.shape[-1]

    def _batch_shape(self):
        return _Shape(torch.broadcast_shapes(self.loc.shape[:-1], self.scale_tril.shape[:-2]))

    def _batch_shape_tensor(self):
        return tuple(self._batch_shape())

    def _event_shape(self):
        return _Shape((self.d,))

    def _mean(self):
        return self.loc.expand(tuple(self._batch_shape()) + (self.d,))

    def covariance(self, name=None):
        cov = self.scale_tril @ self.scale_tril.transpose(-1, -2)
        return cov.expand(tuple(self._batch_shape()) + (self.d, self.d))

    def _variance(self):
        return torch.diagonal(self.covariance(), dim1=-2, dim2=-1)

    def _log_prob(self, x):
        shape = torch.broadcast_shapes(x.shape[:-1], tuple(self._batch_shape()))
        r = (x - self.loc).expand(shape + (self.d,)).unsqueeze(-1)
        tril = self.scale_tril.expand(shape + (self.d, self.d))
        y = torch.linalg.solve_triangular(tril, r, upper=False).squeeze(-1)
        log_det = torch.log(torch.abs(torch.diagonal(tril, dim1=-2, dim2=-1))).sum(dim=-1)
        return -0.5 * (y * y).sum(dim=-1) - log_det - 0.5 * self.d * math.log(2.0 * math.pi)

    def _sample(self, sample_shape):
        shape = tuple(sample_shape) + tuple(self._batch_shape()) + (self.d,)
        if STATE.noise:
            eps = _t(STATE.noise.pop(0), STATE.dtype).reshape(shape)
        else:
            eps = torch.randn(shape, generator=STATE.generator, dtype=STATE.dtype)
        return self.loc + (self.scale_tril @ eps.unsqueeze(-1)).squeeze(-1)


class MultivariateNormalFullCovariance(_NotBuilt):
    pass


# --------------------------------------------------------------------------------------------
# Module assembly
# --------------------------------------------------------------------------------------------

class _Module(types.ModuleType):
    """Module with explicit attributes; anything else is an inert mock (import-time names of
    the distributions this stand-in does not cover)."""

    def __getattr__(self, key):
        if key.startswith("__"):
            raise AttributeError(key)
        value = mock.MagicMock(name=self.__name__ + "." + key)
        setattr(self, key, value)
        return value


class _Finder:
    ROOTS = ("tensorflow", "tensorflow_probability")

    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in self.ROOTS and name not in sys.modules:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        return _Module(spec.name)

    def exec_module(self, module):
        pass


def _module(name, **attributes):
    module = _Module(name)
    module.__path__ = []
    for key, value in attributes.items():
        setattr(module, key, value)
    sys.modules[name] = module
    parent, _, leaf = name.rpartition(".")
    if parent in sys.modules:       # ``from parent import leaf`` must find the real submodule
        setattr(sys.modules[parent], leaf, module)
    return module


def install():
    """Put the stand-in modules into ``sys.modules`` under the TensorFlow / TFP names."""
    this = sys.modules[__name__]
    ops_names = [
        "placeholder", "constant", "convert_to_tensor", "identity", "cast", "expand_dims",
        "squeeze", "reshape", "tile", "concat", "shape", "size", "broadcast_to", "zeros", "ones",
        "ones_like", "one_hot", "reduce_sum", "reduce_mean", "reduce_max", "add_n", "exp", "log",
        "sqrt", "square", "sigmoid", "multiply", "where", "cond", "clip_by_value",
        "random_normal", "diag_part", "variable_scope", "name_scope", "control_dependencies",
        "Variable", "trainable_variables", "add_to_collection", "get_collection", "group",
        "assert_positive", "global_variables_initializer", "Graph", "GraphKeys", "unstack",
        "softmax", "log_softmax",
    ]
    everything = {name: getattr(this, name) for name in ops_names}
    tf = _module("tensorflow", float32=float32, float64=float64, int32=int32, int64=int64,
                 bool=bool_, **everything)
    tf.nn = _module("tensorflow.nn", relu=lambda x, name=None: torch.relu(_t(x)),
                    softplus=lambda x, name=None: torch.nn.functional.softplus(_t(x), threshold=1e4),
                    softmax=softmax, log_softmax=log_softmax, sigmoid=sigmoid)
    tf.train = _module("tensorflow.train", AdamOptimizer=AdamOptimizer,
                       Saver=lambda *a, **kw: None)
    tf.summary = _module("tensorflow.summary", scalar=lambda *a, **kw: None,
                         histogram=lambda *a, **kw: None, merge=lambda *a, **kw: None)
    tf.contrib = _module("tensorflow.contrib")
    tf.contrib.layers = _module("tensorflow.contrib.layers", fully_connected=fully_connected,
                                batch_norm=batch_norm, dropout=dropout)
    python = _module("tensorflow.python")
    python.framework = _module("tensorflow.python.framework")
    python.ops = _module("tensorflow.python.ops")
    _module("tensorflow.python.framework.ops", convert_to_tensor=convert_to_tensor,
            name_scope=name_scope, control_dependencies=control_dependencies)
    _module("tensorflow.python.framework.dtypes", int32=int32, float32=float32)
    _module("tensorflow.python.framework.tensor_util", constant_value=lambda v: v)
    _module("tensorflow.python.ops.array_ops", shape=shape, size=size, unstack=unstack)
    _module("tensorflow.python.ops.check_ops", assert_positive=assert_positive)
    _module("tensorflow.python.ops.clip_ops", clip_by_value=clip_by_value)
    _module("tensorflow.python.ops.math_ops", log=log, exp=exp, square=square, cast=cast,
            add_n=add_n)
    _module("tensorflow.python.ops.nn_ops", softmax=softmax, log_softmax=log_softmax)
    distributions = _module(
        "tensorflow_probability.distributions", Normal=Normal, Poisson=Poisson,
        NegativeBinomial=NegativeBinomial, Categorical=Categorical, kl_divergence=kl_divergence,
        Distribution=Distribution, MultivariateNormalDiag=MultivariateNormalDiag,
        MultivariateNormalTriL=MultivariateNormalTriL,
        MultivariateNormalFullCovariance=MultivariateNormalFullCovariance,
        fill_triangular=fill_triangular)
    tfp = _module("tensorflow_probability", distributions=distributions)
    tfp.python = _module("tensorflow_probability.python")
    _module("tensorflow_probability.python.distributions")
    _module("tensorflow_probability.python.internal")
    _module("tensorflow_probability.python.distributions.distribution",
            Distribution=Distribution)
    _module("tensorflow_probability.python.distributions.categorical", Categorical=Categorical)
    _module("tensorflow_probability.python.internal.reparameterization",
            NOT_REPARAMETERIZED="NOT_REPARAMETERIZED",
            FULLY_REPARAMETERIZED="FULLY_REPARAMETERIZED")
    sys.meta_path.insert(0, _Finder())
    return tf, tfp
