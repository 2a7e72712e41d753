"""Step engines: the per-minibatch hot path of the reference as a fixed sequence of
``libscvae_b200`` kernel launches over statically allocated HBM buffers.

``VAEEngine`` replaces the TensorFlow graph built by
``VariationalAutoencoder._setup_model_graph/_setup_loss_function/_setup_optimiser``
(scvae/models/variational_autoencoder.py:2219-2770) and its execution through
``session.run([optimiser, lower_bound])`` (:1026-1029).  Backward is written out by hand (the
analytic gradients of SURVEY A.8) -- there is no autograd tape and no torch math on the path;
torch only owns the device allocations and the stream.

Data layout in HBM (all fp32, row-major):
  * activations are "augmented": logical width K stored as Kp = round4(K+1) columns with
    column K == 1 and the rest 0, so biases are column K of the weight matrices;
  * weights are stored (out, in_p) -- K-major for the forward product -- inside ONE flat
    parameter buffer; gradients and both Adam slots are flat buffers of the same layout, so
    the optimiser is one fused launch and data-parallel training is one all-reduce;
  * likelihood-head weights of the P heads are concatenated: (P * Gn, Hp), Gn = round4(G),
    and their pre-activations land in one (rows, P * Gn) buffer.
"""

import math
from collections import OrderedDict

import torch

from . import kernels as K

ADAM_BETA1, ADAM_BETA2, ADAM_EPSILON = 0.9, 0.999, 1e-8   # tf.train.AdamOptimizer defaults
GRADIENT_CLIP = 1.0                                          # VAE:2753


def round4(n):
    return (n + 3) & ~3


def aug(n):
    """Stored width of an augmented activation with n logical columns."""
    return round4(n + 1)


class _Layer:
    """One dense layer: weight (n_out, in_p) with the bias in column n_in, optional BN.
    ``n_extra`` further input columns behind the bias column carry the decoder-input extras
    (one-hot batch index, count sum; VAE:2400-2441): data, so no gradient flows into them."""

    def __init__(self, name, n_in, n_out, bn, n_extra=0):
        self.name, self.n_in, self.n_out, self.bn = name, n_in, n_out, bn
        self.n_extra = n_extra
        self.k_in = n_in + 1 + n_extra          # reduction length of the forward product
        self.in_p = round4(self.k_in)
        self.w = self.dw = None          # views into the flat buffers
        self.beta = self.dbeta = None
        self.moving_mean = self.moving_var = None


class ParameterStore:
    """Flat fp32 buffers: parameters, gradients, Adam m / v; 16-byte aligned views."""

    def __init__(self, device):
        self.device = device
        self.specs = []     # (key, shape)
        self.offsets = {}
        self.total = 0

    def add(self, key, shape):
        n = 1
        for s in shape:
            n *= s
        self.offsets[key] = (self.total, tuple(shape))
        self.specs.append(key)
        self.total += round4(n)

    def allocate(self):
        dev = self.device
        self.param = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.m = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.v = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.step = torch.zeros(1, dtype=torch.int64, device=dev)

    def view(self, buf, key):
        off, shape = self.offsets[key]
        n = 1
        for s in shape:
            n *= s
        return buf[off:off + n].view(shape)


class VAEEngine:
    def __init__(self, feature_size, latent_size, hidden_sizes=(100,),
                 reconstruction_distribution="poisson", latent_distribution="gaussian",
                 minibatch_normalisation=True, kl_weight=1.0, device="cuda", seed=0,
                 tensor_cores=True, fused_heads=True, number_of_batches=0, count_sum_feature=False,
                 inference_architecture="MLP", generative_architecture="MLP",
                 number_of_reconstruction_classes=0, analytical_kl_term=True,
                 dropout_keep_probabilities=None):
        if reconstruction_distribution not in K.LIKELIHOOD_KINDS:
            raise ValueError("reconstruction distribution `{}` is not supported by the "
                             "B200 hot path".format(reconstruction_distribution))
        if latent_distribution not in ("gaussian", "unit-variance gaussian"):
            raise ValueError("latent distribution `{}` not supported for the VAE".format(
                latent_distribution))
        self.G, self.L = int(feature_size), int(latent_size)
        self.hidden_sizes = [int(h) for h in hidden_sizes]
        self.kind_name = reconstruction_distribution
        self.kind = K.LIKELIHOOD_KINDS[reconstruction_distribution]
        self.heads = K.LIKELIHOOD_HEADS[reconstruction_distribution]
        self.P = len(self.heads)
        self.unit_variance = latent_distribution == "unit-variance gaussian"
        # sampled KL term, log q(z|x) - log p(z) at the drawn z (VAE:2628-2640), instead of the
        # closed form (VAE:2624-2627); one KL value per (sample, cell) row
        self.sampled_kl = not bool(analytical_kl_term)
        # dropout keep probabilities [hidden, x, z] (VAE:246-269; a scalar means hidden only);
        # False / None / 0 / 1 switch a kind off (MU:45-46)
        keep = dropout_keep_probabilities
        if isinstance(keep, (list, tuple)):
            keep = list(keep) + [False] * (3 - len(keep))
        else:
            keep = [keep, False, False]
        self.keep_h, self.keep_x, self.keep_z = [
            float(k) if (k and k != 1) else None for k in keep[:3]]
        self.dropout_active = any(k is not None for k in (self.keep_h, self.keep_x, self.keep_z))
        self.dropout_seed = int(seed) + 104729
        self.bn = bool(minibatch_normalisation)
        self.kl_weight = float(kl_weight)
        self.device = torch.device(device)
        self.tensor_cores = bool(tensor_cores)
        # constrained Poisson: softmax over the genes of a cell, N = count sum of the cell as a
        # parameter (VAE:2492-2496) -- a row kernel of its own, never the fused heads
        self.constrained = self.kind == K.CONSTRAINED_POISSON
        # continuous / binary reconstruction distributions: row kernels of their own
        self.continuous = self.kind in K.CONTINUOUS_KINDS
        # piecewise-categorical likelihood (`-k`, CAT:210-274): k_max + 1 class-logit heads
        # behind the P heads of the count distribution, a row kernel of its own
        self.k_max = int(number_of_reconstruction_classes or 0)
        if self.k_max and (self.constrained or self.continuous):
            raise ValueError("piecewise-categorical likelihoods wrap the Poisson / NB family")
        self.PT = self.P + (self.k_max + 1 if self.k_max else 0)      # head blocks of width Gn
        # (dropout: every head multiplies its own dropped copy of the decoder output, so neither
        # the fused heads kernel -- one shared operand -- nor the 16-bit-only minibatch assembly
        # that feeds it applies)
        self.fused_heads = (bool(fused_heads) and self.tensor_cores and not self.constrained
                            and not self.continuous and not self.k_max and not self.dropout_active)
        self.Gn = round4(self.G)
        self.Gp = aug(self.G)
        self.Gh = (self.G + 63) & ~63      # head stride of the fp16 buffers of the fused heads
        self.world_size = 1
        self._all_reduce = None
        self._plans = {}
        self._side = None                  # second stream: head weight gradients / shadow copies
        self.overlap_streams = True
        self.side_gemm_ctas = int(__import__("os").environ.get("SCVAE_SIDE_GEMM_CTAS", "116"))
        # the (cells x ~100) middle of a 16-bit training / lean evaluation step as one persistent
        # kernel per direction (csrc/mid_layers.cu) instead of ~20 small launches
        self.mid_fused = __import__("os").environ.get("SCVAE_MID_FUSED", "1") != "0"
        # first-layer weight gradient dW1 = dY1^T X behind the middle kernel: dY1 as ONE loss-scaled
        # fp16 matrix (default; measured error 2e-4 .. 4e-4 of max|dW1| at the C2 / C3 shapes, the bias
        # column comes from the fp32 values) or, SCVAE_DY1_SPLIT=1, as fp16 + rounding remainder (two
        # MMA sets, 1.5 x the operand tiles, ~2 % of the step)
        self.dy1_split = __import__("os").environ.get("SCVAE_DY1_SPLIT", "0") != "0"
        # CTAs of the backward middle kernel when it shares the SMs with the side-stream GEMMs
        # (0: never share, run it first on all SMs)
        self.mid_bwd_ctas = int(__import__("os").environ.get("SCVAE_MID_BWD_CTAS", "74"))
        self.mid_fwd_ctas = int(__import__("os").environ.get("SCVAE_MID_FWD_CTAS", "148"))
        # CTAs of the head-slice gradient exchange on the side stream (data parallel).  The kernel needs
        # no shared memory, so its CTAs sit beside the tensor-core GEMMs' (2 GPUs: 58 / 100 / 148 CTAs ->
        # 0.626 / 0.610 / 0.608 ms per step)
        self.dp_side_ctas = int(__import__("os").environ.get("SCVAE_DP_SIDE_CTAS", "148"))

        for arch in (inference_architecture, generative_architecture):
            if arch not in ("MLP", "LFM"):
                raise ValueError("architectures are `MLP` or `LFM` (linear factor model)")
        # LFM: no hidden layers on that side (VAE:2221-2239, :2443-2462)
        self.lfm_inference = inference_architecture == "LFM"
        self.lfm_generative = generative_architecture == "LFM"
        # decoder-input extras concatenated to z (VAE:2400-2441)
        self.number_of_batches = int(number_of_batches or 0)
        self.count_sum_feature = bool(count_sum_feature)
        self.n_extra = self.number_of_batches + (1 if self.count_sum_feature else 0)
        n = len(self.hidden_sizes)
        self.enc = []
        width = self.G
        for i, h in enumerate([] if self.lfm_inference else self.hidden_sizes):
            self.enc.append(_Layer("ENCODER/{}".format(i + 1), width, h, self.bn))
            width = h
        self.nL = self.L if self.unit_variance else 2 * self.L
        self.post = _Layer("POSTERIOR", width, self.nL, False)
        self.dec = []
        width = self.L
        for i, h in enumerate([] if self.lfm_generative else self.hidden_sizes[::-1]):
            self.dec.append(_Layer("DECODER/{}".format(n - i), width, h, self.bn,
                                   n_extra=self.n_extra if i == 0 else 0))
            width = h
        self.head = _Layer("X_TILDE", width, self.PT * self.Gn, False,
                           n_extra=0 if self.dec else self.n_extra)
        self.Zp = round4(self.L + 1 + self.n_extra)       # stored width of the latent samples

        store = ParameterStore(self.device)
        for layer in self.enc + [self.post] + self.dec + [self.head]:
            store.add(layer.name + "/W", (layer.n_out, layer.in_p))
            if layer.bn:
                store.add(layer.name + "/beta", (layer.n_out,))
        store.allocate()
        self.store = store
        self.state = {}
        self._peer = None                  # distributed.PeerExchange (fused exchange + optimiser)
        # {learning rate, warm-up weight} of the current step on the device: kernels of a captured
        # step read them here, so one CUDA graph serves every epoch of a warm-up schedule
        self.scalars = torch.tensor([0.0, 1.0], dtype=torch.float32, device=self.device)
        # CTA counter of the optimiser launches of one step (scvae_adam_clip_step advance_counter).
        # Created here, not lazily: its first use is on the side stream, which does not wait for
        # a fill enqueued on the main stream
        self._adam_counter = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._scalars_host = (None, None)
        for layer in self.enc + [self.post] + self.dec + [self.head]:
            if layer.bn:
                layer.moving_mean = torch.zeros(layer.n_out, dtype=torch.float32,
                                                device=self.device)
                layer.moving_var = torch.ones(layer.n_out, dtype=torch.float32,
                                              device=self.device)
        self._bind_views()
        self.initialise(seed)

    def _bind_views(self):
        store = self.store
        for layer in self.enc + [self.post] + self.dec + [self.head]:
            layer.w = store.view(store.param, layer.name + "/W")
            layer.dw = store.view(store.grad, layer.name + "/W")
            if layer.bn:
                layer.beta = store.view(store.param, layer.name + "/beta")
                layer.dbeta = store.view(store.grad, layer.name + "/beta")

    def rebind_flat_buffers(self, param, grad):
        """Move the flat parameter / gradient buffers (e.g. into symmetric memory); the values
        must already be in place.  Plans hold no views of them; captured graphs do and must be
        re-captured (TrainLoop objects created before this call are stale)."""
        assert param.numel() == self.store.total and grad.numel() == self.store.total
        self._shadow_valid = False
        self.store.param, self.store.grad = param, grad
        self._bind_views()

    def set_peer_exchange(self, peer):
        self._peer = peer

    # ------------------------------------------------------------------ parameters ---------
    def _tf_names(self):
        """(engine layer, row slice, TF scope) triples in reference variable order."""
        out = []
        for layer in self.enc:
            out.append((layer, slice(0, layer.n_out), layer.name))
        if self.unit_variance:
            out.append((self.post, slice(0, self.L), "POSTERIOR/MU"))
        else:
            out.append((self.post, slice(0, self.L), "POSTERIOR/MU"))
            out.append((self.post, slice(self.L, 2 * self.L), "POSTERIOR/LOG_SIGMA"))
        for layer in self.dec:
            out.append((layer, slice(0, layer.n_out), layer.name))
        for p, head in enumerate(self.heads):
            out.append((self.head, slice(p * self.Gn, p * self.Gn + self.G),
                        "X_TILDE/" + head.upper()))
        return out

    def initialise(self, seed=0):
        """Xavier-uniform weights, zero biases (tf.contrib fully_connected defaults), BN
        beta 0 / moving mean 0 / moving variance 1; Adam slots and step reset."""
        gen = torch.Generator().manual_seed(int(seed))
        params = OrderedDict()
        for layer, rows, scope in self._tf_names():
            fan_in, fan_out = layer.n_in + layer.n_extra, rows.stop - rows.start
            limit = math.sqrt(6.0 / (fan_in + fan_out))
            w = torch.rand((fan_in, fan_out), generator=gen, dtype=torch.float64)
            params[scope + "/DENSE/weights"] = ((2.0 * w - 1.0) * limit).float()
            params[scope + "/DENSE/biases"] = torch.zeros(fan_out)
        if self.k_max:      # one FC of width G (k_max + 1) (VAE:2507-2518)
            fan_in, fan_out = self.head.n_in + self.head.n_extra, self.G * (self.k_max + 1)
            limit = math.sqrt(6.0 / (fan_in + fan_out))
            w = torch.rand((fan_in, fan_out), generator=gen, dtype=torch.float64)
            params["X_TILDE/P_K/DENSE/weights"] = ((2.0 * w - 1.0) * limit).float()
            params["X_TILDE/P_K/DENSE/biases"] = torch.zeros(fan_out)
        self.import_parameters(params, strict=False)
        for buf in (self.store.grad, self.store.m, self.store.v):
            buf.zero_()
        self.store.step.zero_()
        for layer in self.enc + self.dec:
            if layer.bn:
                layer.beta.zero_()
                layer.moving_mean.zero_()
                layer.moving_var.fill_(1.0)

    def import_parameters(self, params, strict=True):
        """Load variables given by their TF names in the reference layout (weights (in,out))."""
        self._shadow_valid = False
        for layer, rows, scope in self._tf_names():
            w = params[scope + "/DENSE/weights"].to(self.device, torch.float32)
            b = params[scope + "/DENSE/biases"].to(self.device, torch.float32)
            layer.w[rows, :layer.n_in] = w[:layer.n_in].t()
            layer.w[rows, layer.n_in] = b
            layer.w[rows, layer.n_in + 1:] = 0
            if layer.n_extra:
                layer.w[rows, layer.n_in + 1:layer.k_in] = w[layer.n_in:].t()
            if layer.bn:
                for key, dst in (("beta", layer.beta), ("moving_mean", layer.moving_mean),
                                 ("moving_variance", layer.moving_var)):
                    name = scope + "/BATCH_NORM/" + key
                    if name in params:
                        dst.copy_(params[name].to(self.device, torch.float32))
                    elif strict:
                        raise KeyError(name)
        if self.k_max:
            # the reference's P_K variable is class-minor (column g (k_max + 1) + c, reshaped to
            # (rows, G, k_max + 1)); the engine keeps one block of Gn rows per class
            K1, layer = self.k_max + 1, self.head
            w = params["X_TILDE/P_K/DENSE/weights"].to(self.device, torch.float32)
            b = params["X_TILDE/P_K/DENSE/biases"].to(self.device, torch.float32)
            for c in range(K1):
                rows = self._pk_rows(c)
                wc = w[:, c::K1]
                layer.w[rows, :layer.n_in] = wc[:layer.n_in].t()
                layer.w[rows, layer.n_in] = b[c::K1]
                layer.w[rows, layer.n_in + 1:] = 0
                if layer.n_extra:
                    layer.w[rows, layer.n_in + 1:layer.k_in] = wc[layer.n_in:].t()

    def _pk_rows(self, c):
        return slice((self.P + c) * self.Gn, (self.P + c) * self.Gn + self.G)

    def _export_pk(self, buf, out):
        if not self.k_max:
            return
        K1, layer = self.k_max + 1, self.head
        w = torch.zeros(layer.n_in + layer.n_extra, self.G * K1)
        b = torch.zeros(self.G * K1)
        for c in range(K1):
            rows = self._pk_rows(c)
            w[:, c::K1] = self._tf_weight(buf, layer, rows)
            b[c::K1] = buf[rows, layer.n_in].cpu()
        out["X_TILDE/P_K/DENSE/weights"] = w
        out["X_TILDE/P_K/DENSE/biases"] = b

    @staticmethod
    def _tf_weight(buf, layer, rows):
        """(in [+ extras], out) matrix in the reference layout from the stored (out, in_p) one."""
        w = buf[rows, :layer.n_in]
        if layer.n_extra:
            w = torch.cat([w, buf[rows, layer.n_in + 1:layer.k_in]], dim=1)
        return w.t().contiguous().cpu()

    def export_parameters(self):
        out = OrderedDict()
        for layer, rows, scope in self._tf_names():
            out[scope + "/DENSE/weights"] = self._tf_weight(layer.w, layer, rows)
            out[scope + "/DENSE/biases"] = layer.w[rows, layer.n_in].contiguous().cpu()
            if layer.bn:
                out[scope + "/BATCH_NORM/beta"] = layer.beta.cpu().clone()
                out[scope + "/BATCH_NORM/moving_mean"] = layer.moving_mean.cpu().clone()
                out[scope + "/BATCH_NORM/moving_variance"] = layer.moving_var.cpu().clone()
        self._export_pk(self.head.w, out)
        return out

    def export_gradients(self):
        """Last computed raw gradients (before clipping) under the TF variable names."""
        out = OrderedDict()
        for layer, rows, scope in self._tf_names():
            out[scope + "/DENSE/weights"] = self._tf_weight(layer.dw, layer, rows)
            out[scope + "/DENSE/biases"] = layer.dw[rows, layer.n_in].contiguous().cpu()
            if layer.bn:
                out[scope + "/BATCH_NORM/beta"] = layer.dbeta.cpu().clone()
        self._export_pk(self.head.dw, out)
        return out

    def exchanged_ranges(self):
        """Flat ranges exchanged separately by optimiser_step (peer mode owns a slice of each)."""
        off = self.store.offsets[self.head.name + "/W"][0]
        return [(0, off), (off, self.store.total)] if self.fused_heads and self.overlap_streams \
            else [(0, self.store.total)]

    def state_dict(self):
        if self._peer is not None and self.world_size > 1:
            self._peer.gather_slots(self._last_ranges if getattr(self, "_last_ranges", None)
                                    else self.exchanged_ranges())
        sd = {"param": self.store.param.cpu(), "m": self.store.m.cpu(), "v": self.store.v.cpu(),
              "step": self.store.step.cpu()}
        for layer in self.enc + self.dec:
            if layer.bn:
                sd[layer.name + "/moving_mean"] = layer.moving_mean.cpu()
                sd[layer.name + "/moving_variance"] = layer.moving_var.cpu()
        return sd

    def load_state_dict(self, sd):
        self._shadow_valid = False
        self.store.param.copy_(sd["param"])
        self.store.m.copy_(sd["m"])
        self.store.v.copy_(sd["v"])
        self.store.step.copy_(sd["step"])
        for layer in self.enc + self.dec:
            if layer.bn:
                layer.moving_mean.copy_(sd[layer.name + "/moving_mean"])
                layer.moving_var.copy_(sd[layer.name + "/moving_variance"])

    def bn_layers(self):
        return [l for l in self.enc + self.dec if l.bn]

    @property
    def global_step(self):
        return int(self.store.step.item())

    # ------------------------------------------------------------------ buffers ------------
    def _plan(self, B, RS):
        key = (B, RS)
        if key in self._plans:
            return self._plans[key]
        dev, f32 = self.device, torch.float32
        M = RS * B
        p = type("Plan", (), {})()
        p.B, p.RS, p.M = B, RS, M

        def zeros(*shape):
            return torch.zeros(*shape, dtype=f32, device=dev)

        p.X = zeros(B, self.Gp)
        p.X[:, self.G] = 1.0
        p.T = None                       # separate targets only when t != x
        p.row_const = zeros(B)
        p.have_row_const = False
        p.encY = [zeros(B, round4(l.n_out)) for l in self.enc]
        p.encH = [zeros(B, aug(l.n_out)) for l in self.enc]
        p.enc_mean = [zeros(l.n_out) for l in self.enc]
        p.enc_rstd = [zeros(l.n_out) for l in self.enc]
        p.PH = zeros(B, round4(self.nL))
        p.eps = zeros(M, self.L)
        p.Z = zeros(M, self.Zp)
        p.count_sum_parameter = zeros(B) if self.constrained else None   # N of the constrained Poisson
        p.lse = zeros(M) if self.constrained else None
        p.batch_index = zeros(B) if self.number_of_batches else None   # batch ids as floats
        p.count_sum = zeros(B) if self.count_sum_feature else None
        p.kl_row = zeros(B)
        p.kl_rows = zeros(M) if self.sampled_kl else None
        p.drop = {}                      # dropout sites: noise, dropped operand copy
        p.drop_on = False
        p.drop_injected = False          # tests: noise tensors set by hand, not drawn
        p.kl_elem = zeros(B, self.L)
        p.kl_neurons = zeros(self.L)
        p.decY = [zeros(M, round4(l.n_out)) for l in self.dec]
        p.decH = [zeros(M, aug(l.n_out)) for l in self.dec]
        p.dec_mean = [zeros(l.n_out) for l in self.dec]
        p.dec_rstd = [zeros(l.n_out) for l in self.dec]
        p.A = zeros(M, self.PT * self.Gn)
        p.logp = zeros(M)
        p.go = zeros(M)
        p.bound = zeros(4)
        # backward
        p.dA = None
        p.bwd_ready = False
        scratch = 1
        for l, rows in [(l, B) for l in self.enc] + [(l, M) for l in self.dec]:
            scratch = max(scratch, K.bn_scratch_floats(rows, l.n_out, 1))
        p.bn_scratch = zeros(scratch)
        p.workspace = None
        p.ws_bytes = 0
        p.have_t16 = False
        p.have_x = True
        p.t16_is_x16 = False
        p.fused_ready = False
        p.fused_done = False
        self._plans[key] = p
        return p

    # ---- fused likelihood heads (heads_fused.cu) -------------------------------------------
    def _fused_possible(self, M, B):
        """n_in + 1 <= 128 hidden columns, genes a multiple of 8, targets tiling in 128 rows."""
        return (self.fused_heads and self.head.k_in <= 128 and self.G % 8 == 0
                and (M == B or B % 128 == 0))

    def _plan_fused(self, p, M, backward=True):
        """Buffers of the fused heads kernel, sized for the plan's row count; the fp16 gradient
        buffer (rows x P genes) only when a backward pass will follow."""
        dev = self.device
        rows = max(M, p.M)
        if not p.fused_ready:
            p.D16 = torch.zeros(rows, 128, dtype=torch.float16, device=dev)
            self._shadow_buffers(p)
            p.fused_ws = torch.zeros(K.heads_fused_workspace_floats(rows, self.G),
                                     dtype=torch.float32, device=dev)
            p.dA16 = None
            p.fused_ready = True
        if backward and p.dA16 is None:
            p.dA16 = torch.zeros(rows, self.P * self.Gh, dtype=torch.float16, device=dev)

    def _t16(self, p):
        if getattr(p, "T16", None) is None:
            p.T16 = torch.zeros(p.B, self.Gh, dtype=torch.int16, device=self.device)
        return p.T16

    def _x16(self, p):
        """fp16 augmented copy of the minibatch: operand of the fp16 first layer and, when all
        counts are <= 2048 (exact in fp16), also the target matrix of the fused heads."""
        if getattr(p, "X16", None) is None:
            # row pitch = a multiple of 128 bytes: every 128-byte row of a TMA box is then ONE aligned
            # cache line (with the 16-byte-aligned pitch of G = 20 000 each row straddled two lines and
            # odd rows five 32-byte sectors: dW1 = dY1^T X 60 -> 42 us, tools/wgrad_probe.py)
            w = (self.G + 8) & ~7
            p.X16 = torch.zeros(p.B, (w + 63) & ~63, dtype=torch.float16, device=self.device)[:, :w]
        return p.X16

    def _plan_backward(self, p):
        if p.bwd_ready:
            return
        dev, f32 = self.device, torch.float32

        def zeros(*shape):
            return torch.zeros(*shape, dtype=f32, device=dev)

        B, M = p.B, p.M
        p.dA = zeros(M, self.PT * self.Gn)
        p.d_decH = [zeros(M, aug(l.n_out)) for l in self.dec]
        p.d_decY = [zeros(M, round4(l.n_out)) for l in self.dec]
        p.dZ = zeros(M, self.Zp)
        p.dPH = zeros(B, round4(self.nL))
        p.d_encH = [zeros(B, aug(l.n_out)) for l in self.enc]
        p.d_encY = [zeros(B, round4(l.n_out)) for l in self.enc]
        p.bwd_ready = True

    def _side_stream(self):
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def _shadow_buffers(self, p):
        """fp16 operand copies of the gene-axis weights.  They depend on the parameters only, so
        the ENGINE owns them (plans alias them): head weights W16, first encoder weight as fp16 +
        rounding remainder (scvae_gemm_f16_split)."""
        if getattr(self, "W16", None) is None:
            dev = self.device
            self.W16 = torch.zeros(self.P * self.Gh, 128, dtype=torch.float16, device=dev)
            self.W1_16 = self.W1_16_lo = None
            if self.enc:
                self.W1_16 = torch.zeros(self.enc[0].n_out, (self.G + 8) & ~7, dtype=torch.float16,
                                         device=dev)
                self.W1_16_lo = torch.zeros_like(self.W1_16)
            self._shadow_valid = False
        p.W16, p.W1_16, p.W1_16_lo = self.W16, self.W1_16, self.W1_16_lo

    def invalidate_shadows(self):
        """The parameters changed behind the optimiser kernel's back (import, restore, broadcast)."""
        self._shadow_valid = False

    def _adam_shadows(self, lo, hi):
        """scvae_shadow descriptors of the shadowed blocks inside the flat range [lo, hi): the
        optimiser kernel rewrites the fp16 copies in the pass that updates the masters."""
        if getattr(self, "W16", None) is None:
            return []
        out = []
        if self.enc:
            l = self.enc[0]
            off = self.store.offsets[l.name + "/W"][0]
            if lo <= off and off + l.n_out * l.in_p <= hi:
                out.append(K.shadow(off - lo, off - lo + l.n_out * l.in_p, l.in_p, l.n_in + 1,
                                    self.W1_16, self.W1_16_lo))
        l = self.head
        off = self.store.offsets[l.name + "/W"][0]
        if lo <= off and off + self.P * self.Gn * l.in_p <= hi:
            out.append(K.shadow(off - lo, off - lo + self.P * self.Gn * l.in_p, l.in_p, l.in_p,
                                self.W16, None, src_block_rows=self.Gn, dst_block_rows=self.Gh))
        return out

    def _refresh_shadows(self, p, M):
        """fp16 copies of the first encoder weight and of the head weights for this step -- only
        when the last optimiser step did not already rewrite them (scvae_adam_clip_step shadows).
        They depend on the parameters only, so they run on the side stream beside the
        densify / noise kernels that precede the first product; the main stream joins here."""
        self._plan_fused(p, M, backward=False)
        if self._shadow_valid:
            p.shadow_fork = None
            return
        self._shadow_valid = True
        main = torch.cuda.current_stream()
        use_side = self.overlap_streams and getattr(p, "shadow_fork", None) is not None
        if use_side:
            side = self._side_stream()
            side.wait_event(p.shadow_fork)
            ctx = torch.cuda.stream(side)
        else:
            import contextlib
            ctx = contextlib.nullcontext()
        with ctx:
            if self.enc:
                l = self.enc[0]
                # weights as fp16 + their fp16 rounding remainder: the counts are exact in fp16,
                # so x W1^T keeps ~22 bits of the fp32 weights (scvae_gemm_f16_split)
                K.f32_to_f16_split(l.w, l.n_in + 1, p.W1_16, p.W1_16_lo)
            l = self.head
            for h in range(self.P):
                K.f32_to_f16(l.w[h * self.Gn:(h + 1) * self.Gn], l.in_p,
                             p.W16[h * self.Gh:h * self.Gh + self.Gn])
            if use_side:
                p.shadow_join.record(side)
        if use_side:
            main.wait_event(p.shadow_join)
            p.shadow_fork = None

    def fork_shadows(self, p):
        """Call on the main stream BEFORE enqueueing the minibatch assembly of a training step:
        marks the point after which the weight shadow copies may start on the side stream."""
        if not self.overlap_streams:
            return
        if getattr(p, "shadow_join", None) is None:
            p.shadow_join = torch.cuda.Event()
            p.shadow_fork_ev = torch.cuda.Event()
        p.shadow_fork_ev.record(torch.cuda.current_stream())
        p.shadow_fork = p.shadow_fork_ev

    def _use_tc(self, M, N, Kd):
        """tcgen05 kernel for every product with a long reduction or a very large output; the
        exact-fp32 FFMA kernel keeps the hidden-layer-sized ones."""
        # (kind::tf32 truncates its fp32 operands -- a -5e-4 relative bias per product that batch norm
        # does not remove behind the posterior heads -- so the small products stay exact)
        return self.tensor_cores and (Kd >= 512 or M * N * Kd >= (1 << 28))

    def _gemm(self, p, layout, M, N, Kd, A, Bm, C, accumulate=False):
        tc = self._use_tc(M, N, Kd)
        ws = None
        if tc:
            need = K.gemm_workspace_bytes(layout, M, N, Kd)
            if need > 0:
                if p.ws_bytes < need:
                    p.workspace = torch.empty(need // 4, dtype=torch.float32, device=self.device)
                    p.ws_bytes = need
                ws = p.workspace
        K.gemm(layout, M, N, Kd, A, Bm, C, accumulate=accumulate, tensor_cores=tc, workspace=ws)

    def _gemm16(self, p, layout, M, N, Kd, A, Bm, C, accumulate=False, alpha=1.0):
        need = K.gemm_f16_workspace_bytes(layout, M, N, Kd)
        ws = None
        if need > 0:
            if p.ws_bytes < need:
                p.workspace = torch.empty(need // 4, dtype=torch.float32, device=self.device)
                p.ws_bytes = need
            ws = p.workspace
        K.gemm_f16(layout, M, N, Kd, A, Bm, C, accumulate=accumulate, alpha=alpha, workspace=ws)

    # ---- fused middle (mid_layers.cu) -------------------------------------------------------
    def _mid_possible(self, p, M, drop):
        """Hidden widths < 128, latent size <= 128 (padded latent row <= 128), one sample per cell,
        Gaussian posterior with the analytic KL, no dropout, cells <= 64 per SM."""
        if not (self.mid_fused and self.tensor_cores and M == p.B and self.enc and self.dec
                and not drop and not self.unit_variance and not self.sampled_kl):
            return False
        if len(self.enc) > 4 or len(self.dec) > 4 or self.Zp > 128 or self.post.in_p > 128:
            return False
        if any(l.n_out >= 128 or (i > 0 and l.in_p > 128) for i, l in enumerate(self.enc)):
            return False
        if any(l.n_out >= 128 or l.in_p > 128 for l in self.dec):
            return False
        return p.B <= 64 * self._sm_count()

    def _sm_count(self):
        if getattr(self, "_sms", None) is None:
            self._sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        return self._sms

    def _mid_bwd_grid(self, p):
        """CTAs of the backward middle kernel when it runs beside the side-stream GEMMs (0: it
        does not fit beside them: more than half of the SMs)."""
        if not (self.mid_bwd_ctas > 0 and self.overlap_streams):
            return 0
        rows = self._mid_rows(p.B, min(self._sm_count(), self.mid_bwd_ctas))
        grid = -(-p.B // rows)
        return grid if grid <= self._sm_count() // 2 else 0

    def _mid_rows(self, B, ctas):
        """Cells per CTA (a multiple of 4, <= 64) so that at most ``ctas`` CTAs cover B cells."""
        rows = max(4, -(-B // max(ctas, 1)))
        # a CTA covers 32 cells per pass whatever it is given: never spread a small minibatch
        # over more CTAs (more weight staging, wider grid barriers) than that needs
        rows = max(rows, min(32, B))
        return min(64, (rows + 3) & ~3)

    def _mid_desc(self, p, backward):
        """The descriptor of vae_mid_fwd / vae_mid_bwd for this plan (all pointers are static)."""
        key = "_mid_bwd" if backward else "_mid_fwd"
        d = getattr(p, key, None)
        if d is not None:
            return d
        dev, B, L = self.device, p.B, self.L
        if getattr(p, "mid_bar", None) is None:
            p.mid_bar = torch.zeros(4, dtype=torch.int32, device=dev)
            p.mid_err = torch.zeros(4, dtype=torch.int32, device=dev)
            p.mid_ws = None
        self._plan_fused(p, p.M, backward=backward)
        d = K.mid_desc()
        d.B, d.L, d.n_enc, d.n_dec = B, L, len(self.enc), len(self.dec)
        sms = self._sm_count()
        d.rows_per_cta = self._mid_rows(B, min(sms, self.mid_bwd_ctas) if (
            backward and self.mid_bwd_ctas > 0 and self.overlap_streams) else min(sms, self.mid_fwd_ctas))
        for i, l in enumerate(self.enc):
            d.enc[i] = K.mid_layer(
                w=l.w if i else None, dw=l.dw if (i or (backward and not self.dy1_split)) else None,
                beta=l.beta if l.bn else None,
                dbeta=l.dbeta if l.bn else None, moving_mean=l.moving_mean, moving_var=l.moving_var,
                mean=p.enc_mean[i], rstd=p.enc_rstd[i], y=p.encY[i], n_in=l.n_in, k_in=l.k_in,
                n_out=l.n_out)
        l = self.post
        d.post = K.mid_layer(w=l.w, dw=l.dw, n_in=l.n_in, k_in=l.k_in, n_out=l.n_out)
        for j, l in enumerate(self.dec):
            d.dec[j] = K.mid_layer(
                w=l.w, dw=l.dw, beta=l.beta if l.bn else None, dbeta=l.dbeta if l.bn else None,
                moving_mean=l.moving_mean, moving_var=l.moving_var, mean=p.dec_mean[j],
                rstd=p.dec_rstd[j], y=p.decY[j], n_in=l.n_in, k_in=l.k_in, n_out=l.n_out)
        d.ph, d.ldph = p.PH.data_ptr(), p.PH.stride(0)
        d.eps = p.eps.data_ptr()
        d.z, d.ldz = p.Z.data_ptr(), p.Z.stride(0)
        d.kl_row, d.kl_elem = p.kl_row.data_ptr(), p.kl_elem.data_ptr()
        if self.number_of_batches:
            d.batch_index, d.n_batches = p.batch_index.data_ptr(), self.number_of_batches
        if self.count_sum_feature:
            d.count_sum = p.count_sum.data_ptr()
        d.d16, d.ldd16 = p.D16.data_ptr(), p.D16.stride(0)
        d.kl_weight = self.kl_weight
        d.scalars = self.scalars.data_ptr()
        d.barrier, d.error = p.mid_bar.data_ptr(), p.mid_err.data_ptr()
        if __import__("os").environ.get("SCVAE_MID_TIMELINE"):       # development aid
            tl = torch.zeros(160 * 32, dtype=torch.int64, device=dev)
            setattr(p, key + "_timeline", tl)
            d.timeline = tl.data_ptr()
        if backward:
            self._plan_backward(p)
            self._dy1_16(p)
            d.logp, d.bound = p.logp.data_ptr(), p.bound.data_ptr()
            d.dy1_16, d.lddy1 = p.dY1_16.data_ptr(), p.dY1_16.stride(0)
            d.dy1_16_lo = p.dY1_16_lo.data_ptr() if self.dy1_split else None
        setattr(p, key, d)
        need = K.vae_mid_workspace_floats(d)
        if p.mid_ws is None or p.mid_ws.numel() < need:
            p.mid_ws = torch.zeros(need, dtype=torch.float32, device=dev)
        for other in ("_mid_fwd", "_mid_bwd"):        # (both share the workspace)
            o = getattr(p, other, None)
            if o is not None:
                o.workspace, o.workspace_floats = p.mid_ws.data_ptr(), p.mid_ws.numel()
        return d

    def step_on_device_scalars(self, p, R, S, u16_ok):
        """True when a training step on this plan takes the 16-bit path with the fused middle: its
        kernels read the warm-up weight (and draw the noise) on the device, so ONE captured graph
        serves every epoch of a warm-up schedule."""
        M = R * S * p.B
        return bool(R == 1 and u16_ok and self.fused_heads and self._fused_possible(M, p.B)
                    and self._mid_possible(p, M, bool(self.dropout_active)))

    def mid_error(self, p):
        """True when a grid barrier of the fused middle timed out (results invalid)."""
        return getattr(p, "mid_err", None) is not None and bool(p.mid_err[0].item())

    def _gemm16_split(self, p, layout, M, N, Kd, A, Bm, X, which, C, alpha=1.0):
        need = K.gemm_f16_workspace_bytes(layout, M, N, Kd)
        ws = None
        if need > 0:
            if p.ws_bytes < need:
                p.workspace = torch.empty(need // 4, dtype=torch.float32, device=self.device)
                p.ws_bytes = need
            ws = p.workspace
        K.gemm_f16_split(layout, M, N, Kd, A, Bm, X, which, C, alpha=alpha, workspace=ws)

    def _wgrad1_16(self, p, l, B, single):
        """dW1 = dY1^T X on the tensor cores from the loss-scaled fp16 dY1: with its rounding
        remainder, or (``single``, behind vae_mid_bwd) from the one fp16 matrix."""
        if not single:
            self._gemm16_split(p, K.GEMM_TN, l.n_out, l.in_p, B, p.dY1_16, p.X16, p.dY1_16_lo, 1, l.dw,
                               alpha=1.0 / p.fused_scale)
        else:
            # the gene columns only: the bias column (= the column sums of dY1, zero up to rounding
            # behind a batch norm) was written by vae_mid_bwd from the fp32 values
            n = l.n_in
            self._gemm16(p, K.GEMM_TN, l.n_out, n, B, p.dY1_16, p.X16[:, :n], l.dw[:, :n],
                         alpha=1.0 / p.fused_scale)

    def _dy1_16(self, p):
        if getattr(p, "dY1_16", None) is None:
            w = (self.enc[0].n_out + 7) & ~7
            wp = (w + 63) & ~63                  # 128-byte row pitch, as for X16
            p.dY1_16 = torch.zeros(p.B, wp, dtype=torch.float16, device=self.device)[:, :w]
            p.dY1_16_lo = torch.zeros(p.B, wp, dtype=torch.float16, device=self.device)[:, :w]
        return p.dY1_16

    # ------------------------------------------------------------------ inputs -------------
    def set_batch_dense(self, p, x, t=None):
        """Dense (B, G) minibatch already on the device (tests / small data)."""
        p.X[:, :self.G].copy_(x)
        p.have_row_const = False
        p.have_t16 = False
        p.have_x = True
        if self.fused_heads and (t is None or t is x):
            # 16-bit copies are exact only for integer counts below 65536 (fp16: <= 2048)
            if bool(((x == x.round()) & (x >= 0) & (x <= 65504)).all()):
                p.t16_is_x16 = bool((x <= 2048).all())
                K.f32_to_f16(p.X, self.G + 1, self._x16(p))
                if not p.t16_is_x16:
                    K.f32_to_u16(p.X, self.G, self._t16(p))
                p.have_t16 = True
        if t is not None and t is not x:
            if p.T is None:
                p.T = torch.zeros(p.B, self.Gn, dtype=torch.float32, device=self.device)
            p.T[:, :self.G].copy_(t)
            p.use_T = True
        else:
            p.use_T = False

    def set_batch_csr(self, p, indptr, indices, values, rows=None, rebase=False, u16_ok=False,
                      f16_exact=False, train16=False, row_const_all=None):
        """Gather + densify B rows of a device-resident CSR matrix (a1).
        ``u16_ok``: counts are integers < 65536; ``f16_exact``: all counts <= 2048.
        ``train16``: the caller will run the fused 16-bit training step, so only the 16-bit
        copies are produced (fp16 input + uint16 targets when fp16 is not exact) and the fp32
        matrix is not written at all."""
        use16 = bool(train16 and u16_ok and self.fused_heads and self._fused_possible(p.M, p.B))
        # sum_g lgamma(1 + x) per cell: gathered from the per-data-set table when there is one
        rc_out = None if row_const_all is not None else p.row_const
        rc_join = None
        if row_const_all is not None and self.overlap_streams and self.device.type == "cuda":
            # the gather of the per-cell constants needs nothing from the densify: beside it on the
            # side stream instead of ~4 us behind it on the critical path
            main, side = torch.cuda.current_stream(), self._side_stream()
            if getattr(p, "rc_fork", None) is None:
                p.rc_fork, p.rc_done = torch.cuda.Event(), torch.cuda.Event()
            p.rc_fork.record(main)
            side.wait_event(p.rc_fork)
            with torch.cuda.stream(side):
                K.gather_f32(row_const_all, rows, p.row_const)
                p.rc_done.record(side)
            rc_join = p.rc_done
        if use16:
            p.t16_is_x16 = bool(f16_exact)
            # (LFM inference: the posterior heads read the fp32 minibatch directly)
            need_x = (not self.enc) or getattr(self, "_needs_fp32_x", False)
            K.csr_densify(indptr, indices, values, rows, self.G, p.X if need_x else None, rc_out,
                          rebase=rebase, t16=None if f16_exact else self._t16(p), x16=self._x16(p))
        else:
            K.csr_densify(indptr, indices, values, rows, self.G, p.X, rc_out, rebase=rebase)
        if rc_join is not None:
            torch.cuda.current_stream().wait_event(rc_join)
        elif row_const_all is not None:
            K.gather_f32(row_const_all, rows, p.row_const)
        p.have_x = (not use16) or (not self.enc) or getattr(self, "_needs_fp32_x", False)
        p.have_row_const = True
        p.have_t16 = use16
        p.use_T = False

    def set_batch_packed(self, p, slab, f16_exact, rows=None):
        """Minibatch from a packed row slab (``PackedStream``; scvae_csr_densify_packed): only the
        16-bit copies exist afterwards, so it serves the fused 16-bit passes (training with
        R = 1, lean evaluation passes).  ``rows`` = (r0, r1): the slab holds that row range of the
        minibatch only (the hybrid feeder delivers a minibatch as two slabs)."""
        if not (self.fused_heads and self._fused_possible(p.M, p.B)) or getattr(self, "_needs_fp32_x", False) \
                or not self.enc:
            raise NotImplementedError("packed row slabs feed the fused 16-bit path only")
        r0, r1 = rows if rows is not None else (0, p.B)
        p.t16_is_x16 = bool(f16_exact)
        K.csr_densify_packed(slab, r1 - r0, self.G, row_const=p.row_const[r0:r1],
                             t16=None if f16_exact else self._t16(p)[r0:r1], x16=self._x16(p)[r0:r1])
        p.have_x, p.have_row_const, p.have_t16, p.use_T = False, True, True, False

    # ------------------------------------------------------------------ forward ------------
    def forward(self, p, is_training, R, S, warm_up_weight=1.0, deterministic=False,
                update_moving=None, want_go=False, fused_backward=False, keep_heads=True):
        """encoder -> posterior -> z -> decoder -> heads -> log-likelihood -> bound.
        With ``fused_backward`` (R == 1) the likelihood kernel also emits dA in the same pass."""
        B = p.B
        RS = 1 if deterministic else R * S
        M = RS * B
        if update_moving is None:
            update_moving = is_training
        # keep_heads=False: the caller needs log p only, not the (rows x P genes) head
        # pre-activations (per-epoch evaluation passes): forward-only fused heads
        drop = p.drop_on = bool(self.dropout_active and is_training)
        use16 = bool((fused_backward or not keep_heads) and p.have_t16 and not drop
                     and not getattr(p, "use_T", False) and self._fused_possible(M, B))
        if not use16 and not getattr(p, "have_x", True):
            raise RuntimeError("this minibatch was densified for the fused 16-bit training step "
                               "only; call set_batch_csr without train16 for other passes")
        p.mid_done = False
        mid = bool(use16 and self._mid_possible(p, M, drop))
        if mid:
            # first encoder product on the tensor cores, everything up to the fp16 operand of the
            # fused heads in ONE persistent kernel (batch norm / ReLU / sample / KL as epilogues)
            l = self.enc[0]
            self._refresh_shadows(p, M)
            self._gemm16_split(p, K.GEMM_NT, B, l.n_out, l.n_in + 1, p.X16, p.W1_16, p.W1_16_lo, 2,
                               p.encY[0])
            md = self._mid_desc(p, backward=False)
            md.training, md.update_moving = int(bool(is_training)), int(bool(update_moving))
            md.deterministic = int(bool(deterministic))
            md.y1_parts, md.y1_ld = p.encY[0].data_ptr(), p.encY[0].stride(0)
            md.y1_slice, md.y1_nsplit, md.y1_alpha = 0, 1, 1.0
            src = getattr(p, "eps_source", None)
            if src is not None and not deterministic:
                # the noise is drawn inside the kernel (same Philox stream as scvae_fill_normal)
                md.generate_eps, md.seed, md.offset = 1, int(src[0]), 0
                md.offset_dev = src[1].data_ptr()
            else:
                md.generate_eps, md.offset_dev = 0, None
            K.vae_mid_fwd(md)
        h, h_cols = p.X, self.G
        for i, l in enumerate([] if mid else self.enc):
            if i == 0 and use16:
                # (cells x genes) operand in fp16: half the HBM traffic of the tf32 path
                self._refresh_shadows(p, M)
                self._gemm16_split(p, K.GEMM_NT, B, l.n_out, l.n_in + 1, p.X16, p.W1_16,
                                   p.W1_16_lo, 2, p.encY[i])
            else:
                keep = (self.keep_x if i == 0 else self.keep_h) if drop else None
                src = self._drop_site(p, l.name, h, B, l.n_in, l.n_in, keep).copy if keep else h
                self._gemm(p, K.GEMM_NT, B, l.n_out, l.n_in + 1, src, l.w, p.encY[i])
            if l.bn:
                K.bn_act_fwd(p.encY[i], l.n_out, l.beta, l.moving_mean, l.moving_var, p.encH[i],
                             p.enc_mean[i], p.enc_rstd[i], p.bn_scratch, training=is_training,
                             update_moving=update_moving, relu=True)
            else:
                K.act_fwd(p.encY[i], l.n_out, p.encH[i], relu=True)
            h = p.encH[i]
        l = self.post
        if mid:
            pass
        elif drop and self.keep_h:
            # one mask per posterior parameter over the same activation (VAE:2280-2289)
            L = self.L
            for part, site in enumerate(["POSTERIOR/MU"] + (
                    [] if self.unit_variance else ["POSTERIOR/LOG_SIGMA"])):
                st = self._drop_site(p, site, h, B, l.n_in, l.n_in, self.keep_h)
                K.gemm(K.GEMM_NT, B, L, l.n_in + 1, st.copy, l.w[part * L:(part + 1) * L],
                       p.PH[:, part * L:(part + 1) * L], tensor_cores=False)
        else:
            self._gemm(p, K.GEMM_NT, B, l.n_out, l.n_in + 1, h, l.w, p.PH)
        if not mid:
            K.gaussian_latent_fwd(p.PH, B, self.L, RS, p.eps, p.Z, p.kl_row, p.kl_elem,
                                  unit_variance=self.unit_variance, deterministic=deterministic)
            self._decoder_features(p, M)
        d = p.Z
        for j, l in enumerate([] if mid else self.dec):
            keep = (self.keep_z if j == 0 else self.keep_h) if drop else None
            src = self._drop_site(p, l.name, d, M, l.n_in + l.n_extra, l.n_in,
                                  keep).copy if keep else d
            self._gemm(p, K.GEMM_NT, M, l.n_out, l.k_in, src, l.w, p.decY[j][:M])
            if l.bn:
                K.bn_act_fwd(p.decY[j][:M], l.n_out, l.beta, l.moving_mean, l.moving_var,
                             p.decH[j][:M], p.dec_mean[j], p.dec_rstd[j], p.bn_scratch,
                             training=is_training, update_moving=update_moving, relu=True)
            else:
                K.act_fwd(p.decY[j][:M], l.n_out, p.decH[j][:M], relu=True)
            d = p.decH[j]
        l = self.head
        tgt = p.T if getattr(p, "use_T", False) else p.X
        rc = p.row_const if p.have_row_const else None
        weight = warm_up_weight * self.kl_weight
        p.fused_done = False
        if use16 and not fused_backward:
            # heads GEMM + likelihood in one kernel, nothing but log p leaves the chip
            self._plan_fused(p, M, backward=False)
            if not mid:
                K.f32_to_f16(d, l.in_p, p.D16)
            if not self.enc:
                self._refresh_shadows(p, M)
            t16 = p.X16 if p.t16_is_x16 else p.T16
            K.heads_fused_fwd(self.kind, p.D16, p.W16, self.Gh, t16, M, self.G, p.logp,
                              p.fused_ws, row_const=rc)
            self._bound(p, R, S, weight, p.go if want_go else None, deterministic)
            return p
        if use16:
            # heads GEMM + likelihood + decoder gradient in one kernel; da stays fp16
            assert R == 1 and not deterministic
            self._plan_backward(p)
            self._plan_fused(p, M)
            if not mid:
                K.f32_to_f16(d, l.in_p, p.D16)
            if not self.enc:
                self._refresh_shadows(p, M)
            p.fused_scale = 2.0 ** round(math.log2(max(S * B, 16) / 16.0))
            dd = p.d_decH[-1] if self.dec else p.dZ
            t16 = p.X16 if p.t16_is_x16 else p.T16
            K.heads_fused_bwd(self.kind, p.D16, p.W16, self.Gh, t16, M, self.G, p.dA16, dd,
                              l.n_in, p.logp, p.fused_ws, row_const=rc, go=None,
                              go_scalar=-1.0 / (S * B), scale=p.fused_scale)
            if mid:
                p.mid_done = True      # the bound comes out of vae_mid_bwd (backward())
            else:
                self._bound(p, R, S, weight)
            p.fused_done = True
            return p
        if drop and self.keep_h:
            for r0, nr, site in self._head_blocks():
                st = self._drop_site(p, site, d, M, l.n_in + l.n_extra, l.n_in, self.keep_h)
                K.gemm(K.GEMM_NT, M, nr, l.k_in, st.copy, l.w[r0:r0 + nr], p.A[:M, r0:r0 + nr],
                       tensor_cores=False)
        else:
            self._gemm(p, K.GEMM_NT, M, l.n_out, l.k_in, d, l.w, p.A[:M])
        if fused_backward:
            assert R == 1 and not deterministic
            self._plan_backward(p)
            self._likelihood(p, tgt, p.A[:M], M, rc, logp=p.logp, da=p.dA[:M],
                             go_scalar=-1.0 / (S * B))
            self._bound(p, R, S, weight)
        else:
            self._likelihood(p, tgt, p.A[:M], M, rc, logp=p.logp)
            self._bound(p, R, S, weight, p.go if want_go else None, deterministic)
        return p

    # ---- dropout (MU:45-50): one mask per site, the dropped copy feeds the site's product ------
    def _drop_site(self, p, site, src, rows, n, skip_col, keep):
        """Draw (unless injected) the site's mask and write the dropped copy of ``src``."""
        st = p.drop.get(site)
        if st is None:
            import statistics
            st = p.drop[site] = type("DropSite", (), {})()
            st.index = len(p.drop)
            st.rows, st.n, st.skip, st.keep = rows, n, skip_col, keep
            st.threshold = statistics.NormalDist().inv_cdf(keep)
            # independent streams per site and, data-parallel, per rank (the shards differ)
            rank = torch.distributed.get_rank() if (
                torch.distributed.is_available() and torch.distributed.is_initialized()) else 0
            st.seed = self.dropout_seed + 7919 * st.index + 15485863 * rank
            st.noise = torch.zeros(rows, n, dtype=torch.float32, device=self.device)
            st.copy = torch.zeros(rows, src.shape[1], dtype=torch.float32, device=self.device)
        if p.drop_injected:
            # keep (1) -> far below, drop (0) -> far above any threshold
            st.noise.copy_((0.5 - p.drop_masks[site].to(self.device, torch.float32)) * 2e9)
        else:
            K.fill_normal(st.noise, st.seed, 0, self.store.step)
        K.dropout_fwd(src, rows, n, skip_col, st.noise, st.threshold, keep, st.copy,
                      src.shape[1])
        return st

    def inject_dropout_masks(self, p, masks):
        """Parity tests: 0/1 keep masks by site name (rows x masked columns) instead of draws."""
        p.drop_injected = masks is not None
        p.drop_masks = masks

    def _drop_bwd(self, st, dx, dsrc=None, accumulate=False):
        K.dropout_bwd(dx, st.rows, st.n, st.skip, st.noise, st.threshold, st.keep, dsrc=dsrc,
                      accumulate=accumulate)

    def _head_blocks(self):
        """(first row, rows, site) of every likelihood head in the stacked head weight: each
        head is a dense layer of its own in the reference, with its own dropout mask."""
        blocks = [(h * self.Gn, self.Gn, "X_TILDE/" + head.upper())
                  for h, head in enumerate(self.heads)]
        if self.k_max:
            blocks.append((self.P * self.Gn, (self.k_max + 1) * self.Gn, "X_TILDE/P_K"))
        return blocks

    def _drop_tmp(self, p, rows, cols):
        key = "_drop_tmp_{}".format(rows)
        buf = getattr(p, key, None)
        if buf is None or buf.shape[1] < cols:
            buf = torch.zeros(rows, round4(cols), dtype=torch.float32, device=self.device)
            setattr(p, key, buf)
        return buf

    def _bound(self, p, R, S, weight, go=None, deterministic=False):
        """ELBO terms (+ d loss / d log p) from the per-row log-likelihoods and the KL term."""
        if deterministic:
            R = S = 1
        if self.sampled_kl:
            K.gaussian_sampled_kl(p.PH, p.B, self.L, R * S, p.eps, p.kl_rows, p.kl_elem,
                                  unit_variance=self.unit_variance, deterministic=deterministic)
            K.vae_bound_rows(p.logp, p.kl_rows, R, S, p.B, weight, p.bound, go)
        else:
            K.vae_bound(p.logp, p.kl_row, R, S, p.B, weight, p.bound, go)

    def set_batch_count_sum_parameter(self, p, count_sum):
        """N of the constrained Poisson for the current minibatch ([B] raw count sums)."""
        p.count_sum_parameter.copy_(count_sum.reshape(-1))

    def _likelihood(self, p, tgt, A, M, rc, logp=None, da=None, go=None, go_scalar=1.0):
        """log p (+ gradient when ``da`` is given) of the M head rows in A (stand-alone kernels)."""
        if self.k_max:
            K.piecewise_likelihood(self.kind, self.k_max, tgt, A, self.Gn, M, self.G, logp=logp,
                                   go=go, go_scalar=go_scalar, da=da)
        elif self.constrained:
            K.constrained_poisson(tgt, A, M, self.G, p.count_sum_parameter, logp=logp, row_const=rc,
                                  go=go, go_scalar=go_scalar, da=da, lse=p.lse)
        elif self.continuous:
            K.continuous_likelihood(self.kind, tgt, A, self.Gn, M, self.G, logp=logp, go=go,
                                    go_scalar=go_scalar, da=da)
        elif da is not None:
            K.likelihood_bwd(self.kind, tgt, A, self.Gn, M, self.G, da, logp=logp, row_const=rc,
                             go=go, go_scalar=go_scalar)
        else:
            K.likelihood_fwd(self.kind, tgt, A, self.Gn, M, self.G, logp, row_const=rc)

    def set_batch_features(self, p, batch_index=None, count_sum=None):
        """Decoder-input extras of the current minibatch: batch ids ([B], any numeric dtype)
        and/or normalised count sums ([B])."""
        if self.number_of_batches:
            p.batch_index.copy_(batch_index.reshape(-1))
        if self.count_sum_feature:
            p.count_sum.copy_(count_sum.reshape(-1))

    def _decoder_features(self, p, M):
        if self.n_extra:
            K.decoder_features(p.Z, M, p.B, self.L + 1, p.batch_index, self.number_of_batches,
                               p.count_sum)

    def decode(self, p, rows):
        """Decoder-only entry point (``session.run(p_x_mean, feed_dict={z: ...})``, VAE:1706-1721):
        p.Z[:rows, :L] holds the latent samples; runs decoder + heads with moving statistics."""
        K.act_fwd(p.Z[:rows], self.L, p.Z[:rows], relu=False)   # (re)write the augmented columns
        if self.n_extra:
            raise NotImplementedError("decoding free latent samples with batch correction or "
                                      "count-sum features (as in the reference, VAE:1639-1649)")
        d = p.Z
        for j, l in enumerate(self.dec):
            self._gemm(p, K.GEMM_NT, rows, l.n_out, l.k_in, d, l.w, p.decY[j][:rows])
            if l.bn:
                K.bn_act_fwd(p.decY[j][:rows], l.n_out, l.beta, l.moving_mean, l.moving_var,
                             p.decH[j][:rows], p.dec_mean[j], p.dec_rstd[j], p.bn_scratch,
                             training=False, update_moving=False, relu=True)
            else:
                K.act_fwd(p.decY[j][:rows], l.n_out, p.decH[j][:rows], relu=True)
            d = p.decH[j]
        l = self.head
        self._gemm(p, K.GEMM_NT, rows, l.n_out, l.k_in, d, l.w, p.A[:rows])

    # ------------------------------------------------------------------ backward -----------
    def backward(self, p, R, S, warm_up_weight=1.0, dA_ready=False, defer_join=False):
        """Gradients of -lower_bound_weighted w.r.t. every parameter into the flat grad buffer."""
        self._plan_backward(p)
        B, M = p.B, p.M
        tgt = p.T if getattr(p, "use_T", False) else p.X
        rc = p.row_const if p.have_row_const else None
        if not dA_ready:
            self._likelihood(p, tgt, p.A, M, rc, logp=None, da=p.dA, go=p.go)
        l = self.head
        d_in = p.decH[-1] if self.dec else p.Z
        dd_in = p.d_decH[-1] if self.dec else p.dZ
        p.head_join = None
        mid = getattr(p, "mid_done", False)

        def mid_backward():
            # decoder -> sample / KL -> posterior -> encoder backward, log p finish and the bound in
            # ONE persistent kernel
            md = self._mid_desc(p, backward=True)
            dd = p.d_decH[-1]
            md.dd_parts, md.dd_ld, md.dd_slice, md.dd_nsplit = dd.data_ptr(), dd.stride(0), 0, 1
            md.logp_parts, md.logp_slice, md.logp_nsplit = p.logp.data_ptr(), 0, 1
            md.row_const = None
            md.go_scalar = -1.0 / (S * B)
            md.dy1_scale = p.fused_scale
            md.deterministic = 0
            self.set_step_scalars(None, warm_up_weight)
            K.vae_mid_bwd(md)

        # The kernel's grid barriers need all of its CTAs resident.  Beside the persistent head
        # weight-gradient GEMMs of the side stream that holds when the two share the SMs by
        # construction (GEMM CTAs capped at #SMs - mid CTAs, one CTA per SM each); otherwise the
        # kernel runs first.
        share = bool(mid and p.fused_done and self.overlap_streams and self._mid_bwd_grid(p) > 0)
        side_ctas = self._sm_count() - self._mid_bwd_grid(p) if share else self.side_gemm_ctas
        if mid and not share:
            mid_backward()
        if p.fused_done:
            # dd came out of the fused kernel; dW = da^T d from the fp16 da (per head).  These
            # two HBM-bound products (and, data-parallel, the all-reduce of their 2/3 of the
            # gradient buffer) are independent of the rest of the backward chain: they run on
            # the side stream beside the latency-bound decoder / encoder backward.
            main = torch.cuda.current_stream()
            if self.overlap_streams:
                if getattr(p, "head_ev", None) is None:
                    p.head_ev = (torch.cuda.Event(), torch.cuda.Event())
                side = self._side_stream()
                p.head_ev[0].record(main)
                side.wait_event(p.head_ev[0])
                ctx = torch.cuda.stream(side)
            else:
                import contextlib
                ctx = contextlib.nullcontext()
            with ctx:
                # leave some SMs to the small kernels of the main stream
                old = K.gemm_sm_limit(side_ctas) if self.overlap_streams else 0
                for h in range(self.P):
                    K.gemm_f16(K.GEMM_TN, self.Gn, l.in_p, M,
                               p.dA16[:, h * self.Gh:h * self.Gh + self.Gn], p.D16,
                               l.dw[h * self.Gn:(h + 1) * self.Gn], alpha=1.0 / p.fused_scale)
                if self.overlap_streams:
                    K.gemm_sm_limit(old)
                if self.overlap_streams:
                    if self._all_reduce is not None and self._peer is None:
                        off = self.store.offsets[l.name + "/W"][0]
                        self._all_reduce(self.store.grad[off:])
                        p.head_reduced_from = off
                    p.head_ev[1].record(side)
                    p.head_join = p.head_ev[1]
        elif p.drop_on and self.keep_h:
            # per head: dW from the head's own dropped operand; the input gradients of the
            # heads meet in dd through their masks
            tmp = self._drop_tmp(p, M, l.in_p)
            for b, (r0, nr, site) in enumerate(self._head_blocks()):
                st = p.drop[site]
                K.gemm(K.GEMM_TN, nr, l.in_p, M, p.dA[:, r0:r0 + nr], st.copy, l.dw[r0:r0 + nr],
                       tensor_cores=False)
                K.gemm(K.GEMM_NN, M, l.n_in, nr, p.dA[:, r0:r0 + nr], l.w[r0:r0 + nr], tmp,
                       tensor_cores=False)
                self._drop_bwd(st, dd_in, dsrc=tmp, accumulate=b > 0)
        else:
            # wgrad (bias gradient = the augmented ones column) and dgrad of the heads
            self._gemm(p, K.GEMM_TN, l.n_out, l.in_p, M, p.dA, d_in, l.dw)
            self._gemm(p, K.GEMM_NN, M, l.n_in, l.n_out, p.dA, l.w, dd_in)
        if mid:
            if share:
                mid_backward()
            # the first layer's weight gradient on the tensor cores
            l = self.enc[0]
            self._wgrad1_16(p, l, B, single=not self.dy1_split)
            self._finish_backward(p, defer_join)
            return
        for j in range(len(self.dec) - 1, -1, -1):
            l = self.dec[j]
            if l.bn:
                K.bn_act_bwd(p.d_decH[j], p.decY[j], p.decH[j], l.n_out, p.dec_mean[j],
                             p.dec_rstd[j], p.d_decY[j], l.dbeta, p.bn_scratch, relu=True)
            else:
                K.act_bwd(p.d_decH[j], p.decH[j], l.n_out, p.d_decY[j], relu=True)
            d_in = p.decH[j - 1] if j > 0 else p.Z
            dd_in = p.d_decH[j - 1] if j > 0 else p.dZ
            st = p.drop.get(l.name) if p.drop_on else None
            if st is not None:
                d_in = st.copy
            self._gemm(p, K.GEMM_TN, l.n_out, l.in_p, M, p.d_decY[j], d_in, l.dw)
            self._gemm(p, K.GEMM_NN, M, l.n_in, l.n_out, p.d_decY[j], l.w, dd_in)
            if st is not None:
                self._drop_bwd(st, dd_in)       # gradient w.r.t. the un-dropped input, in place
        if self.sampled_kl:
            weight = warm_up_weight * self.kl_weight
            # R == 1: d loss / d log p = -1/(S B) for every row, so d loss / d KL = weight/(S B)
            K.gaussian_sampled_kl_bwd(p.PH, B, self.L, p.RS, p.eps, p.dZ,
                                      p.go if R > 1 else None, weight, weight / (S * B), p.dPH,
                                      unit_variance=self.unit_variance)
        else:
            kl_coef = warm_up_weight * self.kl_weight / B
            K.gaussian_latent_bwd(p.PH, B, self.L, p.RS, p.eps, p.dZ, kl_coef, p.dPH,
                                  unit_variance=self.unit_variance)
        l = self.post
        h_in = p.encH[-1] if self.enc else p.X
        if p.drop_on and self.keep_h:
            L = self.L
            tmp = self._drop_tmp(p, B, l.in_p) if self.enc else None
            for part, site in enumerate(["POSTERIOR/MU"] + (
                    [] if self.unit_variance else ["POSTERIOR/LOG_SIGMA"])):
                st = p.drop[site]
                rows, cols = slice(part * L, (part + 1) * L), slice(part * L, (part + 1) * L)
                K.gemm(K.GEMM_TN, L, l.in_p, B, p.dPH[:, cols], st.copy, l.dw[rows],
                       tensor_cores=False)
                if self.enc:
                    K.gemm(K.GEMM_NN, B, l.n_in, L, p.dPH[:, cols], l.w[rows], tmp,
                           tensor_cores=False)
                    self._drop_bwd(st, p.d_encH[-1], dsrc=tmp, accumulate=part > 0)
        else:
            self._gemm(p, K.GEMM_TN, l.n_out, l.in_p, B, p.dPH, h_in, l.dw)
            if self.enc:
                self._gemm(p, K.GEMM_NN, B, l.n_in, l.n_out, p.dPH, l.w, p.d_encH[-1])
        for i in range(len(self.enc) - 1, -1, -1):
            l = self.enc[i]
            if l.bn:
                K.bn_act_bwd(p.d_encH[i], p.encY[i], p.encH[i], l.n_out, p.enc_mean[i],
                             p.enc_rstd[i], p.d_encY[i], l.dbeta, p.bn_scratch, relu=True)
            else:
                K.act_bwd(p.d_encH[i], p.encH[i], l.n_out, p.d_encY[i], relu=True)
            h_in = p.encH[i - 1] if i > 0 else p.X
            st = p.drop.get(l.name) if p.drop_on else None
            if st is not None:
                h_in = st.copy
            if i == 0 and p.fused_done:
                # dW1 = dY1^T X with the fp16 minibatch (dY1 scaled into fp16 range)
                self._dy1_16(p)
                K.f32_to_f16_split(p.d_encY[0], l.n_out, p.dY1_16, p.dY1_16_lo, scale=p.fused_scale)
                self._wgrad1_16(p, l, B, single=False)
            else:
                self._gemm(p, K.GEMM_TN, l.n_out, l.in_p, B, p.d_encY[i], h_in, l.dw)
            if i > 0:
                self._gemm(p, K.GEMM_NN, B, l.n_in, l.n_out, p.d_encY[i], l.w, p.d_encH[i - 1])
                if st is not None:
                    self._drop_bwd(st, p.d_encH[i - 1])
        self._finish_backward(p, defer_join)

    def set_step_scalars(self, learning_rate=None, warm_up_weight=None):
        """Update {learning rate, warm-up weight} on the device when they changed (outside a
        captured step: TrainLoop calls this before the replay)."""
        lr0, w0 = self._scalars_host
        lr = lr0 if learning_rate is None else float(learning_rate)
        w = w0 if warm_up_weight is None else float(warm_up_weight)
        if (lr, w) != (lr0, w0):
            if self.device.type == "cuda" and torch.cuda.is_current_stream_capturing():
                raise RuntimeError("step scalars changed inside a captured step: set them with "
                                   "set_step_scalars() before the capture / replay")
            self.scalars.copy_(torch.tensor([0.0 if lr is None else lr, 1.0 if w is None else w],
                                            dtype=torch.float32))
            self._scalars_host = (lr, w)

    def _finish_backward(self, p, defer_join):
        self._reduce_upto = None
        self._side_tail = None
        if p.head_join is not None:
            off = self.store.offsets[self.head.name + "/W"][0]
            if self._all_reduce is not None and self._peer is None:
                self._reduce_upto = off                   # the tail is already summed (side stream)
            if defer_join:
                self._side_tail = (p, off)                # optimiser_step finishes the tail there
            else:
                torch.cuda.current_stream().wait_event(p.head_join)

    def _adam(self, lo, hi, advance=None):
        """clip + Adam on the flat range; the learning rate comes from the device scalars.  On one
        GPU the kernel also rewrites the fp16 weight shadows and (advance = (counter, total CTAs))
        advances the step counter once the last optimiser CTA of the step is done."""
        s = self.store
        shadows = self._adam_shadows(lo, hi) if self.world_size == 1 else None
        K.adam_clip_step(s.param[lo:hi], s.grad[lo:hi], s.m[lo:hi], s.v[lo:hi], s.step,
                         1.0, ADAM_BETA1, ADAM_BETA2, ADAM_EPSILON, GRADIENT_CLIP,
                         1.0 / self.world_size, scalars=self.scalars, shadows=shadows,
                         advance_counter=advance[0] if advance else None,
                         advance_total=advance[1] if advance else 0)

    def optimiser_step(self, learning_rate):
        """[all-reduce] -> clip to [-1, 1] -> TF Adam (VAE:2742-2759).  One fused launch over the
        flat buffers; when the head gradients were produced on the side stream (train_step) the
        head slice (2/3 of the parameters) is updated there, beside the rest of the backward."""
        s = self.store
        self.set_step_scalars(learning_rate, None)
        tail = getattr(self, "_side_tail", None)
        self._side_tail = None
        upto = getattr(self, "_reduce_upto", None)
        self._reduce_upto = None
        peer = self._peer if self.world_size > 1 else None
        hi = s.total if tail is None else tail[1]
        # one GPU: the optimiser launches advance the step counter themselves (last CTA done)
        advance = None
        if peer is None and self._all_reduce is None:
            total = K.adam_clip_ctas(hi) + (K.adam_clip_ctas(s.total - hi) if tail is not None else 0)
            advance = (self._adam_counter, total)
        if tail is not None:
            p, off = tail
            side = self._side_stream()
            with torch.cuda.stream(side):
                if peer is not None:        # exchange + Adam of the head slice in one kernel
                    peer.reduce_adam(off, s.total, 1.0, channel=1,
                                     max_ctas=self.dp_side_ctas, scalars=self.scalars)
                else:
                    self._adam(off, s.total, advance)
                p.head_ev[1].record(side)
        self._last_ranges = [(0, hi)] + ([(hi, s.total)] if tail is not None else [])
        if peer is not None:
            peer.reduce_adam(0, hi, 1.0, channel=0, scalars=self.scalars)
        else:
            if self._all_reduce is not None:
                self._all_reduce(s.grad if upto is None else s.grad[:upto])
            self._adam(0, hi, advance)
        if tail is not None:
            torch.cuda.current_stream().wait_event(tail[0].head_ev[1])
        if advance is None:
            K.step_advance(s.step)
        # the optimiser kernel rewrote the fp16 weight shadows only on the single-GPU path
        if getattr(self, "W16", None) is not None:
            self._shadow_valid = self.world_size == 1

    def train_step(self, p, R, S, learning_rate, warm_up_weight=1.0):
        """One ``session.run([optimiser, lower_bound])`` (VAE:1026-1029)."""
        if R == 1:
            self.forward(p, True, R, S, warm_up_weight, fused_backward=True)
            self.backward(p, R, S, warm_up_weight, dA_ready=True, defer_join=True)
        else:
            self.forward(p, True, R, S, warm_up_weight, want_go=True)
            self.backward(p, R, S, warm_up_weight)
        self.optimiser_step(learning_rate)
        return p.bound

    # ------------------------------------------------------------------ noise --------------
    def sample_noise(self, p, seed, offset):
        K.fill_normal(p.eps, seed, offset)

    def set_data_parallel(self, world_size, all_reduce):
        """Data-parallel training: ``all_reduce(flat_grad)`` must sum in place over ranks."""
        self.world_size = int(world_size)
        self._all_reduce = all_reduce if world_size > 1 else None

    # ------------------------------------------------------------------ evaluate extras ----
    def _moments_launch(self, p, RS, mean, stddev, stddev_of_mean):
        """One launch of the moments kernel of this likelihood; any output may be None (skipped),
        the given ones share a row pitch."""
        outs = (mean, stddev, stddev_of_mean)
        if self.k_max:
            K.piecewise_moments(self.kind, self.k_max, p.A, self.Gn, p.B, self.G, RS, *outs)
        elif self.constrained:
            K.constrained_poisson_moments(p.A, p.lse, p.count_sum_parameter, p.B, self.G, RS, *outs)
        elif self.continuous:
            K.continuous_moments(self.kind, p.A, self.Gn, p.B, self.G, RS, 1, None, *outs)
        else:
            K.likelihood_moments(self.kind, p.A, self.Gn, p.B, self.G, RS, 1, None, *outs)

    def _moment_buffers(self, p):
        """Three (B, Gn) result buffers kept with the plan (evaluation passes reuse them)."""
        bufs = getattr(p, "moment_bufs", None)
        if bufs is None:
            bufs = p.moment_bufs = [torch.empty(p.B, self.Gn, dtype=torch.float32, device=self.device)
                                    for _ in range(3)]
        return bufs

    def moments(self, p, R, S, deterministic=False, mean_out=None, want_stddev=True):
        """p_x_mean, p_x_stddev, stddev_of_p_x_given_z_mean (VAE:2665-2713) for the batch.
        ``mean_out`` (B, >= G; any row pitch): p_x_mean goes there (the staging buffer of a
        ``hotloop.ReconstructionSink``) instead of the plan's buffer; ``want_stddev=False`` skips
        the two deviation outputs (returned as None)."""
        RS = 1 if deterministic else R * S
        bufs = self._moment_buffers(p)
        if mean_out is None:
            self._moments_launch(p, RS, bufs[0], bufs[1] if want_stddev else None,
                                 bufs[2] if want_stddev else None)
            mean = bufs[0]
        else:
            self._moments_launch(p, RS, mean_out, None, None)
            if want_stddev:
                self._moments_launch(p, RS, None, bufs[1], bufs[2])
            mean = mean_out
        return [mean[:, :self.G], bufs[1][:, :self.G] if want_stddev else None,
                bufs[2][:, :self.G] if want_stddev else None]

    def kl_neurons(self, p):
        K.col_mean(p.kl_elem, p.B, self.L, p.kl_neurons)
        return p.kl_neurons
