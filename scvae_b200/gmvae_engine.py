"""Step engine of the Gaussian-mixture VAE: the graph of
``GaussianMixtureVariationalAutoencoder._setup_model_graph/_setup_loss_function``
(scvae/models/gaussian_mixture_variational_autoencoder.py:2788-3434) as a fixed launch sequence.

Batching of the K cluster passes (the GMVAE hot loop): all K passes of
q(z|x,y=k) and p(x|z_k) share weights, so they run as K consecutive row groups of one tall
matrix (rows ordered (k, sample, cell)) through the same GEMM / batch-norm (groups = K) /
likelihood kernels as the VAE.  ``x W_x`` of the first q(z|x,y) layer is computed once and the
one-hot part of the concat [x, e_k] becomes a per-group row offset.  The decoder + heads +
likelihood + their backward are processed in chunks of clusters so that the (rows x P*genes)
head buffers stay bounded; weight gradients accumulate across chunks inside the GEMM epilogue
(TMA reduce-add).
"""

import math
from collections import OrderedDict

import torch

from . import kernels as K
from .engine import (ADAM_BETA1, ADAM_BETA2, ADAM_EPSILON, GRADIENT_CLIP, ParameterStore, VAEEngine,
                     _Layer, aug, round4)


class GMVAEEngine(VAEEngine):
    PK_SCOPE = "X/DISTRIBUTION/P_K"

    def __init__(self, feature_size, latent_size, number_of_latent_clusters, hidden_sizes=(100,),
                 reconstruction_distribution="poisson", minibatch_normalisation=True, kl_weight=1.0,
                 prior_probabilities_method="uniform", prior_probabilities=None,
                 proportion_of_free_nats_for_y_kl_divergence=0.0, device="cuda", seed=0,
                 tensor_cores=True, head_buffer_bytes=4 << 30, number_of_batches=0,
                 count_sum_feature=False, number_of_reconstruction_classes=0,
                 dropout_keep_probabilities=None, latent_distribution="gaussian mixture"):
        if reconstruction_distribution not in K.LIKELIHOOD_KINDS:
            raise ValueError("reconstruction distribution `{}` is not supported by the "
                             "B200 hot path".format(reconstruction_distribution))
        if prior_probabilities_method not in ("uniform", "learn", "custom"):
            raise ValueError("unknown prior probabilities method `{}`".format(
                prior_probabilities_method))
        self.G, self.L, self.K = int(feature_size), int(latent_size), int(number_of_latent_clusters)
        self.hidden_sizes = [int(h) for h in hidden_sizes]
        if not self.hidden_sizes:
            raise ValueError("the GMVAE engine needs at least one hidden layer")
        self.kind_name = reconstruction_distribution
        self.kind = K.LIKELIHOOD_KINDS[reconstruction_distribution]
        self.heads = K.LIKELIHOOD_HEADS[reconstruction_distribution]
        self.P = len(self.heads)
        self.bn = bool(minibatch_normalisation)
        self.kl_weight = float(kl_weight)
        self.prior_method = prior_probabilities_method
        self.free_nats = float(proportion_of_free_nats_for_y_kl_divergence)
        self.device = torch.device(device)
        self.tensor_cores = bool(tensor_cores)
        # fused heads kernel for the K cluster passes of a training step (rows = clusters x
        # samples x cells tile the B target rows: B % 128 == 0); the encoders read the fp32 minibatch
        self.fused_heads = bool(tensor_cores)
        self._needs_fp32_x = True
        self.Gh = (self.G + 63) & ~63
        self.head_buffer_bytes = int(head_buffer_bytes)
        self.Gn, self.Gp = round4(self.G), aug(self.G)
        self.world_size, self._all_reduce, self._plans = 1, None, {}
        self._side, self.overlap_streams, self._peer = None, False, None   # (VAE-engine-only features)
        self.scalars = torch.tensor([0.0, 1.0], dtype=torch.float32, device=self.device)
        self._adam_counter = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._scalars_host = (None, None)
        self.mid_fused = False
        # dropout keep probabilities [hidden, x, z, y] (GMVAE:276-296; a scalar means hidden only;
        # False / None / 0 / 1 switch a kind off, MU:45-46).  Every build of a shared layer -- one
        # per cluster -- is a dropout op of its own (reuse=True shares variables, not masks); the
        # fourth value drops the one-hot input of the p(z|y) heads (GMVAE:3036-3040).
        keep = dropout_keep_probabilities
        keep = (list(keep) if isinstance(keep, (list, tuple)) else [keep]) + [False] * 4
        self.keep_h, self.keep_x, self.keep_z, self.keep_y = [
            float(k) if (k and k != 1) else None for k in keep[:4]]
        self.dropout_active = any(k is not None for k in (self.keep_h, self.keep_x, self.keep_z,
                                                          self.keep_y))
        self.dropout_seed = int(seed) + 104729
        if self.dropout_active:
            self.fused_heads = False      # every head multiplies its own dropped operand copy
        # decoder-input extras concatenated to every z_k (GMVAE:3097-3132), as in the VAE engine
        self.number_of_batches = int(number_of_batches or 0)
        self.count_sum_feature = bool(count_sum_feature)
        self.n_extra = self.number_of_batches + (1 if self.count_sum_feature else 0)
        self.Zp = round4(int(latent_size) + 1 + self.n_extra)
        self.lfm_inference = self.lfm_generative = False
        # constrained Poisson (rate = N softmax_g(a), N = the cell's count sum fed as
        # count_sum_parameter, GMVAE:419-429, :3170-3176): its own row kernel, never the fused heads
        self.constrained = self.kind == K.CONSTRAINED_POISSON
        self.continuous = self.kind in K.CONTINUOUS_KINDS      # row kernels of csrc/continuous.cu
        # piecewise-categorical likelihood (`-k`, head P_K GMVAE:3192-3218): k_max + 1 class-logit
        # head blocks behind the P heads; its own row kernel, never the fused heads
        self.k_max = int(number_of_reconstruction_classes or 0)
        self.PT = self.P + (self.k_max + 1 if self.k_max else 0)
        if self.k_max and self.continuous:
            raise ValueError("piecewise-categorical likelihoods wrap the Poisson / NB family")
        if self.k_max or self.constrained or self.continuous:
            self.fused_heads = False
        self.unit_variance = False
        # "gaussian mixture": softplus Gaussians, heads [mean | softplus_scale];
        # "full-covariance gaussian mixture": multivariate Gaussians with a lower-triangular scale,
        # heads [locations | L (L + 1) / 2 scales] (DU:75-93, :345-348; csrc/gmvae_full.cu)
        if latent_distribution not in ("gaussian mixture", "full-covariance gaussian mixture"):
            raise ValueError("unknown GMVAE latent distribution `{}`".format(latent_distribution))
        self.full_cov = latent_distribution.startswith("full-covariance")
        self.T = self.L * (self.L + 1) // 2
        self.nL = self.L + self.T if self.full_cov else 2 * self.L      # head pre-activations per row
        if self.full_cov:
            if self.L > 128:
                raise ValueError("the full-covariance mixture is built for latent sizes <= 128")
            if self.keep_h is not None or self.keep_y is not None:
                raise NotImplementedError("hidden / y dropout around the full-covariance mixture heads")
        self.z_scope = "MULTIVARIATE_GAUSSIAN" if self.full_cov else "SOFTPLUS_GAUSSIAN"
        self.z_heads = ("LOCATIONS", "SCALES") if self.full_cov else ("MEAN", "SOFTPLUS_SCALE")
        self.z_head_cols = [slice(0, self.L), slice(self.L, self.nL)]

        def stack(prefix, first_in):
            layers, width = [], first_in
            for i, h in enumerate(self.hidden_sizes):
                layers.append(_Layer("{}/LAYER_{}".format(prefix, i + 1), width, h, self.bn))
                width = h
            return layers, width

        self.qy_enc, width = stack("Y/CATEGORICAL/ENCODER", self.G)
        self.qy_logits = _Layer("Y/CATEGORICAL/LOGITS", width, self.K, False)
        self.qz_enc, width = stack("Z/Q/ENCODER", self.G)
        self.qz_head = _Layer("Z/Q/" + self.z_scope, width, self.nL, False)
        self.dec, width = [], self.L
        for i, h in enumerate(self.hidden_sizes[::-1]):
            self.dec.append(_Layer("X/DECODER/LAYER_{}".format(i + 1), width, h, self.bn,
                                   n_extra=self.n_extra if i == 0 else 0))
            width = h
        self.head = _Layer("X/DISTRIBUTION", width, self.PT * self.Gn, False)
        self.enc = self.qy_enc + self.qz_enc          # every batch-normed encoder layer
        self._dense = self.qy_enc + [self.qy_logits] + self.qz_enc + [self.qz_head] + self.dec + \
            [self.head]

        H1p = round4(self.hidden_sizes[0])
        store = ParameterStore(self.device)
        for layer in self._dense:
            store.add(layer.name + "/W", (layer.n_out, layer.in_p))
            if layer.bn:
                store.add(layer.name + "/beta", (layer.n_out,))
        store.add("Z/Q/WY", (self.K, H1p))            # rows of LAYER_1 weights acting on e_k
        store.add("Z/P/W", (self.K, self.nL))         # p(z|y): [mean | softplus_scale] weights
        store.add("Z/P/B", (1, self.nL))
        if self.prior_method == "learn":
            store.add("Y/P/LOGITS", (self.K,))
        store.allocate()
        self.store = store
        for layer in self._dense:
            layer.w = store.view(store.param, layer.name + "/W")
            layer.dw = store.view(store.grad, layer.name + "/W")
            if layer.bn:
                layer.beta = store.view(store.param, layer.name + "/beta")
                layer.dbeta = store.view(store.grad, layer.name + "/beta")
                layer.moving_mean = torch.zeros(layer.n_out, dtype=torch.float32, device=self.device)
                layer.moving_var = torch.ones(layer.n_out, dtype=torch.float32, device=self.device)
        for key, attr in (("Z/Q/WY", "qz_wy"), ("Z/P/W", "pz_w"), ("Z/P/B", "pz_b")):
            setattr(self, attr, store.view(store.param, key))
            setattr(self, "d_" + attr, store.view(store.grad, key))
        if self.prior_method == "learn":
            self.py_logits = store.view(store.param, "Y/P/LOGITS")
            self.d_py_logits = store.view(store.grad, "Y/P/LOGITS")
        else:
            self.py_logits = self.d_py_logits = None
        if self.prior_method == "custom":
            probs = torch.as_tensor(prior_probabilities, dtype=torch.float64)
            self.log_py = torch.log(probs / probs.sum()).float().to(self.device)
        else:
            self.log_py = torch.full((self.K,), -math.log(self.K), dtype=torch.float32,
                                     device=self.device)
        self.initialise(seed)

    # ------------------------------------------------------------------ parameters ---------
    def _tf_dense(self):
        """(layer, output-row slice, TF scope, input-row slice of the TF weight) in the
        reference's variable creation order (GMVAE:2788-2934)."""
        out = []
        for layer in self.qy_enc:
            out.append((layer, slice(0, layer.n_out), layer.name, slice(0, layer.n_in)))
        out.append((self.qy_logits, slice(0, self.K), self.qy_logits.name, slice(0, self.qy_logits.n_in)))
        for layer in self.qz_enc:
            out.append((layer, slice(0, layer.n_out), layer.name, slice(0, layer.n_in)))
        base = "Z/Q/{}/".format(self.z_scope)
        n_in = self.qz_head.n_in
        for name, cols in zip(self.z_heads, self.z_head_cols):
            out.append((self.qz_head, cols, base + name, slice(0, n_in)))
        return out

    def _tf_tail(self):
        out = [(layer, slice(0, layer.n_out), layer.name, slice(0, layer.n_in)) for layer in self.dec]
        for p, head in enumerate(self.heads):
            out.append((self.head, slice(p * self.Gn, p * self.Gn + self.G),
                        "X/DISTRIBUTION/" + head.upper(), slice(0, self.head.n_in)))
        return out

    def initialise(self, seed=0):
        gen = torch.Generator().manual_seed(int(seed))

        def xavier(fan_in, fan_out):
            limit = math.sqrt(6.0 / (fan_in + fan_out))
            w = torch.rand((fan_in, fan_out), generator=gen, dtype=torch.float64)
            return ((2.0 * w - 1.0) * limit).float()

        params = OrderedDict()
        if self.prior_method == "learn":
            params["Y/P/LOGITS"] = torch.zeros(self.K)
        for layer, rows, scope, _ in self._tf_dense():
            fan_in = layer.n_in + (self.K if scope == "Z/Q/ENCODER/LAYER_1" else 0)
            params[scope + "/DENSE/weights"] = xavier(fan_in, rows.stop - rows.start)
            params[scope + "/DENSE/biases"] = torch.zeros(rows.stop - rows.start)
            if scope == "Z/Q/{}/{}".format(self.z_scope, self.z_heads[1]):
                for name, cols in zip(self.z_heads, self.z_head_cols):
                    width = cols.stop - cols.start
                    params["Z/P/{}/{}/DENSE/weights".format(self.z_scope, name)] = xavier(self.K, width)
                    params["Z/P/{}/{}/DENSE/biases".format(self.z_scope, name)] = torch.zeros(width)
        for layer, rows, scope, _ in self._tf_tail():
            params[scope + "/DENSE/weights"] = xavier(layer.n_in + layer.n_extra, rows.stop - rows.start)
            params[scope + "/DENSE/biases"] = torch.zeros(rows.stop - rows.start)
        if self.k_max:
            fan_out = self.G * (self.k_max + 1)
            params[self.PK_SCOPE + "/DENSE/weights"] = xavier(self.head.n_in, fan_out)
            params[self.PK_SCOPE + "/DENSE/biases"] = torch.zeros(fan_out)
        self.import_parameters(params, strict=False)
        for buf in (self.store.grad, self.store.m, self.store.v):
            buf.zero_()
        self.store.step.zero_()
        for layer in self.bn_layers():
            layer.beta.zero_()
            layer.moving_mean.zero_()
            layer.moving_var.fill_(1.0)

    def import_parameters(self, params, strict=True):
        dev = self.device
        for layer, rows, scope, _ in self._tf_dense() + self._tf_tail():
            w = params[scope + "/DENSE/weights"].to(dev, torch.float32)
            b = params[scope + "/DENSE/biases"].to(dev, torch.float32)
            if scope == "Z/Q/ENCODER/LAYER_1":       # rows [G, G+K) act on the one-hot e_k
                self.qz_wy[:, :layer.n_out] = w[self.G:]
                self.qz_wy[:, layer.n_out:] = 0
                w = w[:self.G]
            layer.w[rows, :layer.n_in] = w[:layer.n_in].t()
            layer.w[rows, layer.n_in] = b
            layer.w[rows, layer.n_in + 1:] = 0
            if layer.n_extra:      # rows [n_in, n_in + extras) of the TF weight: batch one-hot, count sum
                layer.w[rows, layer.n_in + 1:layer.k_in] = w[layer.n_in:].t()
            if layer.bn:
                for key, dst in (("beta", layer.beta), ("moving_mean", layer.moving_mean),
                                 ("moving_variance", layer.moving_var)):
                    name = scope + "/BATCH_NORM/" + key
                    if name in params:
                        dst.copy_(params[name].to(dev, torch.float32))
                    elif strict:
                        raise KeyError(name)
        if self.k_max:      # class-minor P_K variable -> one block of Gn rows per class
            K1, layer = self.k_max + 1, self.head
            w = params[self.PK_SCOPE + "/DENSE/weights"].to(dev, torch.float32)
            b = params[self.PK_SCOPE + "/DENSE/biases"].to(dev, torch.float32)
            for c in range(K1):
                rows = self._pk_rows(c)
                layer.w[rows, :layer.n_in] = w[:, c::K1].t()
                layer.w[rows, layer.n_in] = b[c::K1]
                layer.w[rows, layer.n_in + 1:] = 0
        for name, cols in zip(self.z_heads, self.z_head_cols):
            scope = "Z/P/{}/{}".format(self.z_scope, name)
            self.pz_w[:, cols] = params[scope + "/DENSE/weights"].to(dev, torch.float32)
            self.pz_b[0, cols] = params[scope + "/DENSE/biases"].to(dev, torch.float32)
        if self.prior_method == "learn" and "Y/P/LOGITS" in params:
            self.py_logits.copy_(params["Y/P/LOGITS"].to(dev, torch.float32))

    def _export(self, grads):
        out = OrderedDict()
        pick = (lambda l: l.dw) if grads else (lambda l: l.w)
        if self.prior_method == "learn":
            out["Y/P/LOGITS"] = (self.d_py_logits if grads else self.py_logits).cpu().clone()

        def dense(layer, rows, scope):
            w = self._tf_weight(pick(layer), layer, rows)
            if scope == "Z/Q/ENCODER/LAYER_1":
                wy = (self.d_qz_wy if grads else self.qz_wy)[:, :layer.n_out].cpu()
                w = torch.cat([w, wy], dim=0)
            out[scope + "/DENSE/weights"] = w
            out[scope + "/DENSE/biases"] = pick(layer)[rows, layer.n_in].contiguous().cpu()
            if layer.bn:
                out[scope + "/BATCH_NORM/beta"] = (layer.dbeta if grads else layer.beta).cpu().clone()
                if not grads:
                    out[scope + "/BATCH_NORM/moving_mean"] = layer.moving_mean.cpu().clone()
                    out[scope + "/BATCH_NORM/moving_variance"] = layer.moving_var.cpu().clone()

        for layer, rows, scope, _ in self._tf_dense():
            dense(layer, rows, scope)
        for name, cols in zip(self.z_heads, self.z_head_cols):
            scope = "Z/P/{}/{}".format(self.z_scope, name)
            out[scope + "/DENSE/weights"] = (self.d_pz_w if grads else self.pz_w)[:, cols].cpu().clone()
            out[scope + "/DENSE/biases"] = (self.d_pz_b if grads else self.pz_b)[0, cols].cpu().clone()
        for layer, rows, scope, _ in self._tf_tail():
            dense(layer, rows, scope)
        if self.k_max:
            pk = OrderedDict()
            self._export_pk(pick(self.head), pk)
            out[self.PK_SCOPE + "/DENSE/weights"] = pk["X_TILDE/P_K/DENSE/weights"]
            out[self.PK_SCOPE + "/DENSE/biases"] = pk["X_TILDE/P_K/DENSE/biases"]
        return out

    def export_parameters(self):
        return self._export(False)

    def export_gradients(self):
        return self._export(True)

    # ------------------------------------------------------------------ buffers ------------
    def _plan(self, B, RS):
        key = (B, RS)
        if key in self._plans:
            return self._plans[key]
        dev, f32 = self.device, torch.float32
        Kc, L = self.K, self.L
        KB, M = Kc * B, Kc * RS * B
        # clusters per decoder chunk: bound the two (rows, P*Gn) head buffers
        per_cluster = RS * B * self.PT * self.Gn * 4 * 2
        chunk = max(1, min(Kc, self.head_buffer_bytes // max(per_cluster, 1)))
        p = type("Plan", (), {})()
        p.B, p.RS, p.M, p.KB, p.chunk = B, RS, M, KB, chunk
        Mc = chunk * RS * B

        def zeros(*shape):
            return torch.zeros(*shape, dtype=f32, device=dev)

        p.X = zeros(B, self.Gp)
        p.X[:, self.G] = 1.0
        p.T, p.use_T = None, False
        p.row_const, p.have_row_const = zeros(B), False
        # q(y|x)
        p.qyY = [zeros(B, round4(l.n_out)) for l in self.qy_enc]
        p.qyH = [zeros(B, aug(l.n_out)) for l in self.qy_enc]
        p.qy_mean = [zeros(l.n_out) for l in self.qy_enc]
        p.qy_rstd = [zeros(l.n_out) for l in self.qy_enc]
        p.logits = zeros(B, round4(Kc))
        p.y, p.logy = zeros(B, Kc), zeros(B, Kc)
        # q(z|x,y)
        H1 = self.hidden_sizes[0]
        p.XW = zeros(B, round4(H1))
        p.qzY = [zeros(KB, round4(l.n_out)) for l in self.qz_enc]
        p.qzH = [zeros(KB, aug(l.n_out)) for l in self.qz_enc]
        p.qz_mean = [zeros(Kc * l.n_out) for l in self.qz_enc]
        p.qz_rstd = [zeros(Kc * l.n_out) for l in self.qz_enc]
        p.QH = zeros(KB, round4(self.nL))
        p.PZ = zeros(Kc, self.nL)
        if self.full_cov:      # activated prior scale matrices, S_p^-1 (z - loc_p) of every row
            p.PL = zeros(Kc, L, L)
            p.Wsol = zeros(M, L)
        p.eps = zeros(M, L)
        p.Z = zeros(M, self.Zp)
        p.batch_index = zeros(B) if self.number_of_batches else None
        p.count_sum = zeros(B) if self.count_sum_feature else None
        p.klz = zeros(M)
        p.kl_elem = None
        # constrained Poisson: N of the cells, row log-sum-exps of all K*RS*B rows (each decoder
        # chunk writes its slice through the view p.lse)
        p.count_sum_parameter = zeros(B) if self.constrained else None
        p.lse_all = zeros(M) if self.constrained else None
        p.lse = None
        p.go, p.coef, p.logp = zeros(M), zeros(M), zeros(M)
        # decoder chunk buffers
        p.decY = [zeros(Mc, round4(l.n_out)) for l in self.dec]
        p.decH = [zeros(Mc, aug(l.n_out)) for l in self.dec]
        p.dec_mean = [zeros(chunk * l.n_out) for l in self.dec]
        p.dec_rstd = [zeros(chunk * l.n_out) for l in self.dec]
        p.A = zeros(Mc, self.PT * self.Gn)
        p.bound = zeros(6)
        p.ll_mean, p.klz_mean = zeros(Kc, B), zeros(Kc, B)
        p.z_mean = zeros(B, L)
        scratch = 1
        for l in self.qy_enc:
            scratch = max(scratch, K.bn_scratch_floats(B, l.n_out, 1))
        for l in self.qz_enc:
            scratch = max(scratch, K.bn_scratch_floats(KB, l.n_out, Kc))
        for l in self.dec:
            scratch = max(scratch, K.bn_scratch_floats(Mc, l.n_out, chunk))
        p.bn_scratch = zeros(scratch)
        p.workspace, p.ws_bytes = None, 0
        p.bwd_ready = False
        p.have_t16 = False
        p.have_x = True
        p.t16_is_x16 = False
        p.fused_ready = False
        p.fused_done = False
        p.drop, p.drop_on, p.drop_injected, p.drop_masks = {}, False, False, None
        self._plans[key] = p
        return p

    def _plan_backward(self, p):
        if p.bwd_ready:
            return
        dev, f32 = self.device, torch.float32

        def zeros(*shape):
            return torch.zeros(*shape, dtype=f32, device=dev)

        B, KB, M, Kc, L = p.B, p.KB, p.M, self.K, self.L
        Mc = p.chunk * p.RS * B
        p.dA = zeros(Mc, self.PT * self.Gn)
        p.d_decH = [zeros(Mc, aug(l.n_out)) for l in self.dec]
        p.d_decY = [zeros(Mc, round4(l.n_out)) for l in self.dec]
        p.dZ = zeros(M, self.Zp)
        p.dQH = zeros(KB, round4(self.nL))
        p.dPZ = zeros(Kc, self.nL)
        if self.full_cov:
            p.CU = zeros(M, L)
        p.d_qzH = [zeros(KB, aug(l.n_out)) for l in self.qz_enc]
        p.d_qzY = [zeros(KB, round4(l.n_out)) for l in self.qz_enc]
        p.dXW = zeros(B, round4(self.hidden_sizes[0]))
        p.dlogits = zeros(B, round4(Kc))
        p.dlogits_c = zeros(B, Kc)
        p.d_qyH = [zeros(B, aug(l.n_out)) for l in self.qy_enc]
        p.d_qyY = [zeros(B, round4(l.n_out)) for l in self.qy_enc]
        p.bwd_ready = True

    # ------------------------------------------------------------------ layer helpers -------
    def _bn_fwd(self, p, layer, Y, H, mean, rstd, training, update_moving, groups):
        if layer.bn:
            K.bn_act_fwd(Y, layer.n_out, layer.beta, layer.moving_mean, layer.moving_var, H, mean,
                         rstd, p.bn_scratch, training=training, update_moving=update_moving,
                         relu=True, groups=groups)
        else:
            K.act_fwd(Y, layer.n_out, H, relu=True)

    def _bn_bwd(self, p, layer, dH, Y, H, mean, rstd, dY, groups, accumulate):
        if layer.bn:
            K.bn_act_bwd(dH, Y, H, layer.n_out, mean, rstd, dY, layer.dbeta, p.bn_scratch,
                         relu=True, groups=groups, accumulate_dbeta=accumulate)
        else:
            K.act_bwd(dH, H, layer.n_out, dY, relu=True)

    # ---- dropout (MU:45-50; GMVAE:276-296) -------------------------------------------------------
    # A site covers `total` tall rows ordered (cluster, sample, cell): the K builds of a shared
    # layer are K row groups with independent masks (recorded as `site`, `site#1`, ... in the
    # reference-graph fixtures).  The decoder sites are visited chunk by chunk (rows [r0, r0+rows)).
    def _gdrop(self, p, site, src, r0, rows, total, n, skip_col, keep, groups):
        st = p.drop.get(site)
        if st is None:
            import statistics
            st = p.drop[site] = type("DropSite", (), {})()
            st.index = len(p.drop)
            st.n, st.skip, st.keep, st.total = n, skip_col, keep, total
            st.threshold = statistics.NormalDist().inv_cdf(keep)
            rank = torch.distributed.get_rank() if (
                torch.distributed.is_available() and torch.distributed.is_initialized()) else 0
            st.seed = self.dropout_seed + 7919 * st.index + 15485863 * rank
            st.noise = torch.zeros(total, n, dtype=torch.float32, device=self.device)
            st.copy = torch.zeros(rows, src.shape[1], dtype=torch.float32, device=self.device)
        if r0 == 0:          # once per step: the whole site's masks
            if p.drop_injected:
                per = total // groups
                for k in range(groups):
                    mask = p.drop_masks[site if k == 0 else "{}#{}".format(site, k)]
                    st.noise[k * per:(k + 1) * per].copy_(
                        (0.5 - mask.to(self.device, torch.float32)) * 2e9)
            else:
                K.fill_normal(st.noise, st.seed, 0, self.store.step)
        K.dropout_fwd(src, rows, n, skip_col, st.noise[r0:r0 + rows], st.threshold, keep,
                      st.copy[:rows], src.shape[1])
        st.r0, st.rows = r0, rows
        return st

    def _gdrop_bwd(self, st, dx, dsrc=None, accumulate=False):
        """Gradient through the site's last forward call (rows [st.r0, st.r0 + st.rows))."""
        K.dropout_bwd(dx, st.rows, st.n, st.skip, st.noise[st.r0:st.r0 + st.rows], st.threshold,
                      st.keep, dsrc=dsrc, accumulate=accumulate)

    def inject_dropout_masks(self, p, masks):
        """Parity tests: 0/1 keep masks by site name (`site`, `site#k` per cluster build)."""
        p.drop_injected = masks is not None
        p.drop_masks = masks

    def _drop_tmp(self, p, rows, cols):
        key = "_drop_tmp_{}".format(rows)
        buf = getattr(p, key, None)
        if buf is None or buf.shape[1] < cols:
            buf = torch.zeros(rows, round4(cols), dtype=torch.float32, device=self.device)
            setattr(p, key, buf)
        return buf

    def _head_blocks(self):
        """(first row, rows, site) of every likelihood head in the stacked head weight."""
        blocks = [(h * self.Gn, self.Gn, "X/DISTRIBUTION/" + head.upper())
                  for h, head in enumerate(self.heads)]
        if self.k_max:
            blocks.append((self.P * self.Gn, (self.k_max + 1) * self.Gn, self.PK_SCOPE))
        return blocks

    def _qz_first_layer_dropped(self, p):
        """First q(z|x,y) layer under input dropout (keep_x): every cluster build drops its own
        copy of [x | e_k], so the shared x W_x product does not apply -- the K dropped copies are K
        row groups of one tall operand against [W_x | W_y | b] (GMVAE:2942-2956)."""
        l, Kc, B, G = self.qz_enc[0], self.K, p.B, self.G
        Wc = round4(G + Kc + 1)
        if getattr(p, "XE", None) is None:
            p.XE = torch.zeros(Kc * B, Wc, dtype=torch.float32, device=self.device)
            view = p.XE.view(Kc, B, Wc)
            for k in range(Kc):
                view[k, :, G + k] = 1.0          # e_k
            view[:, :, G + Kc] = 1.0             # bias column
            p.Wcat = torch.zeros(l.n_out, Wc, dtype=torch.float32, device=self.device)
            p.dWcat = torch.zeros(l.n_out, Wc, dtype=torch.float32, device=self.device)
        p.XE.view(Kc, B, Wc)[:, :, :G].copy_(p.X[:, :G])
        p.Wcat[:, :G].copy_(l.w[:, :G])
        p.Wcat[:, G:G + Kc].copy_(self.qz_wy[:, :l.n_out].t())
        p.Wcat[:, G + Kc].copy_(l.w[:, G])
        st = self._gdrop(p, l.name, p.XE, 0, Kc * B, Kc * B, G + Kc, G + Kc, self.keep_x, Kc)
        self._gemm(p, K.GEMM_NT, Kc * B, l.n_out, G + Kc + 1, st.copy, p.Wcat, p.qzY[0])

    def _qz_first_layer_dropped_bwd(self, p):
        l, Kc, B, G = self.qz_enc[0], self.K, p.B, self.G
        st = p.drop[l.name]
        self._gemm(p, K.GEMM_TN, l.n_out, p.Wcat.shape[1], Kc * B, p.d_qzY[0], st.copy, p.dWcat)
        l.dw.zero_()
        l.dw[:, :G].copy_(p.dWcat[:, :G])
        l.dw[:, G].copy_(p.dWcat[:, G + Kc])
        self.d_qz_wy.zero_()
        self.d_qz_wy[:, :l.n_out].copy_(p.dWcat[:, G:G + Kc].t())

    def _pz_scales(self, p):
        """p(z|y) under dropout of its one-hot input (keep_y): mean / scale rows of cluster k are
        b + (mask_kk / keep) W[k], one mask per head and cluster (GMVAE:3009-3048)."""
        Kc, L = self.K, self.L
        if getattr(p, "eye", None) is None:
            p.eye = torch.zeros(Kc, round4(Kc), dtype=torch.float32, device=self.device)
            p.eye[:, :Kc] = torch.eye(Kc, device=self.device)
            p.pz_scale = torch.ones(Kc, 2 * L, dtype=torch.float32, device=self.device)
            p.pz_wd = torch.zeros_like(self.pz_w)
        for part, name in enumerate(("MEAN", "SOFTPLUS_SCALE")):
            st = self._gdrop(p, "Z/P/SOFTPLUS_GAUSSIAN/" + name, p.eye, 0, Kc, Kc, Kc, Kc,
                             self.keep_y, Kc)
            p.pz_scale[:, part * L:(part + 1) * L] = torch.diagonal(st.copy[:, :Kc]).reshape(Kc, 1)
        torch.mul(self.pz_w, p.pz_scale, out=p.pz_wd)
        return p.pz_wd

    # ---- gene-axis products of the two encoders --------------------------------------------------
    # With the 16-bit minibatch at hand (integer counts) they run as fp16 products with the weight
    # (forward) resp. the output gradient (weight gradient) split into fp16 + rounding remainder
    # (scvae_gemm_f16_split): exact counts x ~22-bit weights, half the HBM traffic of the tf32
    # product and none of its truncation bias.
    def _x_product(self, p, l, out):
        B = p.B
        if self.tensor_cores and p.have_t16:
            key = "_w16_" + l.name
            bufs = getattr(p, key, None)
            if bufs is None:
                hi = torch.zeros(l.n_out, (self.G + 8) & ~7, dtype=torch.float16, device=self.device)
                bufs = (hi, torch.zeros_like(hi))
                setattr(p, key, bufs)
            K.f32_to_f16_split(l.w, l.n_in + 1, bufs[0], bufs[1])
            self._gemm16_split(p, K.GEMM_NT, B, l.n_out, l.n_in + 1, p.X16, bufs[0], bufs[1], 2, out)
        else:
            self._gemm(p, K.GEMM_NT, B, l.n_out, l.n_in + 1, p.X, l.w, out)

    def _x_wgrad(self, p, l, dY):
        B = p.B
        if self.tensor_cores and p.have_t16:
            key = "_dy16_" + l.name
            bufs = getattr(p, key, None)
            if bufs is None:
                hi = torch.zeros(B, (l.n_out + 7) & ~7, dtype=torch.float16, device=self.device)
                bufs = (hi, torch.zeros_like(hi))
                setattr(p, key, bufs)
            scale = 2.0 ** round(math.log2(max(p.RS * B, 16) / 16.0))
            K.f32_to_f16_split(dY, l.n_out, bufs[0], bufs[1], scale=scale)
            self._gemm16_split(p, K.GEMM_TN, l.n_out, l.in_p, B, bufs[0], p.X16, bufs[1], 1, l.dw,
                               alpha=1.0 / scale)
        else:
            self._gemm(p, K.GEMM_TN, l.n_out, l.in_p, B, dY, p.X, l.dw)

    # ------------------------------------------------------------------ forward ------------
    def forward(self, p, is_training, R, S, warm_up_weight=1.0, update_moving=None,
                with_backward=False, deterministic=False):
        """Full pass; with ``with_backward`` the decoder chunks also run their backward (the
        gradient of each row's log-likelihood, -y_bk/(B RS), is known before the decoder runs)."""
        if deterministic:
            raise NotImplementedError("the GMVAE graph has no deterministic-z mode (as in the "
                                      "reference)")
        B, Kc, L, RS = p.B, self.K, self.L, p.RS
        KB = p.KB
        if update_moving is None:
            update_moving = is_training
        weight = warm_up_weight * self.kl_weight
        if with_backward:
            self._plan_backward(p)
        # --- q(y|x) ---------------------------------------------------------------------------
        drop = p.drop_on = bool(self.dropout_active and is_training)
        h = p.X
        for i, l in enumerate(self.qy_enc):
            keep = (self.keep_x if i == 0 else self.keep_h) if drop else None
            if keep:
                st = self._gdrop(p, l.name, h, 0, B, B, l.n_in, l.n_in, keep, 1)
                self._gemm(p, K.GEMM_NT, B, l.n_out, l.n_in + 1, st.copy, l.w, p.qyY[i])
            elif i == 0:
                self._x_product(p, l, p.qyY[i])
            else:
                self._gemm(p, K.GEMM_NT, B, l.n_out, l.n_in + 1, h, l.w, p.qyY[i])
            self._bn_fwd(p, l, p.qyY[i], p.qyH[i], p.qy_mean[i], p.qy_rstd[i], is_training,
                         update_moving, 1)
            h = p.qyH[i]
        l = self.qy_logits
        if drop and self.keep_h:
            h = self._gdrop(p, l.name, h, 0, B, B, l.n_in, l.n_in, self.keep_h, 1).copy
        self._gemm(p, K.GEMM_NT, B, Kc, l.n_in + 1, h, l.w, p.logits)
        K.softmax_fwd(p.logits, B, Kc, p.y, p.logy)
        if self.prior_method == "learn":
            self._refresh_log_py()
        # --- q(z|x, y=k) for all k ----------------------------------------------------------
        l = self.qz_enc[0]
        if drop and self.keep_x:
            self._qz_first_layer_dropped(p)
        else:
            self._x_product(p, l, p.XW)
            K.group_offset_fwd(p.XW, self.qz_wy, Kc, B, l.n_out, p.qzY[0])
        self._bn_fwd(p, l, p.qzY[0], p.qzH[0], p.qz_mean[0], p.qz_rstd[0], is_training,
                     update_moving, Kc)
        for i in range(1, len(self.qz_enc)):
            l = self.qz_enc[i]
            h = p.qzH[i - 1]
            if drop and self.keep_h:
                h = self._gdrop(p, l.name, h, 0, KB, KB, l.n_in, l.n_in, self.keep_h, Kc).copy
            self._gemm(p, K.GEMM_NT, KB, l.n_out, l.n_in + 1, h, l.w, p.qzY[i])
            self._bn_fwd(p, l, p.qzY[i], p.qzH[i], p.qz_mean[i], p.qz_rstd[i], is_training,
                         update_moving, Kc)
        l = self.qz_head
        if drop and self.keep_h:
            # one mask per posterior parameter (and cluster build) over the same activation
            for part, name in enumerate(("MEAN", "SOFTPLUS_SCALE")):
                st = self._gdrop(p, l.name + "/" + name, p.qzH[-1], 0, KB, KB, l.n_in, l.n_in,
                                 self.keep_h, Kc)
                K.gemm(K.GEMM_NT, KB, L, l.n_in + 1, st.copy, l.w[part * L:(part + 1) * L],
                       p.QH[:, part * L:(part + 1) * L], tensor_cores=False)
        else:
            self._gemm(p, K.GEMM_NT, KB, l.n_out, l.n_in + 1, p.qzH[-1], l.w, p.QH)
        # --- p(z|y=k) = FC(e_k) ---------------------------------------------------------------
        pz_w = self._pz_scales(p) if (drop and self.keep_y) else self.pz_w
        K.group_offset_fwd(self.pz_b, pz_w, Kc, 1, self.nL, p.PZ)
        if self.full_cov:
            K.gmvae_full_prior(p.PZ, Kc, L, p.PL)
            K.gmvae_latent_full_fwd(p.QH, p.PZ, p.PL, Kc, B, L, RS, p.eps, p.Z, p.klz, p.Wsol)
        else:
            K.gmvae_latent_fwd(p.QH, p.PZ, Kc, B, L, RS, p.eps, p.Z, p.klz, p.kl_elem)
        self._decoder_features(p, p.M)
        K.gmvae_row_coefficients(p.y, Kc, RS, B, weight, p.go, p.coef)
        # --- p(x|z_k): decoder + heads + likelihood, cluster chunk by cluster chunk -----------
        tgt = p.T if p.use_T else p.X
        rc = p.row_const if p.have_row_const else None
        rows_per_k = RS * B
        fused = bool(with_backward and p.have_t16 and not p.use_T
                     and self._fused_possible(min(p.chunk, Kc) * rows_per_k, B))
        if fused:
            self._plan_fused_chunks(p)
            l = self.head
            for h in range(self.P):        # fp16 shadow of the head weights for this step
                K.f32_to_f16(l.w[h * self.Gn:(h + 1) * self.Gn], l.in_p,
                             p.W16[h * self.Gh:h * self.Gh + self.Gn])
            p.fused_scale = 2.0 ** round(math.log2(max(RS * B, 16) / 16.0))
        p.fused_done = fused
        for c0 in range(0, Kc, p.chunk):
            kc = min(p.chunk, Kc - c0)
            r0, rows = c0 * rows_per_k, kc * rows_per_k
            if self.constrained:
                p.lse = p.lse_all[r0:r0 + rows]
            d = p.Z[r0:r0 + rows]
            for j, l in enumerate(self.dec):
                keep = (self.keep_z if j == 0 else self.keep_h) if drop else None
                src = d
                if keep:
                    src = self._gdrop(p, l.name, d[:rows], r0, rows, p.M, l.n_in + l.n_extra,
                                      l.n_in, keep, Kc).copy
                self._gemm(p, K.GEMM_NT, rows, l.n_out, l.k_in, src, l.w, p.decY[j][:rows])
                self._bn_fwd(p, l, p.decY[j][:rows], p.decH[j][:rows], p.dec_mean[j], p.dec_rstd[j],
                             is_training, update_moving, kc)
                d = p.decH[j]
            l = self.head
            if fused:
                # heads GEMM + likelihood + gradient + head dgrad in one kernel (heads_fused.cu);
                # row m of the chunk reads target row m % B, its upstream gradient is go[m]
                K.f32_to_f16(d[:rows], l.in_p, p.D16[:rows])
                dd = p.d_decH[-1] if self.dec else p.dZ[r0:r0 + rows]
                t16 = p.X16 if p.t16_is_x16 else p.T16
                K.heads_fused_bwd(self.kind, p.D16[:rows], p.W16, self.Gh, t16, rows, self.G,
                                  p.dA16[:rows], dd[:rows], l.n_in, p.logp[r0:r0 + rows],
                                  p.fused_ws, row_const=rc, go=p.go[r0:r0 + rows],
                                  scale=p.fused_scale)
                for h in range(self.P):
                    K.gemm_f16(K.GEMM_TN, self.Gn, l.in_p, rows,
                               p.dA16[:rows, h * self.Gh:h * self.Gh + self.Gn], p.D16[:rows],
                               l.dw[h * self.Gn:(h + 1) * self.Gn], accumulate=c0 > 0,
                               alpha=1.0 / p.fused_scale)
                self._decoder_backward(p, r0, rows, kc, accumulate=c0 > 0, heads_done=True)
                if getattr(p, "on_chunk", None) is not None:
                    p.on_chunk(c0, kc, rows)
                continue
            if drop and self.keep_h:
                # every head is a dense layer of its own with its own mask (GMVAE:3140-3218)
                for b0, nr, site in self._head_blocks():
                    st = self._gdrop(p, site, d[:rows], r0, rows, p.M, l.n_in, l.n_in,
                                     self.keep_h, Kc)
                    K.gemm(K.GEMM_NT, rows, nr, l.n_in + 1, st.copy, l.w[b0:b0 + nr],
                           p.A[:rows, b0:b0 + nr], tensor_cores=False)
            else:
                self._gemm(p, K.GEMM_NT, rows, l.n_out, l.n_in + 1, d, l.w, p.A[:rows])
            if with_backward:
                self._likelihood(p, tgt, p.A[:rows], rows, rc, logp=p.logp[r0:r0 + rows],
                                 da=p.dA[:rows], go=p.go[r0:r0 + rows])
                self._decoder_backward(p, r0, rows, kc, accumulate=c0 > 0)
            else:
                self._likelihood(p, tgt, p.A[:rows], rows, rc, logp=p.logp[r0:r0 + rows])
            if getattr(p, "on_chunk", None) is not None:
                p.on_chunk(c0, kc, rows)
        # free nats: the kernel forms threshold = proportion * H[p(y)] (GMVAE:3260-3261) from
        # log_py on the device -- no host read inside a captured step, and a learnt prior's
        # threshold follows the prior
        K.gmvae_bound(p.y, p.logy, p.logp, p.klz, self.log_py, Kc, RS, B, weight, self.free_nats,
                      self.prior_method == "uniform", p.bound,
                      p.dlogits_c if with_backward else None,
                      self.d_py_logits if (with_backward and self.prior_method == "learn") else None,
                      p.ll_mean, p.klz_mean)
        return p

    def _refresh_log_py(self):
        # log_softmax of K learned logits (K floats: host-side glue of the shell)
        self.log_py = torch.log_softmax(self.py_logits.detach(), dim=0)

    def _plan_fused_chunks(self, p):
        """fp16 operand / gradient buffers of the fused heads kernel for one decoder chunk."""
        if p.fused_ready:
            return
        dev = self.device
        Mc = p.chunk * p.RS * p.B
        p.D16 = torch.zeros(Mc, 128, dtype=torch.float16, device=dev)
        p.W16 = torch.zeros(self.P * self.Gh, 128, dtype=torch.float16, device=dev)
        p.dA16 = torch.zeros(Mc, self.P * self.Gh, dtype=torch.float16, device=dev)
        p.fused_ws = torch.zeros(K.heads_fused_workspace_floats(Mc, self.G), dtype=torch.float32,
                                 device=dev)
        p.fused_ready = True

    def _decoder_backward(self, p, r0, rows, kc, accumulate, heads_done=False):
        l = self.head
        d_in = p.decH[-1] if self.dec else p.Z[r0:r0 + rows]
        dd_in = p.d_decH[-1] if self.dec else p.dZ[r0:r0 + rows]
        if not heads_done and p.drop_on and self.keep_h:
            tmp = self._drop_tmp(p, p.decH[-1].shape[0] if self.dec else p.M, l.in_p)
            for b, (b0, nr, site) in enumerate(self._head_blocks()):
                st = p.drop[site]
                K.gemm(K.GEMM_TN, nr, l.in_p, rows, p.dA[:rows, b0:b0 + nr], st.copy[:rows],
                       l.dw[b0:b0 + nr], accumulate=accumulate, tensor_cores=False)
                K.gemm(K.GEMM_NN, rows, l.n_in, nr, p.dA[:rows, b0:b0 + nr], l.w[b0:b0 + nr],
                       tmp[:rows], tensor_cores=False)
                self._gdrop_bwd(st, dd_in[:rows], dsrc=tmp[:rows], accumulate=b > 0)
        elif not heads_done:
            self._gemm(p, K.GEMM_TN, l.n_out, l.in_p, rows, p.dA[:rows], d_in[:rows], l.dw,
                       accumulate=accumulate)
            self._gemm(p, K.GEMM_NN, rows, l.n_in, l.n_out, p.dA[:rows], l.w, dd_in[:rows])
        for j in range(len(self.dec) - 1, -1, -1):
            l = self.dec[j]
            self._bn_bwd(p, l, p.d_decH[j][:rows], p.decY[j][:rows], p.decH[j][:rows],
                         p.dec_mean[j], p.dec_rstd[j], p.d_decY[j][:rows], kc, accumulate)
            d_in = p.decH[j - 1][:rows] if j > 0 else p.Z[r0:r0 + rows]
            dd_in = p.d_decH[j - 1][:rows] if j > 0 else p.dZ[r0:r0 + rows]
            st = p.drop.get(l.name) if p.drop_on else None
            if st is not None:
                d_in = st.copy[:rows]
            self._gemm(p, K.GEMM_TN, l.n_out, l.in_p, rows, p.d_decY[j][:rows], d_in, l.dw,
                       accumulate=accumulate)
            self._gemm(p, K.GEMM_NN, rows, l.n_in, l.n_out, p.d_decY[j][:rows], l.w, dd_in)
            if st is not None:
                self._gdrop_bwd(st, dd_in)      # gradient w.r.t. the un-dropped input, in place

    # ------------------------------------------------------------------ backward -----------
    def backward(self, p, R, S, warm_up_weight=1.0):
        """Everything upstream of z (the decoder part already ran chunk-wise in forward)."""
        B, Kc, L, RS, KB = p.B, self.K, self.L, p.RS, p.KB
        if self.full_cov:
            K.gmvae_latent_full_bwd(p.QH, p.PZ, p.PL, Kc, B, L, RS, p.eps, p.dZ, p.coef, p.Wsol, p.CU,
                                    p.dQH, p.dPZ)
        else:
            K.gmvae_latent_bwd(p.QH, p.PZ, Kc, B, L, RS, p.eps, p.dZ, p.coef, p.dQH, p.dPZ)
        # p(z|y) parameters: W[k] gets dPZ[k], the shared bias their sum
        drop = p.drop_on
        if drop and self.keep_y:
            torch.mul(p.dPZ, p.pz_scale, out=self.d_pz_w)     # W[k] enters as (mask_kk / keep) W[k]
        else:
            self.d_pz_w.copy_(p.dPZ)
        K.group_offset_bwd(p.dPZ, 1, Kc, self.nL, dt=self.d_pz_b)
        # q(z|x,y) encoder on K*B rows
        l = self.qz_head
        if drop and self.keep_h:
            tmp = self._drop_tmp(p, KB, l.in_p)
            for part, name in enumerate(("MEAN", "SOFTPLUS_SCALE")):
                st = p.drop[l.name + "/" + name]
                cols = slice(part * L, (part + 1) * L)
                K.gemm(K.GEMM_TN, L, l.in_p, KB, p.dQH[:, cols], st.copy, l.dw[cols],
                       tensor_cores=False)
                K.gemm(K.GEMM_NN, KB, l.n_in, L, p.dQH[:, cols], l.w[cols], tmp, tensor_cores=False)
                self._gdrop_bwd(st, p.d_qzH[-1], dsrc=tmp, accumulate=part > 0)
        else:
            self._gemm(p, K.GEMM_TN, l.n_out, l.in_p, KB, p.dQH, p.qzH[-1], l.dw)
            self._gemm(p, K.GEMM_NN, KB, l.n_in, l.n_out, p.dQH, l.w, p.d_qzH[-1])
        for i in range(len(self.qz_enc) - 1, -1, -1):
            l = self.qz_enc[i]
            self._bn_bwd(p, l, p.d_qzH[i], p.qzY[i], p.qzH[i], p.qz_mean[i], p.qz_rstd[i],
                         p.d_qzY[i], Kc, False)
            if i > 0:
                st = p.drop.get(l.name) if drop else None
                h_in = st.copy if st is not None else p.qzH[i - 1]
                self._gemm(p, K.GEMM_TN, l.n_out, l.in_p, KB, p.d_qzY[i], h_in, l.dw)
                self._gemm(p, K.GEMM_NN, KB, l.n_in, l.n_out, p.d_qzY[i], l.w, p.d_qzH[i - 1])
                if st is not None:
                    self._gdrop_bwd(st, p.d_qzH[i - 1])
            elif drop and self.keep_x:
                self._qz_first_layer_dropped_bwd(p)
            else:
                K.group_offset_bwd(p.d_qzY[0], Kc, B, l.n_out, dx=p.dXW, dt=self.d_qz_wy)
                self._x_wgrad(p, l, p.dXW)
        # q(y|x) encoder
        p.dlogits[:, :Kc].copy_(p.dlogits_c)
        l = self.qy_logits
        st = p.drop.get(l.name) if drop else None
        self._gemm(p, K.GEMM_TN, l.n_out, l.in_p, B, p.dlogits,
                   st.copy if st is not None else p.qyH[-1], l.dw)
        self._gemm(p, K.GEMM_NN, B, l.n_in, l.n_out, p.dlogits, l.w, p.d_qyH[-1])
        if st is not None:
            self._gdrop_bwd(st, p.d_qyH[-1])
        for i in range(len(self.qy_enc) - 1, -1, -1):
            l = self.qy_enc[i]
            self._bn_bwd(p, l, p.d_qyH[i], p.qyY[i], p.qyH[i], p.qy_mean[i], p.qy_rstd[i],
                         p.d_qyY[i], 1, False)
            st = p.drop.get(l.name) if drop else None
            if i > 0:
                self._gemm(p, K.GEMM_TN, l.n_out, l.in_p, B, p.d_qyY[i],
                           st.copy if st is not None else p.qyH[i - 1], l.dw)
            elif st is not None:
                self._gemm(p, K.GEMM_TN, l.n_out, l.in_p, B, p.d_qyY[i], st.copy, l.dw)
            else:
                self._x_wgrad(p, l, p.d_qyY[i])
            if i > 0:
                self._gemm(p, K.GEMM_NN, B, l.n_in, l.n_out, p.d_qyY[i], l.w, p.d_qyH[i - 1])
                if st is not None:
                    self._gdrop_bwd(st, p.d_qyH[i - 1])

    def train_step(self, p, R, S, learning_rate, warm_up_weight=1.0):
        """One ``session.run([optimiser, lower_bound])`` (GMVAE:1109-1112)."""
        self.forward(p, True, R, S, warm_up_weight, with_backward=True)
        self.backward(p, R, S, warm_up_weight)
        self.optimiser_step(learning_rate)
        return p.bound

    # ------------------------------------------------------------------ evaluate extras ----
    def z_mean(self, p):
        K.gmvae_z_mean(p.QH, p.y, self.K, p.B, self.L, p.z_mean)
        return p.z_mean

    def _moments_launch(self, p, RS, mean, stddev, stddev_of_mean):
        outs = (mean, stddev, stddev_of_mean)
        if self.k_max:
            K.piecewise_moments(self.kind, self.k_max, p.A, self.Gn, p.B, self.G, p.RS, *outs,
                                K_=self.K, y=p.y)
        elif self.constrained:
            K.constrained_poisson_mixture_moments(p.A, p.lse_all, p.count_sum_parameter, p.B,
                                                  self.G, p.RS, self.K, p.y, *outs)
        elif self.continuous:
            K.continuous_moments(self.kind, p.A, self.Gn, p.B, self.G, p.RS, self.K, p.y, *outs)
        else:
            K.likelihood_moments(self.kind, p.A, self.Gn, p.B, self.G, p.RS, self.K, p.y, *outs)

    def moments(self, p, R, S, deterministic=False, mean_out=None, want_stddev=True):
        """p_x_mean, p_x_stddev, stddev_of_p_x_given_z_mean marginalised over y
        (GMVAE:3312-3386, including the y-weighted per-cluster mean of quirk Q7).  Needs the
        head pre-activations of all K clusters, i.e. a plan whose decoder runs in one chunk.
        ``mean_out`` / ``want_stddev`` as VAEEngine.moments."""
        if p.chunk != self.K:
            raise ValueError("moments need a single-chunk plan: lower the evaluation minibatch "
                             "size or raise head_buffer_bytes")
        bufs = getattr(p, "moment_bufs", None)
        if bufs is None:
            bufs = p.moment_bufs = [torch.empty(p.B, self.Gn, dtype=torch.float32, device=self.device)
                                    for _ in range(3)]
        if mean_out is None:
            self._moments_launch(p, p.RS, bufs[0], bufs[1] if want_stddev else None,
                                 bufs[2] if want_stddev else None)
            mean = bufs[0]
        else:
            self._moments_launch(p, p.RS, mean_out, None, None)
            if want_stddev:
                self._moments_launch(p, p.RS, None, bufs[1], bufs[2])
            mean = mean_out
        return [mean[:, :self.G], bufs[1][:, :self.G] if want_stddev else None,
                bufs[2][:, :self.G] if want_stddev else None]

    def max_single_chunk_minibatch(self, RS):
        """Largest minibatch whose K cluster passes fit the head buffers in one chunk."""
        per_cell = self.K * RS * self.PT * self.Gn * 4 * 2
        return max(1, self.head_buffer_bytes // per_cell)
