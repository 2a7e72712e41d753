"""CPU stand-ins for the kernel wrappers of ``scvae_b200.kernels`` -- TEST INFRASTRUCTURE.

Each function has the signature of the wrapper it replaces and the semantics documented for the
entry point in ``include/scvae_b200.h`` (operand layouts, augmented ones columns, leading
dimensions, in-place outputs), restated with PyTorch-CPU ops in fp32 storage.  They exist so
that the HOST logic of the engines -- buffer layout, launch order, which gradient lands where,
the optimiser wiring -- can be exercised by ``-m "not gpu"`` tests against the golden vectors
recorded from the reference's graph code (``tests/test_engine_host_logic.py``).  They are never
imported by the product: without the CUDA library and a device every product entry point raises.
Only the exact-fp32 paths of the VAE and GMVAE engines are covered (no tensor-core, fused-heads,
CSR or data-parallel kernels).
"""
import math

import numpy
import torch

from oracle import scvae_oracle as O
from scvae_b200 import kernels as RealK
from test_sampled_kl_math import bound_rows, sampled_kl_bwd, sampled_kl_rows

# constants and pure host helpers are the real ones
GEMM_NT, GEMM_NN, GEMM_TN = RealK.GEMM_NT, RealK.GEMM_NN, RealK.GEMM_TN
LIKELIHOOD_KINDS = RealK.LIKELIHOOD_KINDS
LIKELIHOOD_HEADS = RealK.LIKELIHOOD_HEADS
CONSTRAINED_POISSON = RealK.CONSTRAINED_POISSON
CONTINUOUS_KINDS = RealK.CONTINUOUS_KINDS
_KIND_NAMES = {v: k for k, v in LIKELIHOOD_KINDS.items()}
BN_EPSILON, BN_DECAY = 1e-3, 0.999

launches = []        # names of the stand-ins called, in order (the launch sequence of a step)


def _log(name):
    launches.append(name)


def _view(t, rows, cols):
    """(rows, cols) window at the tensor's base pointer with its leading dimension -- the
    kernels see only (pointer, ld), not the declared shape."""
    if t.dim() == 1:
        return t.as_strided((rows, cols), (cols, 1), t.storage_offset())
    return t.as_strided((rows, cols), (t.stride(0), 1), t.storage_offset())


def bn_scratch_floats(M, H, groups):
    return 1


def gemm_workspace_bytes(layout, M, N, K):
    return 0


def gemm(layout, M, N, K, A, B, C, accumulate=False, tensor_cores=True, workspace=None):
    _log("gemm")
    assert not tensor_cores
    if layout == GEMM_NT:
        c = _view(A, M, K).double() @ _view(B, N, K).double().t()
    elif layout == GEMM_NN:
        c = _view(A, M, K).double() @ _view(B, K, N).double()
    else:
        c = _view(A, K, M).double().t() @ _view(B, K, N).double()
    out = _view(C, M, N)
    out.copy_((c + out.double()) if accumulate else c)


def _augment(out, H):
    if out.shape[1] > H:
        out[:, H] = 1.0
        out[:, H + 1:] = 0.0


def act_fwd(y, H, out, relu=True):
    _log("act_fwd")
    v = y[:, :H].clone()
    out[:, :H] = torch.relu(v) if relu else v
    _augment(out, H)


def act_bwd(dout, out, H, dy, relu=True):
    _log("act_bwd")
    g = dout[:, :H]
    dy[:, :H] = g * (out[:, :H] > 0) if relu else g


def bn_act_fwd(y, H, beta, moving_mean, moving_var, out, save_mean, save_rstd, scratch,
               training=True, update_moving=True, relu=True, groups=1):
    _log("bn_act_fwd")
    M = y.shape[0]
    n = M // groups
    v = y[:, :H].double().reshape(groups, n, H)
    if training:
        mean = v.mean(dim=1)
        var = ((v - mean.unsqueeze(1)) ** 2).mean(dim=1)
        rstd = 1.0 / torch.sqrt(var + BN_EPSILON)
        save_mean[:groups * H] = mean.reshape(-1)
        save_rstd[:groups * H] = rstd.reshape(-1)
        if update_moving:
            for g in range(groups):
                moving_mean -= ((1.0 - BN_DECAY) * (moving_mean.double() - mean[g])).float()
                unbiased = var[g] * (n / max(n - 1, 1))
                moving_var -= ((1.0 - BN_DECAY) * (moving_var.double() - unbiased)).float()
    else:
        mean = moving_mean.double().expand(groups, H)
        rstd = (1.0 / torch.sqrt(moving_var.double() + BN_EPSILON)).expand(groups, H)
    r = (v - mean.unsqueeze(1)) * rstd.unsqueeze(1) + beta.double()
    if relu:
        r = torch.relu(r)
    out[:, :H] = r.reshape(M, H)
    _augment(out, H)


def bn_act_bwd(dout, y, out, H, save_mean, save_rstd, dy, dbeta, scratch, relu=True, groups=1,
               accumulate_dbeta=False):
    _log("bn_act_bwd")
    M = y.shape[0]
    n = M // groups
    g = dout[:, :H].double()
    if relu:
        g = g * (out[:, :H] > 0)
    g = g.reshape(groups, n, H)
    mean = save_mean[:groups * H].double().reshape(groups, 1, H)
    rstd = save_rstd[:groups * H].double().reshape(groups, 1, H)
    xhat = (y[:, :H].double().reshape(groups, n, H) - mean) * rstd
    total = g.sum(dim=(0, 1))
    dbeta[:H] = (dbeta[:H].double() + total) if accumulate_dbeta else total
    d = rstd * (g - g.mean(dim=1, keepdim=True) - xhat * (g * xhat).mean(dim=1, keepdim=True))
    dy[:, :H] = d.reshape(M, H)


def _physical(n, skip_col):
    return [c if c < skip_col else c + 1 for c in range(n)]


def dropout_fwd(x, rows, n, skip_col, noise, threshold, keep, out, width):
    _log("dropout_fwd")
    out[:rows, :width] = x[:rows, :width]
    cols = _physical(n, skip_col)
    mask = (noise[:rows, :n] < threshold).float()
    out[:rows, cols] = x[:rows, cols] * mask / keep


def dropout_bwd(dx, rows, n, skip_col, noise, threshold, keep, dsrc=None, accumulate=False):
    _log("dropout_bwd")
    cols = _physical(n, skip_col)
    mask = (noise[:rows, :n] < threshold).float()
    source = dx if dsrc is None else dsrc
    g = source[:rows, cols] * mask / keep
    dx[:rows, cols] = dx[:rows, cols] + g if (accumulate and dsrc is not None) else g


def gaussian_latent_fwd(ph, B, L, RS, eps, z, kl_row, kl_elem=None, unit_variance=False,
                        deterministic=False):
    _log("gaussian_latent_fwd")
    mu = ph[:B, :L].double()
    ls = torch.zeros_like(mu) if unit_variance else torch.clamp(ph[:B, L:2 * L].double(), -3, 3)
    sigma = torch.exp(ls)
    k = 0.5 * mu * mu + 0.5 * (sigma * sigma - 1.0) - ls
    if kl_elem is not None:
        kl_elem.reshape(-1)[:B * L] = k.reshape(-1)
    if kl_row is not None:
        kl_row[:B] = k.sum(dim=1)
    nrep = 1 if deterministic else RS
    for s in range(nrep):
        rows = slice(s * B, (s + 1) * B)
        z[rows, :L] = mu if deterministic else mu + sigma * eps[rows, :L].double()
        z[rows, L] = 1.0
        z[rows, L + 1:] = 0.0


def gaussian_latent_bwd(ph, B, L, RS, eps, dz, kl_coef, dph, unit_variance=False):
    _log("gaussian_latent_bwd")
    mu = ph[:B, :L].double()
    raw = torch.zeros_like(mu) if unit_variance else ph[:B, L:2 * L].double()
    sigma = torch.exp(torch.clamp(raw, -3, 3))
    d = dz[:RS * B, :L].double().reshape(RS, B, L)
    e = eps[:RS * B, :L].double().reshape(RS, B, L)
    dph[:B, :L] = d.sum(dim=0) + kl_coef * mu
    if not unit_variance:
        mask = ((raw >= -3) & (raw <= 3)).double()
        dph[:B, L:2 * L] = ((d * e).sum(dim=0) * sigma + kl_coef * (sigma * sigma - 1.0)) * mask


def gaussian_sampled_kl(ph, B, L, RS, eps, kl_rows, kl_elem=None, unit_variance=False,
                        deterministic=False):
    _log("gaussian_sampled_kl")
    nL = L if unit_variance else 2 * L
    rows, elem = sampled_kl_rows(ph[:B, :nL].double().numpy(),
                                 None if eps is None else eps.double().numpy(), B, L, RS,
                                 unit_variance, deterministic)
    kl_rows[:rows.size] = torch.as_tensor(rows)
    if kl_elem is not None:
        kl_elem.reshape(-1)[:B * L] = torch.as_tensor(elem).reshape(-1)


def gaussian_sampled_kl_bwd(ph, B, L, RS, eps, dz, go, weight, coef_scalar, dph,
                            unit_variance=False):
    _log("gaussian_sampled_kl_bwd")
    nL = L if unit_variance else 2 * L
    d = sampled_kl_bwd(ph[:B, :nL].double().numpy(), eps.double().numpy(),
                       dz[:RS * B, :L].double().numpy(),
                       None if go is None else go.double().numpy(), weight, coef_scalar, B, L,
                       RS, unit_variance)
    dph[:B, :nL] = torch.as_tensor(d)


def vae_bound_rows(logp, kl_rows, R, S, B, weight, out, go=None):
    _log("vae_bound_rows")
    o, g = bound_rows(logp.double().numpy(), kl_rows.double().numpy(), R, S, B, weight)
    out[:4] = torch.as_tensor(o)
    if go is not None:
        go[:R * S * B] = torch.as_tensor(g)


def vae_bound(logp, kl_row, R, S, B, weight, out, go=None):
    _log("vae_bound")
    rows = numpy.tile(kl_row[:B].double().numpy(), R * S)
    o, g = bound_rows(logp.double().numpy(), rows, R, S, B, weight)
    out[:4] = torch.as_tensor(o)
    if go is not None:
        go[:R * S * B] = torch.as_tensor(g)


def decoder_features(z, M, B, col0, batch_index=None, n_batches=0, count_sum=None):
    _log("decoder_features")
    for m in range(M):
        b = m % B
        col = col0
        if batch_index is not None and n_batches:
            z[m, col:col + n_batches] = 0.0
            z[m, col + int(batch_index[b])] = 1.0
            col += n_batches
        if count_sum is not None:
            z[m, col] = count_sum[b]


def _theta(kind, a, head_stride, M, G):
    name = _KIND_NAMES[kind]
    heads = LIKELIHOOD_HEADS[name]
    return name, {head: O._clip_head(a[:M, h * head_stride:h * head_stride + G], head)
                  for h, head in enumerate(heads)}


def _targets(t, M, G):
    rows = torch.arange(M) % t.shape[0]
    return t[rows, :G].double()


LOGIT_FLOOR = -87.33654475      # log(float32 tiny): sigmoid heads are clipped to [tiny, 1]


def _safe_log_prob(name, x, heads):
    """Count log-pmf from the head PRE-activations, as the kernels evaluate it (DESIGN §2,
    "known deviations": the sigmoid pre-activation is the NB / ZI logit): identical to the
    oracle's formulas wherever those are finite, and finite where sigmoid(a) rounds to one."""
    logsig = torch.nn.functional.logsigmoid
    if "poisson" in name:
        ll = torch.clamp(heads["log_lambda"], -10.0, 10.0)
        lp = x * ll - torch.lgamma(1.0 + x) - torch.exp(ll)
    else:
        a = torch.clamp(heads["p"], min=LOGIT_FLOOR)
        r = torch.exp(torch.clamp(heads["log_r"], -10.0, 10.0))
        lp = (torch.lgamma(x + r) - torch.lgamma(1.0 + x) - torch.lgamma(r)
              + r * logsig(-a) + x * logsig(a))
    if name.startswith("zero-inflated"):
        b = torch.clamp(heads["pi"], min=LOGIT_FLOOR)
        lp = torch.where(x > 0, logsig(-b) + lp, torch.logaddexp(logsig(b), logsig(-b) + lp))
    return lp


def _safe_moments(name, heads):
    if "poisson" in name:
        mean = var = torch.exp(torch.clamp(heads["log_lambda"], -10.0, 10.0))
    else:
        a = torch.clamp(heads["p"], min=LOGIT_FLOOR)
        mean = torch.exp(torch.clamp(heads["log_r"], -10.0, 10.0)) * torch.exp(a)
        var = mean * (1.0 + torch.exp(a))
    if name.startswith("zero-inflated"):
        keep = torch.sigmoid(-torch.clamp(heads["pi"], min=LOGIT_FLOOR))      # 1 - pi
        zi_mean = keep * mean
        return zi_mean, keep * (var + mean * mean) - zi_mean * zi_mean
    return mean, var


def _heads(kind, a, head_stride, M, G):
    name = _KIND_NAMES[kind]
    return name, {head: a[:M, h * head_stride:h * head_stride + G]
                  for h, head in enumerate(LIKELIHOOD_HEADS[name])}


def _rows_log_prob(kind, t, a, head_stride, M, G, k_max=0, count_sum=None):
    if not k_max and count_sum is None:
        name, heads = _heads(kind, a, head_stride, M, G)
        return _safe_log_prob(name, _targets(t, M, G), heads).sum(dim=-1)
    name, theta = _theta(kind, a, head_stride, M, G)
    x = _targets(t, M, G)
    n = None
    if count_sum is not None:
        n = count_sum.double()[torch.arange(M) % t.shape[0]].reshape(M, 1)
    if k_max:
        P = len(theta)
        logits = torch.stack([a[:M, (P + c) * head_stride:(P + c) * head_stride + G]
                              for c in range(k_max + 1)], dim=-1)
        log_p = O.piecewise_log_prob(name, x, theta, torch.log_softmax(logits, dim=-1), k_max, n)
    else:
        log_p = O.likelihood_log_prob(name, x, theta, n)
    return log_p.sum(dim=-1)


def _fwd_bwd(kind, t, a, head_stride, M, G, logp, da, go, go_scalar, k_max=0, count_sum=None):
    leaf = a.double().detach().clone().requires_grad_(da is not None)
    rows = _rows_log_prob(kind, t, leaf, head_stride, M, G, k_max, count_sum)
    if logp is not None:
        logp[:M] = rows.detach()
    if da is not None:
        upstream = go[:M].double() if go is not None else torch.full((M,), float(go_scalar),
                                                                     dtype=torch.float64)
        grad, = torch.autograd.grad(rows, leaf, grad_outputs=upstream)
        da[:M] = grad[:M, :da.shape[1]]


def likelihood_fwd(kind, t, a, head_stride, M, G, logp, row_const=None):
    _log("likelihood_fwd")
    _fwd_bwd(kind, t, a, head_stride, M, G, logp, None, None, 1.0)


def likelihood_bwd(kind, t, a, head_stride, M, G, da, logp=None, row_const=None, go=None,
                   go_scalar=1.0):
    _log("likelihood_bwd")
    _fwd_bwd(kind, t, a, head_stride, M, G, logp, da, go, go_scalar)


def piecewise_likelihood(kind, k_max, t, a, head_stride, M, G, logp=None, go=None, go_scalar=1.0,
                         da=None):
    _log("piecewise_likelihood")
    _fwd_bwd(kind, t, a, head_stride, M, G, logp, da, go, go_scalar, k_max=k_max)


def constrained_poisson(t, a, M, G, count_sum, logp=None, row_const=None, go=None, go_scalar=1.0,
                        da=None, lse=None):
    _log("constrained_poisson")
    _fwd_bwd(CONSTRAINED_POISSON, t, a, 0, M, G, logp, da, go, go_scalar, count_sum=count_sum)
    if lse is not None:
        lse[:M] = torch.logsumexp(a[:M, :G].double(), dim=1)


def continuous_likelihood(kind, t, a, head_stride, M, G, logp=None, go=None, go_scalar=1.0, da=None):
    """csrc/continuous.cu: activation + clip of the heads, closed-form log density, gradient."""
    _log("continuous_likelihood")
    leaf = a.double().detach().clone().requires_grad_(da is not None)
    name, theta = _theta(kind, leaf, head_stride, M, G)
    rows = O.continuous_log_prob(name, _targets(t, M, G), theta).sum(dim=-1)
    if logp is not None:
        logp[:M] = rows.detach()
    if da is not None:
        upstream = go[:M].double() if go is not None else torch.full((M,), float(go_scalar),
                                                                     dtype=torch.float64)
        grad, = torch.autograd.grad(rows, leaf, grad_outputs=upstream)
        da[:M] = grad[:M, :da.shape[1]]


def continuous_moments(kind, a, head_stride, B, G, RS, K_, y, p_x_mean, p_x_stddev, stddev_of_mean):
    _log("continuous_moments")
    name, theta = _theta(kind, a.double(), head_stride, K_ * RS * B, G)
    m, v = O.continuous_moments(name, theta)
    if K_ == 1 and y is None:
        _write_moments(m, v, B, G, RS, (p_x_mean, p_x_stddev, stddev_of_mean))
    else:
        _write_mixture_moments(m, v, y, K_, B, G, RS, (p_x_mean, p_x_stddev, stddev_of_mean))


def _write_moments(m, v, B, G, RS, outs):
    m = m.reshape(RS, B, G)
    v = v.reshape(RS, B, G)
    mean = m.mean(dim=0)
    var_of_mean = ((m - mean) ** 2).mean(dim=0)
    for out, value in zip(outs, (mean, torch.sqrt(var_of_mean + v.mean(dim=0)),
                                 torch.sqrt(var_of_mean))):
        if out is not None:
            out[:B, :G] = value


def likelihood_moments(kind, a, head_stride, B, G, RS, K, y, p_x_mean, p_x_stddev,
                       stddev_of_mean):
    _log("likelihood_moments")
    assert K == 1 and y is None
    name, heads = _heads(kind, a.double(), head_stride, RS * B, G)
    m, v = _safe_moments(name, heads)
    _write_moments(m, v, B, G, RS, (p_x_mean, p_x_stddev, stddev_of_mean))


def piecewise_moments(kind, k_max, a, head_stride, B, G, RS, p_x_mean, p_x_stddev, stddev_of_mean,
                      K=1, y=None):
    _log("piecewise_moments")
    a = a.double()
    name, theta = _theta(kind, a, head_stride, RS * B, G)
    P = len(theta)
    logits = torch.stack([a[:RS * B, (P + c) * head_stride:(P + c) * head_stride + G]
                          for c in range(k_max + 1)], dim=-1)
    m, v = O.likelihood_moments(name, theta)
    m, v = O.piecewise_moments(m, v, torch.log_softmax(logits, dim=-1), k_max)
    _write_moments(m, v, B, G, RS, (p_x_mean, p_x_stddev, stddev_of_mean))


def constrained_poisson_moments(a, lse, count_sum, B, G, RS, p_x_mean, p_x_stddev,
                                stddev_of_mean):
    _log("constrained_poisson_moments")
    theta = {"lambda": O._clip_head(a[:RS * B, :G].double(), "lambda")}
    n = count_sum.double()[torch.arange(RS * B) % B].reshape(-1, 1)
    m, v = O.likelihood_moments("constrained poisson", theta, n)
    _write_moments(m, v, B, G, RS, (p_x_mean, p_x_stddev, stddev_of_mean))


def col_mean(x, rows, cols, out):
    _log("col_mean")
    out[:cols] = x[:rows, :cols].double().mean(dim=0)


def adam_clip_ctas(n):
    return max(1, min(((int(n) >> 2) + 255) // 256, 148 * 4))


def adam_clip_step(param, grad, m, v, step, lr, beta1=0.9, beta2=0.999, epsilon=1e-8, clip=1.0,
                   grad_scale=1.0, scalars=None, shadows=None, advance_counter=None, advance_total=0):
    _log("adam_clip_step")
    assert not shadows, "fp16 shadows belong to the tensor-core path"
    if scalars is not None:
        lr = lr * float(scalars[0])
    t = int(step.item()) + 1
    g = torch.clamp(grad.double() * grad_scale, -clip, clip)
    m.copy_(beta1 * m.double() + (1.0 - beta1) * g)
    v.copy_(beta2 * v.double() + (1.0 - beta2) * g * g)
    lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    param.copy_(param.double() - lr_t * m.double() / (torch.sqrt(v.double()) + epsilon))
    if advance_counter is not None:      # the last CTA of the launches sharing the counter advances
        advance_counter += adam_clip_ctas(param.numel())
        if int(advance_counter.item()) >= advance_total:
            step += 1
            advance_counter.zero_()


def step_advance(step):
    _log("step_advance")
    step += 1


def fill_normal(out, seed, offset=0, offset_dev=None):
    _log("fill_normal")
    out.copy_(torch.randn(out.shape, generator=torch.Generator().manual_seed(int(seed))))


# ---- Gaussian-mixture VAE pieces (include/scvae_b200.h, "a9/a10") ------------------------------
F32_HALF_MAX = O.F32_HALF_MAX


def group_offset_fwd(x, t, K_, B, H, y):
    _log("group_offset_fwd")
    for k in range(K_):
        y[k * B:(k + 1) * B, :H] = x[:B, :H] + t[k, :H]


def group_offset_bwd(dy, K_, B, H, dx=None, dt=None, accumulate_dt=False):
    _log("group_offset_bwd")
    d = dy[:K_ * B, :H].double().reshape(K_, B, H)
    if dx is not None:
        dx[:B, :H] = d.sum(dim=0)
    if dt is not None:
        total = d.sum(dim=1)
        dt[:K_, :H] = (dt[:K_, :H].double() + total) if accumulate_dt else total


def softmax_fwd(logits, B, K_, y, logy):
    _log("softmax_fwd")
    lp = torch.log_softmax(logits[:B, :K_].double(), dim=-1)
    logy.reshape(-1)[:B * K_] = lp.reshape(-1)
    y.reshape(-1)[:B * K_] = torch.exp(lp).reshape(-1)


def _gmvae_latent(qh, pz, K_, B, L, RS, eps):
    """z (K, RS, B, L), per-element KL terms (K, RS, B, L) from [mean | s] heads."""
    q = torch.clamp(qh[:K_ * B, :2 * L], -F32_HALF_MAX, F32_HALF_MAX).reshape(K_, 1, B, 2 * L)
    pr = torch.clamp(pz[:K_, :2 * L], -F32_HALF_MAX, F32_HALF_MAX).reshape(K_, 1, 1, 2 * L)
    q_mean, q_scale = q[..., :L], torch.sqrt(torch.nn.functional.softplus(q[..., L:]))
    p_mean, p_scale = pr[..., :L], torch.sqrt(torch.nn.functional.softplus(pr[..., L:]))
    z = q_mean + q_scale * eps.double().reshape(K_, RS, B, L)
    kl = O._normal_log_prob(z, q_mean, q_scale) - O._normal_log_prob(z, p_mean, p_scale)
    return z, kl


def gmvae_latent_fwd(qh, pz, K_, B, L, RS, eps, z, klz, kl_elem=None):
    _log("gmvae_latent_fwd")
    zz, kl = _gmvae_latent(qh.double(), pz.double(), K_, B, L, RS, eps)
    M = K_ * RS * B
    z[:M, :L] = zz.reshape(M, L)
    z[:M, L] = 1.0
    z[:M, L + 1:] = 0.0
    klz[:M] = kl.sum(dim=-1).reshape(M)
    if kl_elem is not None:
        kl_elem.reshape(-1)[:M * L] = kl.reshape(-1)


def gmvae_latent_bwd(qh, pz, K_, B, L, RS, eps, dz, coef, dqh, dpz):
    _log("gmvae_latent_bwd")
    M = K_ * RS * B
    q_leaf = qh[:K_ * B, :2 * L].double().detach().clone().requires_grad_(True)
    p_leaf = pz[:K_, :2 * L].double().detach().clone().requires_grad_(True)
    zz, kl = _gmvae_latent(q_leaf, p_leaf, K_, B, L, RS, eps)
    total = (dz[:M, :L].double() * zz.reshape(M, L)).sum() + \
        (coef[:M].double() * kl.sum(dim=-1).reshape(M)).sum()
    gq, gp = torch.autograd.grad(total, [q_leaf, p_leaf])
    dqh[:K_ * B, :2 * L] = gq
    dpz[:K_, :2 * L] = gp


def _gmvae_latent_full(qh, pz, K_, B, L, RS, eps):
    """(z, kl) of the full-covariance mixture, rows ordered (k, rs, b), via the oracle's helpers."""
    T = L * (L + 1) // 2
    q_loc = qh[:K_ * B, :L].reshape(K_, 1, B, L)
    q_tril = O.fill_triangular(torch.clamp(torch.nn.functional.softplus(qh[:K_ * B, L:L + T]),
                                           min=O.F32_MIN)).reshape(K_, 1, B, L, L)
    p_loc = pz[:K_, :L].reshape(K_, 1, 1, L)
    p_tril = O.fill_triangular(torch.clamp(torch.nn.functional.softplus(pz[:K_, L:L + T]),
                                           min=O.F32_MIN)).reshape(K_, 1, 1, L, L)
    e = eps.reshape(-1)[:K_ * RS * B * L].double().reshape(K_, RS, B, L)
    z = q_loc + (q_tril @ e.unsqueeze(-1)).squeeze(-1)
    kl = O._mvn_log_prob(z, q_loc, q_tril) - O._mvn_log_prob(z, p_loc, p_tril)
    return z, kl, p_tril


def gmvae_full_prior(pz, K_, L, pl):
    _log("gmvae_full_prior")
    T = L * (L + 1) // 2
    pl.reshape(-1)[:K_ * L * L] = O.fill_triangular(torch.clamp(torch.nn.functional.softplus(
        pz[:K_, L:L + T].double()), min=O.F32_MIN)).reshape(-1)


def gmvae_latent_full_fwd(qh, pz, pl, K_, B, L, RS, eps, z, klz, w):
    _log("gmvae_latent_full_fwd")
    zz, kl, p_tril = _gmvae_latent_full(qh.double(), pz.double(), K_, B, L, RS, eps)
    M = K_ * RS * B
    z[:M, :L] = zz.reshape(M, L)
    z[:M, L] = 1.0
    z[:M, L + 1:] = 0.0
    klz[:M] = kl.reshape(M)
    r = (zz - pz[:K_, :L].double().reshape(K_, 1, 1, L)).unsqueeze(-1)
    w.reshape(-1)[:M * L] = torch.linalg.solve_triangular(
        p_tril.expand(K_, RS, B, L, L), r, upper=False).reshape(-1)


def gmvae_latent_full_bwd(qh, pz, pl, K_, B, L, RS, eps, dz, coef, w, cu, dqh, dpz):
    _log("gmvae_latent_full_bwd")
    M, T = K_ * RS * B, L * (L + 1) // 2
    q_leaf = qh[:K_ * B, :L + T].double().detach().clone().requires_grad_(True)
    p_leaf = pz[:K_, :L + T].double().detach().clone().requires_grad_(True)
    zz, kl, _ = _gmvae_latent_full(q_leaf, p_leaf, K_, B, L, RS, eps)
    total = (dz[:M, :L].double() * zz.reshape(M, L)).sum() + (coef[:M].double() * kl.reshape(M)).sum()
    gq, gp = torch.autograd.grad(total, [q_leaf, p_leaf])
    dqh[:K_ * B, :L + T] = gq
    dpz[:K_, :L + T] = gp


def gmvae_full_covariance_mean(qh, K_, B, L, cov):
    _log("gmvae_full_covariance_mean")
    T = L * (L + 1) // 2
    tril = O.fill_triangular(torch.clamp(torch.nn.functional.softplus(qh[:K_ * B, L:L + T].double()),
                                         min=O.F32_MIN)).reshape(K_, B, L, L)
    cov.reshape(-1)[:K_ * L * L] = (tril @ tril.transpose(-1, -2)).mean(dim=1).reshape(-1)


def gmvae_row_coefficients(y, K_, RS, B, weight, go, coef):
    _log("gmvae_row_coefficients")
    w = (y.reshape(-1)[:B * K_].double().reshape(B, K_).t() / (B * RS))     # (K, B)
    rows = w.unsqueeze(1).expand(K_, RS, B).reshape(-1)
    go[:K_ * RS * B] = -rows
    coef[:K_ * RS * B] = weight * rows


def gmvae_bound(y, logy, logp, klz, log_py, K_, RS, B, weight, free_nats_proportion, uniform_prior,
                out, dlogits, dpy_logits, ll_mean, klz_mean):
    _log("gmvae_bound")
    logits = logy.reshape(-1)[:B * K_].double().reshape(B, K_).detach().clone().requires_grad_(True)
    prior = log_py[:K_].double().detach().clone().requires_grad_(True)
    lq = torch.log_softmax(logits, dim=-1)
    q = torch.exp(lq)
    lpy = torch.log_softmax(prior, dim=-1)
    ll = logp[:K_ * RS * B].double().reshape(K_, RS, B).mean(dim=1)         # (K, B)
    kz = klz[:K_ * RS * B].double().reshape(K_, RS, B).mean(dim=1)
    ll_mean.reshape(-1)[:K_ * B] = ll.reshape(-1)
    klz_mean.reshape(-1)[:K_ * B] = kz.reshape(-1)
    reconstruction = (q.t() * ll).sum(dim=0).mean()
    kl_z = (q.t() * kz).sum(dim=0).mean()
    if uniform_prior:
        kl_y = (math.log(K_) + (q * lq).sum(dim=-1)).mean()
    else:
        kl_y = (q * (lq - lpy)).sum(dim=-1).mean()
    if free_nats_proportion:
        # H[p(y)] of the prior; differentiable for a learnt prior (GMVAE:3258-3261)
        threshold = float(free_nats_proportion) * (-(torch.exp(lpy) * lpy).sum())
        kl_y_mod = torch.where(kl_y > threshold, kl_y, threshold)
    else:
        kl_y_mod = kl_y
    weighted = reconstruction - weight * (kl_z + kl_y_mod)
    values = [reconstruction - kl_z - kl_y, weighted, reconstruction, kl_z, kl_y, kl_y_mod]
    out[:6] = torch.stack([v.detach() for v in values])
    if dlogits is not None or dpy_logits is not None:
        g_logits, g_prior = torch.autograd.grad(-weighted, [logits, prior], allow_unused=True)
        if dlogits is not None:
            dlogits.reshape(-1)[:B * K_] = g_logits.reshape(-1)
        if dpy_logits is not None:
            dpy_logits[:K_] = g_prior if g_prior is not None else 0.0


def gmvae_z_mean(qh, y, K_, B, L, z_mean):
    _log("gmvae_z_mean")
    mean = qh[:K_ * B, :L].double().reshape(K_, B, L)
    w = y.reshape(-1)[:B * K_].double().reshape(B, K_).t().unsqueeze(-1)
    z_mean.reshape(-1)[:B * L] = (mean * w).sum(dim=0).reshape(-1)


def _write_mixture_moments(m, v, y, K_, B, G, RS, outs):
    """GMVAE:3312-3386 with the y-weighted per-cluster mean of quirk Q7."""
    m = m.reshape(K_, RS, B, G)
    v = v.reshape(K_, RS, B, G)
    w = y.reshape(-1)[:B * K_].double().reshape(B, K_).t().unsqueeze(-1)      # (K, B, 1)
    means = m.mean(dim=1) * w
    mean_of_var = (v.mean(dim=1) * w).sum(dim=0)
    var_of_mean = (((m - means.unsqueeze(1)) ** 2).mean(dim=1) * w).sum(dim=0)
    for out, value in zip(outs, (means.sum(dim=0), torch.sqrt(mean_of_var + var_of_mean),
                                 torch.sqrt(var_of_mean))):
        if out is not None:
            out[:B, :G] = value


_vae_likelihood_moments = likelihood_moments
_vae_piecewise_moments = piecewise_moments


def likelihood_moments(kind, a, head_stride, B, G, RS, K, y, p_x_mean, p_x_stddev,  # noqa: F811
                       stddev_of_mean):
    if K == 1 and y is None:
        return _vae_likelihood_moments(kind, a, head_stride, B, G, RS, K, y, p_x_mean,
                                       p_x_stddev, stddev_of_mean)
    _log("likelihood_moments")
    name, heads = _heads(kind, a.double(), head_stride, K * RS * B, G)
    m, v = _safe_moments(name, heads)
    _write_mixture_moments(m, v, y, K, B, G, RS, (p_x_mean, p_x_stddev, stddev_of_mean))


def piecewise_moments(kind, k_max, a, head_stride, B, G, RS, p_x_mean, p_x_stddev,  # noqa: F811
                      stddev_of_mean, K_=1, y=None):
    if K_ == 1 and y is None:
        return _vae_piecewise_moments(kind, k_max, a, head_stride, B, G, RS, p_x_mean,
                                      p_x_stddev, stddev_of_mean)
    _log("piecewise_moments")
    a = a.double()
    rows = K_ * RS * B
    name, theta = _theta(kind, a, head_stride, rows, G)
    P = len(theta)
    logits = torch.stack([a[:rows, (P + c) * head_stride:(P + c) * head_stride + G]
                          for c in range(k_max + 1)], dim=-1)
    m, v = O.likelihood_moments(name, theta)
    m, v = O.piecewise_moments(m, v, torch.log_softmax(logits, dim=-1), k_max)
    _write_mixture_moments(m, v, y, K_, B, G, RS, (p_x_mean, p_x_stddev, stddev_of_mean))


# ---- minibatch assembly (a1) -------------------------------------------------------------------
def csr_row_constants(indptr, values, out):
    _log("csr_row_constants")
    v = torch.lgamma(1.0 + values.double())
    csum = torch.cat([torch.zeros(1, dtype=torch.float64), torch.cumsum(v, dim=0)])
    out[:] = csum[indptr[1:].long()] - csum[indptr[:-1].long()]


def gather_f32(src, rows, dst):
    _log("gather_f32")
    n = dst.numel()
    dst.reshape(-1)[:] = src[:n] if rows is None else src[rows[:n].long()]


def csr_densify(indptr, indices, values, rows, G, x, row_const=None, rebase=False, t16=None,
                x16=None):
    _log("csr_densify")
    assert x is not None and t16 is None and x16 is None, "fp32 minibatch only"
    B = x.shape[0]
    base = int(indptr[0]) if rebase else 0
    x[:, :G] = 0.0
    for b in range(B):
        r = b if rows is None else int(rows[b])
        lo, hi = int(indptr[r]) - base, int(indptr[r + 1]) - base
        x[b, indices[lo:hi].long()] = values[lo:hi].float()
    _augment(x, G)
    if row_const is not None:
        row_const[:B] = torch.lgamma(1.0 + x[:, :G].double()).sum(dim=1)


def constrained_poisson_mixture_moments(a, lse, count_sum, B, G, RS, K_, y, p_x_mean, p_x_stddev,
                                        stddev_of_mean):
    _log("constrained_poisson_mixture_moments")
    rows = K_ * RS * B
    rate = count_sum.double()[torch.arange(rows) % B].reshape(-1, 1) * torch.exp(torch.clamp(
        a[:rows, :G].double() - lse[:rows].double().reshape(-1, 1), min=math.log(O.TINY)))
    _write_mixture_moments(rate, rate, y, K_, B, G, RS, (p_x_mean, p_x_stddev, stddev_of_mean))
