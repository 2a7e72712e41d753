"""The arithmetic of the sampled-KL kernels (``csrc/latent.cu``: gaussian_sampled_kl_kernel,
vae_bound_rows_kernel, gaussian_sampled_kl_bwd_kernel) restated loop for loop in numpy and held
against (a) autograd through the textbook definition log q(z|x) - log p(z) (VAE:2628-2640) and
(b) the golden vectors recorded from the reference's own graph code.  CPU only: it checks the
derivation the kernels implement, not the kernels (those: tests/test_zz_gpu_reference_graph.py).
"""
import math

import numpy
import pytest
import torch

from test_reference_graph import load_case, oracle_config, oracle_inputs

from oracle import scvae_oracle as O

D = torch.float64


def sampled_kl_rows(ph, eps, B, L, RS, unit_variance, deterministic=False):
    """gaussian_sampled_kl_kernel: kl_rows[RS*B], kl_elem (B, L)."""
    nrep = 1 if deterministic else RS
    kl_rows = numpy.zeros(nrep * B)
    kl_elem = numpy.zeros((B, L))
    for b in range(B):
        for s in range(nrep):
            m = s * B + b
            acc = 0.0
            for l in range(L):
                mu = ph[b, l]
                ls = 0.0 if unit_variance else min(max(ph[b, L + l], -3.0), 3.0)
                e = 0.0 if deterministic else eps[m, l]
                zv = mu + math.exp(ls) * e
                k = 0.5 * zv * zv - 0.5 * e * e - ls
                acc += k
                kl_elem[b, l] += k / nrep
            kl_rows[m] = acc
    return kl_rows, kl_elem


def bound_rows(logp, kl_rows, R, S, B, weight):
    """vae_bound_rows_kernel: out[4], go[R*S*B]."""
    SB = S * B
    go = numpy.zeros(R * SB)
    lb = lbw = lp_sum = kl_sum = 0.0
    for i in range(SB):
        a = numpy.array([logp[r * SB + i] - kl_rows[r * SB + i] for r in range(R)])
        aw = numpy.array([logp[r * SB + i] - weight * kl_rows[r * SB + i] for r in range(R)])
        mx, mxw = a.max(), aw.max()
        se, sew = numpy.exp(a - mx).sum(), numpy.exp(aw - mxw).sum()
        lb += math.log(se / R) + mx
        lbw += math.log(sew / R) + mxw
        for r in range(R):
            lp_sum += logp[r * SB + i]
            kl_sum += kl_rows[r * SB + i]
            go[r * SB + i] = -math.exp(aw[r] - mxw) / sew / SB
    return numpy.array([lb / SB, lbw / SB, lp_sum / SB / R, kl_sum / SB / R]), go


def sampled_kl_bwd(ph, eps, dz, go, weight, coef_scalar, B, L, RS, unit_variance):
    """gaussian_sampled_kl_bwd_kernel: dph (B, L or 2L)."""
    dph = numpy.zeros_like(ph)
    for b in range(B):
        for l in range(L):
            mu = ph[b, l]
            raw = 0.0 if unit_variance else ph[b, L + l]
            ls = min(max(raw, -3.0), 3.0)
            sigma = math.exp(ls)
            dmu = dls = 0.0
            for s in range(RS):
                m = s * B + b
                e = eps[m, l]
                c = -weight * go[m] if go is not None else coef_scalar
                zv = mu + sigma * e
                d = dz[m, l] + c * zv
                dmu += d
                dls += d * sigma * e - c
            dph[b, l] = dmu
            if not unit_variance:
                dph[b, L + l] = dls * (0.0 if (raw < -3.0 or raw > 3.0) else 1.0)
    return dph


@pytest.mark.parametrize("R,S,unit_variance", [(1, 1, False), (1, 3, False), (3, 2, False),
                                                (2, 2, True)])
def test_sampled_kl_kernels_arithmetic_matches_autograd(R, S, unit_variance):
    B, L, RS, weight = 5, 4, R * S, 0.6
    gen = torch.Generator().manual_seed(R * 10 + S)
    ph = torch.randn(B, L if unit_variance else 2 * L, generator=gen, dtype=D)
    if not unit_variance:
        ph[:, L:] *= 2.5                     # some log_sigma pre-activations beyond the +-3 clip
    eps = torch.randn(RS * B, L, generator=gen, dtype=D)
    a = torch.randn(L, 7, generator=gen, dtype=D)
    ph.requires_grad_(True)

    def decoder_logp(z):                    # any smooth per-row function of z stands in
        return -(torch.tanh(z @ a) ** 2).sum(dim=1) * 3.0

    mu = ph[:, :L]
    log_sigma = torch.zeros_like(mu) if unit_variance else torch.clamp(ph[:, L:], -3.0, 3.0)
    sigma = torch.exp(log_sigma)
    z = (mu.unsqueeze(0) + sigma.unsqueeze(0) * eps.reshape(RS, B, L))
    z_rows = z.reshape(RS * B, L)
    logp = decoder_logp(z_rows)
    log_q = torch.distributions.Normal(mu, sigma).log_prob(z)
    log_pz = torch.distributions.Normal(torch.zeros_like(mu), torch.ones_like(mu)).log_prob(z)
    kl = (log_q - log_pz).sum(dim=-1).reshape(R, S, B)
    lp = logp.reshape(R, S, B)
    lower_bound = O.log_mean_exp(lp - kl, 0).mean()
    loss = -O.log_mean_exp(lp - weight * kl, 0).mean()
    dph_ref, = torch.autograd.grad(loss, ph)

    ph_n, eps_n = ph.detach().numpy(), eps.numpy()
    kl_rows, kl_elem = sampled_kl_rows(ph_n, eps_n, B, L, RS, unit_variance)
    assert numpy.allclose(kl_rows, kl.detach().reshape(-1).numpy(), rtol=1e-12, atol=1e-12)
    assert numpy.allclose(kl_elem.mean(axis=0),
                          (log_q - log_pz).detach().reshape(-1, L).mean(dim=0).numpy())
    out, go = bound_rows(logp.detach().numpy(), kl_rows, R, S, B, weight)
    assert math.isclose(out[0], lower_bound.item(), rel_tol=1e-12)
    assert math.isclose(out[1], -loss.item(), rel_tol=1e-12)
    assert math.isclose(out[2], lp.mean().item(), rel_tol=1e-12)
    assert math.isclose(out[3], kl.mean().item(), rel_tol=1e-12)
    # dz as the decoder backward delivers it: go[m] * d logp_m / d z
    zd = z_rows.detach().clone().requires_grad_(True)
    dz, = torch.autograd.grad(decoder_logp(zd), zd, grad_outputs=torch.as_tensor(go))
    dph = sampled_kl_bwd(ph_n, eps_n, dz.numpy(), go, weight, 0.0, B, L, RS, unit_variance)
    assert numpy.allclose(dph, dph_ref.numpy(), rtol=1e-10, atol=1e-12)
    if R == 1:      # the scalar-coefficient form used when go is never materialised
        dph1 = sampled_kl_bwd(ph_n, eps_n, dz.numpy(), None, weight, weight / (S * B), B, L, RS,
                              unit_variance)
        assert numpy.allclose(dph1, dph_ref.numpy(), rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("name", ["vae_nb_sampled_kl_train", "vae_nb_unit_variance_train"])
def test_sampled_kl_arithmetic_matches_reference_graph(name):
    """Kernel arithmetic on the posterior pre-activations of the golden cases -> the ELBO terms
    the reference's own graph code produced."""
    meta, groups = load_case(name)
    cfg = oracle_config(meta)
    assert not cfg.analytical_kl_term
    params, x, eps, features = oracle_inputs(meta, groups)
    out = O.vae_forward(cfg, params, x, x, eps, is_training=True)
    R, S, B, L = meta["R"], meta["S"], x.shape[0], cfg.latent_size
    unit_variance = cfg.latent_distribution == "unit-variance gaussian"
    # the posterior pre-activations (the oracle returns the clipped log_sigma; all golden values
    # lie inside the clip, so the clipped values serve as the raw ones)
    ph = out["q_z_mean"].numpy() if unit_variance else numpy.concatenate(
        [out["q_z_mean"].numpy(), out["log_sigma"].numpy()], axis=1)
    kl_rows, kl_elem = sampled_kl_rows(ph, eps.reshape(R * S * B, L).numpy(), B, L, R * S,
                                       unit_variance)
    weight = float(groups["in_feed"]["warm_up_weight"]) * cfg.kl_weight
    bound, _ = bound_rows(out["log_p_x_given_z"].reshape(-1).numpy(), kl_rows, R, S, B, weight)
    want = groups["out"]
    for i, key in enumerate(["lower_bound", "lower_bound_weighted", "reconstruction_error",
                             "kl_divergence"]):
        assert math.isclose(bound[i], float(want[key]), rel_tol=1e-9, abs_tol=1e-10), key
    assert numpy.allclose(kl_elem.mean(axis=0), want["kl_divergence_neurons"], rtol=1e-9,
                          atol=1e-10)
