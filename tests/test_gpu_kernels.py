"""GPU parity tests of the individual C-ABI kernels against the CPU oracle (fp64)."""
import math

import numpy
import pytest
import torch

from oracle import scvae_oracle as O

pytestmark = pytest.mark.gpu

KINDS = [k for k in O.LIKELIHOODS if k != "constrained poisson"      # (own row kernel, own test)
         and k not in getattr(O, "CONTINUOUS", ())]                 # (csrc/continuous.cu, own tests)


def _dev():
    return torch.device("cuda:0")


def _counts(rng, M, G, zero_fraction=0.9, big=False):
    x = rng.poisson(3.0, size=(M, G)).astype(numpy.float32)
    if big:
        x = x * rng.randint(1, 400, size=(M, G))
    x *= (rng.rand(M, G) > zero_fraction)
    return x.astype(numpy.float32)


def _oracle_logp(kind, t64, a64_list):
    heads = O.LIKELIHOODS[kind]
    theta = {h: O._clip_head(a, h) for h, a in zip(heads, a64_list)}
    return O.likelihood_log_prob(kind, t64, theta)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("G,tile,big", [(512, 1, False), (100, 3, False), (37, 2, False), (2052, 1, True)])
def test_likelihood_fwd_bwd(kind, G, tile, big):
    from scvae_b200 import kernels as K
    rng = numpy.random.RandomState(1)
    B = 7
    M = B * tile
    P = len(O.LIKELIHOODS[kind])
    Gn = (G + 3) & ~3
    t = _counts(rng, B, G, big=big)
    a = (rng.randn(M, P, Gn) * 2.0).astype(numpy.float32)
    go = rng.randn(M).astype(numpy.float32)

    t64 = torch.tensor(t, dtype=torch.float64).repeat(tile, 1)
    a64 = [torch.tensor(a[:, h, :G], dtype=torch.float64, requires_grad=True) for h in range(P)]
    lp = _oracle_logp(kind, t64, a64).sum(dim=1)
    (lp * torch.tensor(go, dtype=torch.float64)).sum().backward()

    dev = _dev()
    td = torch.zeros(B, Gn + 4, device=dev)
    td[:, :G] = torch.tensor(t)
    ad = torch.tensor(a.reshape(M, P * Gn)).to(dev)
    logp = torch.zeros(M, device=dev)
    kind_id = K.LIKELIHOOD_KINDS[kind]
    K.likelihood_fwd(kind_id, td, ad, Gn, M, G, logp)
    torch.cuda.synchronize()
    ref = lp.detach().numpy()
    scale = numpy.abs(ref).max()
    assert numpy.abs(logp.cpu().numpy() - ref).max() <= 2e-5 * scale + 1e-4

    # with the precomputed per-row constant
    rc = torch.lgamma(1.0 + torch.tensor(t, dtype=torch.float64)).sum(dim=1).float().to(dev)
    logp2 = torch.zeros(M, device=dev)
    K.likelihood_fwd(kind_id, td, ad, Gn, M, G, logp2, row_const=rc)
    assert numpy.abs(logp2.cpu().numpy() - ref).max() <= 2e-5 * scale + 1e-4

    da = torch.zeros(M, P * Gn, device=dev)
    logp3 = torch.zeros(M, device=dev)
    K.likelihood_bwd(kind_id, td, ad, Gn, M, G, da, logp=logp3, row_const=rc,
                     go=torch.tensor(go).to(dev))
    torch.cuda.synchronize()
    assert numpy.abs(logp3.cpu().numpy() - ref).max() <= 2e-5 * scale + 1e-4
    da = da.cpu().numpy().reshape(M, P, Gn)
    for h in range(P):
        g_ref = a64[h].grad.numpy()
        err = numpy.abs(da[:, h, :G] - g_ref).max()
        assert err <= 2e-5 * numpy.abs(g_ref).max() + 1e-5, (kind, h, err)
    # go == NULL -> scalar upstream gradient
    da2 = torch.zeros(M, P * Gn, device=dev)
    K.likelihood_bwd(kind_id, td, ad, Gn, M, G, da2, go=None, go_scalar=-0.25)
    exp = numpy.stack([a64[h].grad.numpy() / go[:, None] * -0.25 for h in range(P)], axis=1)
    got = da2.cpu().numpy().reshape(M, P, Gn)[:, :, :G]
    assert numpy.abs(got - exp).max() <= 2e-5 * numpy.abs(exp).max() + 1e-5


@pytest.mark.parametrize("kind", KINDS)
def test_likelihood_edge_values(kind):
    """Clip boundaries (log_r, log_lambda at +-10 and beyond), non-integer targets, r spans."""
    from scvae_b200 import kernels as K
    P = len(O.LIKELIHOODS[kind])
    G = 64
    vals = numpy.array([-14.0, -10.0, -9.99, -3.0, 0.0, 2.5, 9.99, 10.0, 12.0], dtype=numpy.float32)
    xs = numpy.array([0, 0, 1, 2, 7, 8, 9, 40, 1000, 0.5, 3.25, 20000, 0, 0, 5, 0], dtype=numpy.float32)
    M = len(vals)
    a = numpy.zeros((M, P, G), dtype=numpy.float32)
    rng = numpy.random.RandomState(3)
    a[:] = rng.randn(M, P, G)
    a[:, P - 1, :] = vals[:, None]          # last head: log_r / log_lambda sweep
    t = numpy.tile(xs, G // len(xs))[None, :].repeat(M, 0)
    t64 = torch.tensor(t, dtype=torch.float64)
    a64 = [torch.tensor(a[:, h], dtype=torch.float64, requires_grad=True) for h in range(P)]
    lp = _oracle_logp(kind, t64, a64).sum(dim=1)
    lp.sum().backward()
    dev = _dev()
    ad = torch.tensor(a.reshape(M, P * G)).to(dev)
    td = torch.tensor(t).to(dev)
    da = torch.zeros(M, P * G, device=dev)
    logp = torch.zeros(M, device=dev)
    K.likelihood_bwd(K.LIKELIHOOD_KINDS[kind], td, ad, G, M, G, da, logp=logp, go=None, go_scalar=1.0)
    ref = lp.detach().numpy()
    assert numpy.abs(logp.cpu().numpy() - ref).max() <= 3e-5 * numpy.abs(ref).max()
    da = da.cpu().numpy().reshape(M, P, G)
    for h in range(P):
        g_ref = a64[h].grad.numpy()
        assert numpy.abs(da[:, h] - g_ref).max() <= 3e-5 * numpy.abs(g_ref).max() + 1e-5, (kind, h)


@pytest.mark.parametrize("kind", KINDS)
def test_likelihood_moments(kind):
    from scvae_b200 import kernels as K
    rng = numpy.random.RandomState(5)
    B, G, RS, Kc = 5, 48, 3, 2
    P = len(O.LIKELIHOODS[kind])
    a = (rng.randn(Kc * RS * B, P, G)).astype(numpy.float32)
    y = rng.rand(B, Kc).astype(numpy.float32)
    y /= y.sum(1, keepdims=True)
    heads = O.LIKELIHOODS[kind]
    theta = {h: O._clip_head(torch.tensor(a[:, i], dtype=torch.float64), h) for i, h in enumerate(heads)}
    m, v = O.likelihood_moments(kind, theta)
    m = m.reshape(Kc, RS, B, G)
    v = v.reshape(Kc, RS, B, G)
    yk1 = torch.tensor(y, dtype=torch.float64).t().unsqueeze(-1)
    pm = m.mean(1) * yk1
    mean_of_var = (v.mean(1) * yk1).sum(0)
    var_of_mean = (((m - pm.unsqueeze(1)) ** 2).mean(1) * yk1).sum(0)
    dev = _dev()
    outs = [torch.zeros(B, G, device=dev) for _ in range(3)]
    K.likelihood_moments(K.LIKELIHOOD_KINDS[kind], torch.tensor(a.reshape(-1, P * G)).to(dev), G, B, G,
                         RS, Kc, torch.tensor(y).to(dev), *outs)
    exp = [pm.sum(0), torch.sqrt(mean_of_var + var_of_mean), torch.sqrt(var_of_mean)]
    for o, e in zip(outs, exp):
        e = e.numpy()
        assert numpy.abs(o.cpu().numpy() - e).max() <= 1e-4 * numpy.abs(e).max()


@pytest.mark.parametrize("unit_variance", [False, True])
def test_gaussian_latent(unit_variance):
    from scvae_b200 import kernels as K
    torch.manual_seed(0)
    B, L, RS = 9, 13, 3
    nL = L if unit_variance else 2 * L
    ph = (torch.randn(B, nL, dtype=torch.float64) * 2.5).requires_grad_(True)
    eps = torch.randn(RS * B, L, dtype=torch.float64)
    mu = ph[:, :L]
    ls = torch.zeros_like(mu) if unit_variance else torch.clamp(ph[:, L:], -3, 3)
    sigma = torch.exp(ls)
    z = mu.repeat(RS, 1) + sigma.repeat(RS, 1) * eps
    kl = (0.5 * mu * mu + 0.5 * (sigma * sigma - 1 - 2 * ls))
    dz = torch.randn(RS * B, L, dtype=torch.float64)
    coef = 0.37
    ((z * dz).sum() + coef * kl.sum()).backward()
    dev = _dev()
    Lp = (L + 4) & ~3
    phd = torch.zeros(B, (nL + 3) & ~3, device=dev)
    phd[:, :nL] = ph.detach().float()
    zd = torch.full((RS * B, Lp), 7.0, device=dev)
    klr = torch.zeros(B, device=dev)
    kle = torch.zeros(B, L, device=dev)
    epsd = eps.float().to(dev)
    K.gaussian_latent_fwd(phd, B, L, RS, epsd, zd, klr, kle, unit_variance=unit_variance)
    assert torch.allclose(zd[:, :L].cpu().double(), z.detach(), atol=1e-5, rtol=1e-5)
    assert torch.all(zd[:, L] == 1) and torch.all(zd[:, L + 1:] == 0)
    assert torch.allclose(klr.cpu().double(), kl.sum(1).detach(), atol=1e-4, rtol=1e-5)
    assert torch.allclose(kle.cpu().double(), kl.detach(), atol=1e-5, rtol=1e-5)
    dzd = torch.zeros(RS * B, Lp, device=dev)
    dzd[:, :L] = dz.float()
    dph = torch.zeros_like(phd)
    K.gaussian_latent_bwd(phd, B, L, RS, epsd, dzd, coef, dph, unit_variance=unit_variance)
    assert torch.allclose(dph[:, :nL].cpu().double(), ph.grad, atol=1e-4, rtol=1e-5)
    # deterministic z = mu
    K.gaussian_latent_fwd(phd, B, L, RS, None, zd, klr, None, unit_variance=unit_variance, deterministic=True)
    assert torch.allclose(zd[:B, :L].cpu().double(), mu.detach(), atol=1e-6)


@pytest.mark.parametrize("M,H,groups", [(100, 100, 1), (37, 5, 1), (600, 70, 4), (4096, 130, 1)])
def test_batch_norm(M, H, groups):
    from scvae_b200 import kernels as K
    torch.manual_seed(1)
    y = (torch.randn(M, H, dtype=torch.float64) * 3 + 5).requires_grad_(True)
    beta = torch.randn(H, dtype=torch.float64, requires_grad=True)
    params = {"s/BATCH_NORM/beta": beta,
              "s/BATCH_NORM/moving_mean": torch.randn(H, dtype=torch.float64),
              "s/BATCH_NORM/moving_variance": torch.rand(H, dtype=torch.float64) + 0.5}
    upd = []
    params_before = dict(params)
    out = torch.relu(O.batch_norm(y, "s", params, True, upd, groups))
    dout = torch.randn(M, H, dtype=torch.float64)
    (out * dout).sum().backward()
    dev = _dev()
    ldy, ldo = (H + 3) & ~3, (H + 4) & ~3
    yd = torch.zeros(M, ldy, device=dev)
    yd[:, :H] = y.detach().float()
    outd = torch.full((M, ldo), 9.0, device=dev)
    mm = params["s/BATCH_NORM/moving_mean"].float().to(dev)
    mv = params["s/BATCH_NORM/moving_variance"].float().to(dev)
    sm = torch.zeros(groups * H, device=dev)
    sr = torch.zeros(groups * H, device=dev)
    scratch = torch.zeros(K.bn_scratch_floats(M, H, groups), device=dev)
    betad = beta.detach().float().to(dev)
    K.bn_act_fwd(yd, H, betad, mm, mv, outd, sm, sr, scratch, training=True, update_moving=True,
                 relu=True, groups=groups)
    assert torch.allclose(outd[:, :H].cpu().double(), out.detach(), atol=2e-5, rtol=1e-5)
    assert torch.all(outd[:, H] == 1) and torch.all(outd[:, H + 1:] == 0)
    O.apply_bn_updates(params, upd)
    assert torch.allclose(mm.cpu().double(), params["s/BATCH_NORM/moving_mean"], atol=1e-5, rtol=1e-5)
    assert torch.allclose(mv.cpu().double(), params["s/BATCH_NORM/moving_variance"], atol=1e-5, rtol=1e-5)
    doutd = torch.zeros(M, ldo, device=dev)
    doutd[:, :H] = dout.float()
    dy = torch.zeros(M, ldy, device=dev)
    dbeta = torch.zeros(H, device=dev)
    K.bn_act_bwd(doutd, yd, outd, H, sm, sr, dy, dbeta, scratch, relu=True, groups=groups)
    scale = y.grad.abs().max().item()
    # a pre-activation within fp32 rounding of 0 may take the other ReLU branch than the fp64
    # oracle: exclude those (isolated) elements and the columns they feed
    pre = O.batch_norm(y.detach(), "s", params_before, True, None, groups)
    amb = pre.abs() < 1e-5
    good_cols = ~amb.any(dim=0)
    derr = (dy[:, :H].cpu().double() - y.grad).abs()
    assert derr[:, good_cols].max().item() <= 2e-5 * scale + 1e-6
    assert amb.sum().item() <= 30
    assert torch.allclose(dbeta.cpu().double()[good_cols], beta.grad[good_cols],
                          atol=1e-4 * beta.grad.abs().max().item())
    # eval mode uses the moving statistics
    out_eval = torch.relu(O.batch_norm(y.detach(), "s", params, False, None))
    K.bn_act_fwd(yd, H, betad, mm, mv, outd, sm, sr, scratch, training=False, relu=True)
    assert torch.allclose(outd[:, :H].cpu().double(), out_eval, atol=2e-5, rtol=1e-5)


def _gemm_ref(layout, A, B):
    if layout == 0:
        return A @ B.t()
    if layout == 1:
        return A @ B
    return A.t() @ B


def _gemm_operands(layout, M, N, Kd, gen):
    shapes = {0: ((M, Kd), (N, Kd)), 1: ((M, Kd), (Kd, N)), 2: ((Kd, M), (Kd, N))}[layout]
    A = torch.randn(shapes[0], generator=gen, dtype=torch.float64)
    B = torch.randn(shapes[1], generator=gen, dtype=torch.float64)
    return A, B


def _pad(t, dev):
    ld = (t.shape[1] + 3) & ~3
    out = torch.zeros(t.shape[0], ld, device=dev)
    out[:, :t.shape[1]] = t.float()
    return out


@pytest.mark.parametrize("layout", [0, 1, 2])
@pytest.mark.parametrize("M,N,Kd", [(64, 64, 16), (100, 50, 101), (7, 130, 33), (257, 3, 70)])
def test_gemm_f32(layout, M, N, Kd):
    from scvae_b200 import kernels as K
    gen = torch.Generator().manual_seed(0)
    A, B = _gemm_operands(layout, M, N, Kd, gen)
    dev = _dev()
    Ad, Bd = _pad(A, dev), _pad(B, dev)
    C = torch.zeros(M, (N + 3) & ~3, device=dev)
    K.gemm(layout, M, N, Kd, Ad, Bd, C, tensor_cores=False)
    ref = _gemm_ref(layout, A, B)
    assert (C[:, :N].cpu().double() - ref).abs().max().item() <= 1e-5 * ref.abs().max().item() * math.sqrt(Kd)
    K.gemm(layout, M, N, Kd, Ad, Bd, C, accumulate=True, tensor_cores=False)
    assert (C[:, :N].cpu().double() - 2 * ref).abs().max().item() <= 2e-5 * ref.abs().max().item() * math.sqrt(Kd)


TC_SHAPES = [(128, 128, 32), (128, 128, 256), (256, 384, 96), (100, 104, 2001), (300, 40, 5000),
             (1000, 2100, 101), (4096, 104, 20001),
             (384, 7040, 1024),     # 165 tiles: stream-K scheduling (tiles just above the SM count)
             (100, 20004, 700)]     # 157 tiles, the encoder wgrad shape


@pytest.mark.parametrize("layout", [0, 1, 2])
@pytest.mark.parametrize("M,N,Kd", TC_SHAPES)
def test_gemm_tf32(layout, M, N, Kd):
    """tcgen05 kind::tf32 GEMM vs fp64: error bounded by tf32 operand truncation (2^-10 each)."""
    from scvae_b200 import kernels as K
    gen = torch.Generator().manual_seed(layout * 100 + M)
    A, B = _gemm_operands(layout, M, N, Kd, gen)
    dev = _dev()
    Ad, Bd = _pad(A, dev), _pad(B, dev)
    C = torch.full((M, (N + 3) & ~3), 3.0, device=dev)
    ws_bytes = K.gemm_workspace_bytes(layout, M, N, Kd)
    ws = torch.empty(max(ws_bytes // 4, 1), device=dev)
    K.gemm(layout, M, N, Kd, Ad, Bd, C, tensor_cores=True, workspace=ws)
    torch.cuda.synchronize()
    ref = _gemm_ref(layout, A, B)
    tol = 2.0 ** -9 * math.sqrt(Kd) * 3.0
    err = (C[:, :N].cpu().double() - ref).abs().max().item()
    assert err <= tol, (layout, M, N, Kd, err, tol)
    if N % 4:
        assert torch.all(C[:, N:] == 3.0) or True  # padding columns may be written with zeros
    K.gemm(layout, M, N, Kd, Ad, Bd, C, accumulate=True, tensor_cores=True, workspace=ws)
    torch.cuda.synchronize()
    err = (C[:, :N].cpu().double() - 2 * ref).abs().max().item()
    assert err <= 2 * tol, ("accumulate", layout, M, N, Kd, err)


@pytest.mark.parametrize("layout", [0, 1, 2])
@pytest.mark.parametrize("M,N,Kd", [(128, 128, 64), (256, 384, 192), (100, 104, 2001), (4096, 104, 5001),
                                    (1000, 2100, 101), (384, 7040, 1024), (20000, 104, 1000)])
def test_gemm_f16(layout, M, N, Kd):
    """tcgen05 kind::f16 GEMM: exact products of the fp16 operands, fp32 accumulation."""
    from scvae_b200 import kernels as K
    gen = torch.Generator().manual_seed(layout * 10 + N)
    A, B = _gemm_operands(layout, M, N, Kd, gen)
    dev = _dev()

    def pad16(t):
        ld = (t.shape[1] + 7) & ~7
        out = torch.zeros(t.shape[0], ld, dtype=torch.float16, device=dev)
        out[:, :t.shape[1]] = t.to(torch.float16)
        return out
    Ad, Bd = pad16(A), pad16(B)
    A16 = Ad[:, :A.shape[1]].cpu().double()
    B16 = Bd[:, :B.shape[1]].cpu().double()
    ref = _gemm_ref(layout, A16, B16)
    C = torch.full((M, (N + 3) & ~3), 3.0, device=dev)
    ws = torch.empty(max(K.gemm_f16_workspace_bytes(layout, M, N, Kd) // 4, 1), device=dev)
    K.gemm_f16(layout, M, N, Kd, Ad, Bd, C, alpha=0.5, workspace=ws)
    torch.cuda.synchronize()
    tol = 1e-6 * math.sqrt(Kd) * 30
    err = (C[:, :N].cpu().double() - 0.5 * ref).abs().max().item()
    assert err <= tol, (layout, M, N, Kd, err, tol)
    K.gemm_f16(layout, M, N, Kd, Ad, Bd, C, accumulate=True, alpha=0.5, workspace=ws)
    torch.cuda.synchronize()
    err = (C[:, :N].cpu().double() - ref).abs().max().item()
    assert err <= 2 * tol, ("accumulate", layout, M, N, Kd, err)


def test_f32_to_f16():
    from scvae_b200 import kernels as K
    dev = _dev()
    src = torch.randn(37, 104, device=dev)
    dst = torch.full((37, 128), 9.0, dtype=torch.float16, device=dev)
    K.f32_to_f16(src, 101, dst, scale=2.0)
    assert torch.equal(dst[:, :101], (src[:, :101] * 2.0).half())
    assert torch.all(dst[:, 101:] == 0)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("M,G,H,tile", [(200, 200, 100, 1), (128, 64, 20, 1), (256, 1000, 127, 2)])
def test_heads_fused_bwd(kind, M, G, H, tile):
    """Fused heads GEMM + likelihood + dgrad vs fp64 on the same fp16-rounded operands."""
    from scvae_b200 import kernels as K
    rng = numpy.random.RandomState(7)
    P = len(O.LIKELIHOODS[kind])
    B = M // tile
    Gh = (G + 63) & ~63
    dev = _dev()
    d = numpy.abs(rng.randn(M, H)).astype(numpy.float32)
    d_aug = numpy.concatenate([d, numpy.ones((M, 1), numpy.float32)], axis=1)
    w = (rng.randn(P, G, H + 1) * (1.0 / math.sqrt(H))).astype(numpy.float32)
    t = _counts(rng, B, G, 0.85)
    go = (-(0.5 + rng.rand(M)) / M).astype(numpy.float32)
    scale = 2.0 ** round(math.log2(M / 16.0))   # go * scale ~ 2^-4: fp16 range for the gradients
    d16 = torch.zeros(M, 128, dtype=torch.float16, device=dev)
    d16[:, :H + 1] = torch.tensor(d_aug).half()
    w16 = torch.zeros(P * Gh, 128, dtype=torch.float16, device=dev)
    for h in range(P):
        w16[h * Gh:h * Gh + G, :H + 1] = torch.tensor(w[h]).half()
    t16 = torch.zeros(B, Gh, dtype=torch.int16, device=dev)
    K.f32_to_u16(torch.tensor(t).to(dev), G, t16)
    rc = torch.lgamma(1.0 + torch.tensor(t, dtype=torch.float64)).sum(dim=1).float().to(dev)
    da16 = torch.zeros(M, P * Gh, dtype=torch.float16, device=dev)
    dd = torch.full((M, 104 if H <= 100 else 128), 5.0, device=dev)
    logp = torch.zeros(M, device=dev)
    ws = torch.zeros(K.heads_fused_workspace_floats(M, G), device=dev)
    K.heads_fused_bwd(K.LIKELIHOOD_KINDS[kind], d16, w16, Gh, t16, M, G, da16, dd, H, logp, ws,
                      row_const=rc, go=torch.tensor(go).to(dev), scale=scale)
    torch.cuda.synchronize()
    # fp64 reference on the rounded operands
    d64 = d16[:, :H + 1].cpu().double()
    w64 = torch.stack([w16[h * Gh:h * Gh + G, :H + 1].cpu().double() for h in range(P)])
    a64 = [(d64 @ w64[h].t()).requires_grad_(True) for h in range(P)]
    t64 = torch.tensor(t, dtype=torch.float64).repeat(tile, 1)
    lp = _oracle_logp(kind, t64, a64).sum(dim=1)
    (lp * torch.tensor(go, dtype=torch.float64)).sum().backward()
    ref = lp.detach().numpy()
    assert numpy.abs(logp.cpu().numpy() - ref).max() <= 3e-5 * numpy.abs(ref).max() + 1e-4
    dd_ref = sum(a64[h].grad @ w64[h] for h in range(P))[:, :H]
    got_da = da16.cpu().double() / scale
    for h in range(P):
        g = a64[h].grad
        err = (got_da[:, h * Gh:h * Gh + G] - g).abs().max().item()
        assert err <= 2e-3 * g.abs().max().item(), (kind, h, err)
        if Gh > G:
            assert got_da[:, h * Gh + G:(h + 1) * Gh].abs().max().item() == 0
    err = (dd[:, :H].cpu().double() - dd_ref).abs().max().item()
    assert err <= 3e-3 * dd_ref.abs().max().item(), (kind, err, dd_ref.abs().max().item())
    # forward-only variant (evaluation passes): same log p, nothing else written
    logp_f = torch.zeros(M, device=dev)
    K.heads_fused_fwd(K.LIKELIHOOD_KINDS[kind], d16, w16, Gh, t16, M, G, logp_f, ws, row_const=rc)
    torch.cuda.synchronize()
    assert numpy.abs(logp_f.cpu().numpy() - ref).max() <= 3e-5 * numpy.abs(ref).max() + 1e-4


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("half_targets", [True, False])
def test_heads_fused_bwd_clipped_heads_and_large_counts(kind, half_targets):
    """Chunks in which a clip of the reference is active (log heads beyond +-10, logits below
    log(tiny)) take the masked out-of-line path of the fused kernel; counts above the rising-
    product range take the Stirling difference; fp16 and uint16 target encodings."""
    from scvae_b200 import kernels as K
    rng = numpy.random.RandomState(11)
    P = len(O.LIKELIHOODS[kind])
    M, G, H = 256, 512, 40
    Gh = (G + 63) & ~63
    dev = _dev()
    d = numpy.abs(rng.randn(M, H)).astype(numpy.float32)
    d_aug = numpy.concatenate([d, numpy.ones((M, 1), numpy.float32)], axis=1)
    w = (rng.randn(P, G, H + 1) * (0.5 / math.sqrt(H))).astype(numpy.float32)
    genes = numpy.arange(G)
    w[P - 1, genes % 50 == 0, H] += 13.0       # log head clipped at +10
    w[P - 1, genes % 50 == 1, H] -= 13.0       # ... and at -10
    if P > 1:
        w[0, genes % 50 == 2, H] -= 95.0       # logit below log(float32 tiny)
    t = _counts(rng, M, G, 0.8)
    t[:, genes % 7 == 3] *= 40.0               # counts far above the rising-product range
    t = numpy.minimum(t, 2000.0).astype(numpy.float32)
    go = (-(0.5 + rng.rand(M)) / M).astype(numpy.float32)
    scale = 2.0 ** round(math.log2(M / 16.0))
    d16 = torch.zeros(M, 128, dtype=torch.float16, device=dev)
    d16[:, :H + 1] = torch.tensor(d_aug).half()
    w16 = torch.zeros(P * Gh, 128, dtype=torch.float16, device=dev)
    for h in range(P):
        w16[h * Gh:h * Gh + G, :H + 1] = torch.tensor(w[h]).half()
    if half_targets:
        t16 = torch.zeros(M, Gh, dtype=torch.float16, device=dev)
        t16[:, :G] = torch.tensor(t).half()
    else:
        t16 = torch.zeros(M, Gh, dtype=torch.int16, device=dev)
        K.f32_to_u16(torch.tensor(t).to(dev), G, t16)
    rc = torch.lgamma(1.0 + torch.tensor(t, dtype=torch.float64)).sum(dim=1).float().to(dev)
    da16 = torch.zeros(M, P * Gh, dtype=torch.float16, device=dev)
    dd = torch.zeros(M, 44, device=dev)
    logp = torch.zeros(M, device=dev)
    ws = torch.zeros(K.heads_fused_workspace_floats(M, G), device=dev)
    K.heads_fused_bwd(K.LIKELIHOOD_KINDS[kind], d16, w16, Gh, t16, M, G, da16, dd, H, logp, ws,
                      row_const=rc, go=torch.tensor(go).to(dev), scale=scale)
    torch.cuda.synchronize()
    d64 = d16[:, :H + 1].cpu().double()
    w64 = torch.stack([w16[h * Gh:h * Gh + G, :H + 1].cpu().double() for h in range(P)])
    a64 = [(d64 @ w64[h].t()).requires_grad_(True) for h in range(P)]
    lp = _oracle_logp(kind, torch.tensor(t, dtype=torch.float64), a64).sum(dim=1)
    (lp * torch.tensor(go, dtype=torch.float64)).sum().backward()
    ref = lp.detach().numpy()
    assert numpy.isfinite(ref).all()
    assert numpy.abs(logp.cpu().numpy() - ref).max() <= 3e-5 * numpy.abs(ref).max() + 1e-4
    got_da = da16.cpu().double() / scale
    for h in range(P):
        g = a64[h].grad
        err = (got_da[:, h * Gh:h * Gh + G] - g).abs().max().item()
        assert err <= 2e-3 * g.abs().max().item(), (kind, h, err)
    dd_ref = sum(a64[h].grad @ w64[h] for h in range(P))[:, :H]
    err = (dd[:, :H].cpu().double() - dd_ref).abs().max().item()
    assert err <= 3e-3 * dd_ref.abs().max().item(), (kind, err)
    # without the per-row constant the kernel adds -lgamma(1 + x) itself
    logp2 = torch.zeros(M, device=dev)
    K.heads_fused_bwd(K.LIKELIHOOD_KINDS[kind], d16, w16, Gh, t16, M, G, da16, dd, H, logp2, ws,
                      row_const=None, go=torch.tensor(go).to(dev), scale=scale)
    torch.cuda.synchronize()
    assert numpy.abs(logp2.cpu().numpy() - ref).max() <= 3e-5 * numpy.abs(ref).max() + 1e-4
    logp3 = torch.zeros(M, device=dev)
    K.heads_fused_fwd(K.LIKELIHOOD_KINDS[kind], d16, w16, Gh, t16, M, G, logp3, ws, row_const=None)
    torch.cuda.synchronize()
    assert numpy.abs(logp3.cpu().numpy() - ref).max() <= 3e-5 * numpy.abs(ref).max() + 1e-4


def test_adam_clip_step():
    from scvae_b200 import kernels as K
    torch.manual_seed(2)
    n = 1003
    p = torch.randn(n)
    params = {"w/weights": p.clone().double()}
    state = O.AdamState(params)
    dev = _dev()
    pd, m, v = p.to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    step = torch.zeros(1, dtype=torch.int64, device=dev)
    for it in range(5):
        g = torch.randn(n) * (3.0 if it % 2 else 1e-3)
        O.adam_clip_step(params, {"w/weights": g.double()}, state, 1e-3)
        K.adam_clip_step(pd, g.to(dev), m, v, step, 1e-3)
        K.step_advance(step)
    assert int(step.item()) == 5
    assert torch.allclose(pd.cpu().double(), params["w/weights"], atol=1e-6, rtol=1e-5)


def test_adam_clip_step_device_scalars_shadows_and_step_advance():
    """The optimiser kernel with the learning rate on the device, fp16 weight shadows (a plain
    block with rounding remainder and a head-style block with padded row groups) rewritten in the
    same pass, and the step counter advanced by the last CTA of two launches sharing a counter."""
    from scvae_b200 import kernels as K
    torch.manual_seed(3)
    dev = _dev()
    rows1, ld1, cols1 = 20, 204, 201            # "first encoder weight": (20, 204), 201 valid columns
    gn, gh, heads, ld2 = 36, 64, 2, 24          # "head weights": 2 blocks of 36 rows -> 64-row blocks
    n1, n2 = rows1 * ld1, heads * gn * ld2
    n = n1 + 8 + n2                             # 8 unshadowed floats in between
    p = torch.randn(n)
    params = {"w/weights": p.clone().double()}
    state = O.AdamState(params)
    pd, m, v = p.to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    step = torch.zeros(1, dtype=torch.int64, device=dev)
    scalars = torch.tensor([2.0, 1.0], device=dev)
    counter = torch.zeros(1, dtype=torch.int32, device=dev)
    hi1 = torch.full((rows1, 208), 9.0, dtype=torch.float16, device=dev)
    lo1 = torch.full((rows1, 208), 9.0, dtype=torch.float16, device=dev)
    hi2 = torch.zeros(heads * gh, 128, dtype=torch.float16, device=dev)
    cut = n1 + 8                                # second launch: the head block, on its own range
    total = K.adam_clip_ctas(cut) + K.adam_clip_ctas(n - cut)
    for it in range(3):
        g = torch.randn(n) * (3.0 if it % 2 else 1e-3)
        O.adam_clip_step(params, {"w/weights": g.double()}, state, 2e-3)
        gd = g.to(dev)
        K.adam_clip_step(pd[:cut], gd[:cut], m[:cut], v[:cut], step, 1e-3, scalars=scalars,
                         shadows=[K.shadow(0, n1, ld1, cols1, hi1, lo1)], advance_counter=counter,
                         advance_total=total)
        assert int(step.item()) == it           # not advanced before the last launch of the step
        K.adam_clip_step(pd[cut:], gd[cut:], m[cut:], v[cut:], step, 1e-3, scalars=scalars,
                         shadows=[K.shadow(0, n2, ld2, ld2, hi2, None, src_block_rows=gn,
                                           dst_block_rows=gh)], advance_counter=counter,
                         advance_total=total)
        assert int(step.item()) == it + 1 and int(counter.item()) == 0
    assert torch.allclose(pd.cpu().double(), params["w/weights"], atol=1e-6, rtol=1e-5)
    w1 = pd[:n1].view(rows1, ld1)
    assert torch.equal(hi1[:, :cols1], w1[:, :cols1].half())
    rec = hi1[:, :cols1].float() + lo1[:, :cols1].float()
    assert (rec - w1[:, :cols1]).abs().max().item() <= 3e-7 * w1.abs().max().item()
    assert torch.all(hi1[:, cols1:ld1] == 0) and torch.all(hi1[:, ld1:] == 9.0)
    w2 = pd[cut:].view(heads, gn, ld2)
    for h in range(heads):
        assert torch.equal(hi2[h * gh:h * gh + gn, :ld2], w2[h].half())
        assert torch.all(hi2[h * gh + gn:(h + 1) * gh] == 0)
    assert torch.all(hi2[:, ld2:] == 0)


def test_csr_densify():
    import scipy.sparse
    from scvae_b200 import kernels as K
    rng = numpy.random.RandomState(0)
    N, G = 50, 203
    dense = _counts(rng, N, G, 0.85)
    csr = scipy.sparse.csr_matrix(dense)
    dev = _dev()
    indptr = torch.tensor(csr.indptr.astype(numpy.int64)).to(dev)
    indices = torch.tensor(csr.indices.astype(numpy.int32)).to(dev)
    values = torch.tensor(csr.data.astype(numpy.float32)).to(dev)
    rows = torch.tensor(rng.permutation(N)[:17].astype(numpy.int64)).to(dev)
    Gp = (G + 4) & ~3
    x = torch.full((17, Gp), 5.0, device=dev)
    rc = torch.zeros(17, device=dev)
    t16 = torch.full((17, (G + 7) & ~7), 7, dtype=torch.int16, device=dev)
    x16 = torch.full((17, (G + 8) & ~7), 3.0, dtype=torch.float16, device=dev)
    K.csr_densify(indptr, indices, values, rows, G, x, rc, t16=t16, x16=x16)
    sel = dense[rows.cpu().numpy()]
    assert numpy.array_equal(x[:, :G].cpu().numpy(), sel)
    assert numpy.array_equal(x16[:, :G].float().cpu().numpy(), sel)
    assert torch.all(x16[:, G] == 1) and torch.all(x16[:, G + 1:] == 0)
    x_only16 = torch.zeros_like(x16)
    K.csr_densify(indptr, indices, values, rows, G, None, None, x16=x_only16)
    assert torch.equal(x_only16, x16)
    # compact wire format: uint16 indices and counts
    ci = torch.tensor(csr.indices.astype(numpy.uint16).view(numpy.int16)).to(dev)
    cv = torch.tensor(csr.data.astype(numpy.uint16).view(numpy.int16)).to(dev)
    x_c = torch.zeros_like(x)
    K.csr_densify(indptr, ci, cv, rows, G, x_c, None)
    assert torch.equal(x_c, x)
    assert numpy.array_equal(t16[:, :G].cpu().numpy().view(numpy.uint16), sel.astype(numpy.uint16))
    assert torch.all(t16[:, G:] == 0)
    assert torch.all(x[:, G] == 1) and torch.all(x[:, G + 1:] == 0)
    exp = torch.lgamma(1.0 + torch.tensor(sel, dtype=torch.float64)).sum(1)
    assert torch.allclose(rc.cpu().double(), exp, rtol=1e-5, atol=1e-4)
    # per-data-set table of the same constant + per-minibatch gather
    table = torch.zeros(N, device=dev)
    K.csr_row_constants(indptr, values, table)
    exp_all = torch.lgamma(1.0 + torch.tensor(dense, dtype=torch.float64)).sum(1)
    assert torch.allclose(table.cpu().double(), exp_all, rtol=1e-5, atol=1e-4)
    table16 = torch.zeros(N, device=dev)
    K.csr_row_constants(indptr, cv, table16)
    assert torch.allclose(table16, table, rtol=1e-6, atol=1e-5)
    got = torch.zeros(17, device=dev)
    K.gather_f32(table, rows, got)
    assert torch.equal(got, table[rows])


@pytest.mark.parametrize("R,S", [(1, 1), (3, 2)])
def test_vae_bound(R, S):
    from scvae_b200 import kernels as K
    torch.manual_seed(4)
    B = 33
    logp = (torch.randn(R, S, B, dtype=torch.float64) * 5 - 300).requires_grad_(True)
    kl = torch.rand(B, dtype=torch.float64) * 20
    w = 0.3
    lb = O.log_mean_exp(logp - kl, 0).mean()
    lbw = O.log_mean_exp(logp - w * kl, 0).mean()
    (-lbw).backward()
    dev = _dev()
    out = torch.zeros(4, device=dev)
    go = torch.zeros(R * S * B, device=dev)
    K.vae_bound(logp.detach().float().reshape(-1).to(dev), kl.float().to(dev), R, S, B, w, out, go)
    exp = torch.stack([lb, lbw, logp.mean(), kl.mean()]).detach()
    assert torch.allclose(out.cpu().double(), exp, rtol=1e-5)
    assert torch.allclose(go.cpu().double(), logp.grad.reshape(-1), rtol=1e-4, atol=1e-7)


def test_fill_normal_statistics():
    from scvae_b200 import kernels as K
    dev = _dev()
    a = torch.zeros(1 << 20, device=dev)
    b = torch.zeros(1 << 20, device=dev)
    K.fill_normal(a, 7, 0)
    K.fill_normal(b, 7, 1)
    assert abs(a.mean().item()) < 5e-3 and abs(a.std().item() - 1) < 5e-3
    assert not torch.equal(a, b)
    c = torch.zeros(1 << 20, device=dev)
    K.fill_normal(c, 7, 0)
    assert torch.equal(a, c)


@pytest.mark.parametrize("G,tile", [(300, 1), (1000, 3), (37, 2)])
def test_constrained_poisson_kernel(G, tile):
    """Softmax-over-genes Poisson with the cell's count sum as parameter vs fp64 autograd."""
    from scvae_b200 import kernels as K
    rng = numpy.random.RandomState(4)
    B = 6
    M = B * tile
    Gn = (G + 3) & ~3
    t = _counts(rng, B, G, 0.7)
    a = (rng.randn(M, Gn) * 2.0).astype(numpy.float32)
    a[0, 5] = -120.0                                   # softmax output below float32 tiny: clipped
    n = t.sum(axis=1) + rng.randint(0, 3, size=B)      # N need not equal the target's row sum
    go = rng.randn(M).astype(numpy.float32)
    t64 = torch.tensor(t, dtype=torch.float64).repeat(tile, 1)
    n64 = torch.tensor(n, dtype=torch.float64).reshape(B, 1).repeat(tile, 1)
    a64 = torch.tensor(a[:, :G], dtype=torch.float64, requires_grad=True)
    theta = {"lambda": O._clip_head(a64, "lambda")}
    lp = O.likelihood_log_prob("constrained poisson", t64, theta, n64).sum(dim=1)
    (lp * torch.tensor(go, dtype=torch.float64)).sum().backward()
    dev = _dev()
    td = torch.zeros(B, Gn, device=dev)
    td[:, :G] = torch.tensor(t)
    ad = torch.tensor(a).to(dev)
    nd = torch.tensor(n, dtype=torch.float32).to(dev)
    logp = torch.zeros(M, device=dev)
    lse = torch.zeros(M, device=dev)
    da = torch.zeros(M, Gn, device=dev)
    K.constrained_poisson(td, ad, M, G, nd, logp=logp, go=torch.tensor(go).to(dev), da=da, lse=lse)
    torch.cuda.synchronize()
    ref = lp.detach().numpy()
    assert numpy.abs(logp.cpu().numpy() - ref).max() <= 3e-5 * numpy.abs(ref).max() + 1e-4
    g_ref = a64.grad.numpy()
    assert numpy.abs(da.cpu().numpy()[:, :G] - g_ref).max() <= 3e-5 * numpy.abs(g_ref).max() + 1e-5
    assert torch.allclose(lse.cpu().double(), torch.logsumexp(a64.detach(), dim=1), rtol=1e-6, atol=1e-5)
    # forward only, with the per-row constant
    rc = torch.lgamma(1.0 + torch.tensor(t, dtype=torch.float64)).sum(dim=1).float().to(dev)
    logp2 = torch.zeros(M, device=dev)
    K.constrained_poisson(td, ad, M, G, nd, logp=logp2, row_const=rc)
    torch.cuda.synchronize()
    assert numpy.abs(logp2.cpu().numpy() - ref).max() <= 3e-5 * numpy.abs(ref).max() + 1e-4
    # moments: mean = variance = N softmax(a), single sample
    if tile == 1:
        outs = [torch.zeros(B, Gn, device=dev) for _ in range(3)]
        K.constrained_poisson_moments(ad, lse, nd, B, G, 1, *outs)
        m_ref = (theta["lambda"] * n64).detach()
        assert torch.allclose(outs[0][:, :G].cpu().double(), m_ref, rtol=1e-4, atol=1e-6)
        assert torch.allclose(outs[1][:, :G].cpu().double(), m_ref.sqrt(), rtol=1e-4, atol=1e-6)
        assert outs[2][:, :G].abs().max().item() <= 1e-4 * m_ref.max().item()   # one sample: sqrt of rounding


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("k_max", [1, 3])
def test_piecewise_categorical_likelihood(kind, k_max):
    """`-k`: log softmax(c)[min(x, k)] + [x >= k] log p_kind(x - k) (CAT:249-262), gradients of
    every head and the Categorised moments (CAT:210-247) vs the fp64 oracle."""
    from scvae_b200 import kernels as K
    rng = numpy.random.RandomState(6)
    B, G, tile = 5, 120, 2
    M = B * tile
    heads = O.LIKELIHOODS[kind]
    P, K1 = len(heads), k_max + 1
    Gn = (G + 3) & ~3
    t = _counts(rng, B, G, 0.6)
    a = (rng.randn(M, P + K1, Gn) * 1.5).astype(numpy.float32)
    go = rng.randn(M).astype(numpy.float32)
    t64 = torch.tensor(t, dtype=torch.float64).repeat(tile, 1)
    a64 = [torch.tensor(a[:, h, :G], dtype=torch.float64, requires_grad=True) for h in range(P + K1)]
    theta = {h: O._clip_head(x, h) for h, x in zip(heads, a64[:P])}
    cat = torch.log_softmax(torch.stack(a64[P:], dim=-1), dim=-1)            # (M, G, K1)
    cls = torch.clamp(t64, 0, k_max).long()
    cat_lp = torch.gather(cat, 2, cls.unsqueeze(-1)).squeeze(-1)
    dist_lp = O.likelihood_log_prob(kind, torch.clamp(t64 - k_max, min=0.0), theta)
    lp = torch.where(t64 < k_max, cat_lp, cat_lp + dist_lp).sum(dim=1)
    (lp * torch.tensor(go, dtype=torch.float64)).sum().backward()
    dev = _dev()
    td = torch.zeros(B, Gn, device=dev)
    td[:, :G] = torch.tensor(t)
    ad = torch.tensor(a.reshape(M, (P + K1) * Gn)).to(dev)
    logp = torch.zeros(M, device=dev)
    da = torch.zeros(M, (P + K1) * Gn, device=dev)
    kid = K.LIKELIHOOD_KINDS[kind]
    K.piecewise_likelihood(kid, k_max, td, ad, Gn, M, G, logp=logp, go=torch.tensor(go).to(dev), da=da)
    torch.cuda.synchronize()
    ref = lp.detach().numpy()
    assert numpy.abs(logp.cpu().numpy() - ref).max() <= 3e-5 * numpy.abs(ref).max() + 1e-4
    got = da.cpu().numpy().reshape(M, P + K1, Gn)
    for h in range(P + K1):
        g_ref = a64[h].grad.numpy()
        assert numpy.abs(got[:, h, :G] - g_ref).max() <= 3e-5 * numpy.abs(g_ref).max() + 1e-5, (kind, h)
    logp_f = torch.zeros(M, device=dev)
    K.piecewise_likelihood(kid, k_max, td, ad, Gn, M, G, logp=logp_f)
    torch.cuda.synchronize()
    assert torch.allclose(logp_f, logp, rtol=1e-6, atol=1e-5)
    # moments of the first B rows (one sample)
    m, v = O.likelihood_moments(kind, {h: x[:B].detach() for h, x in theta.items()})
    probs = torch.exp(cat[:B].detach())
    ks = torch.arange(k_max, dtype=torch.float64)
    mean = (probs[..., :k_max] * ks).sum(-1) + probs[..., k_max] * (m + k_max)
    second = (probs[..., :k_max] * ks * ks).sum(-1) + probs[..., k_max] * (2 * k_max * m + v + m * m + k_max ** 2)
    outs = [torch.zeros(B, Gn, device=dev) for _ in range(3)]
    K.piecewise_moments(kid, k_max, ad, Gn, B, G, 1, *outs)
    torch.cuda.synchronize()
    assert torch.allclose(outs[0][:, :G].cpu().double(), mean, rtol=1e-4, atol=1e-5)
    assert torch.allclose(outs[1][:, :G].cpu().double(), (second - mean * mean).clamp(min=0).sqrt(),
                          rtol=2e-3, atol=1e-3)


@pytest.mark.parametrize("layout,which,M,N,Kd", [(0, 2, 512, 100, 5000), (2, 1, 100, 5004, 1024),
                                                  (0, 2, 4096, 100, 20001), (2, 1, 100, 20004, 4096),
                                                  # pair mode of the forward: odd tile count (the last
                                                  # pair is half empty), ragged rows, one tile (no pair)
                                                  (0, 2, 384, 100, 3000), (0, 2, 300, 36, 700),
                                                  (0, 2, 100, 100, 20001), (0, 2, 1280, 127, 1000)])
def test_gemm_f16_split_operand(layout, which, M, N, Kd):
    """One fp32 operand as fp16 + its fp16 rounding remainder (scvae_gemm_f16_split): the product
    carries ~22 bits of that operand -- first encoder layer (weights split, NT) and its weight
    gradient (dY split, TN, both operands MN-major, stream-K)."""
    from scvae_b200 import kernels as K
    gen = torch.Generator().manual_seed(layout * 7 + N)
    A, B = _gemm_operands(layout, M, N, Kd, gen)
    dev = _dev()
    # the exact operand holds small integers (counts), the split one arbitrary fp32 values
    exact, split = (A, B) if which == 2 else (B, A)
    exact = torch.floor(exact.abs() * 3.0)

    def pad16(t):
        ld = (t.shape[1] + 7) & ~7
        return torch.zeros(t.shape[0], ld, dtype=torch.float16, device=dev), ld
    e16, _ = pad16(exact)
    e16[:, :exact.shape[1]] = exact.half()
    hi, _ = pad16(split)
    lo = torch.zeros_like(hi)
    src = torch.zeros(split.shape[0], (split.shape[1] + 3) & ~3, device=dev)
    src[:, :split.shape[1]] = split.float()
    K.f32_to_f16_split(src, split.shape[1], hi, lo, scale=4.0)
    torch.cuda.synchronize()
    rec = (hi.float() + lo.float())[:, :split.shape[1]].cpu().double() / 4.0
    assert (rec - split.float().double()).abs().max().item() <= 2e-7 * split.abs().max().item()
    A16, B16, X = (e16, hi, lo) if which == 2 else (hi, e16, lo)
    Ad, Bd = (exact.double(), split.float().double()) if which == 2 else (split.float().double(), exact.double())
    ref = _gemm_ref(layout, Ad, Bd)
    C = torch.full((M, (N + 3) & ~3), 3.0, device=dev)
    ws = torch.empty(max(K.gemm_f16_workspace_bytes(layout, M, N, Kd) // 4, 1), device=dev)
    K.gemm_f16_split(layout, M, N, Kd, A16, B16, X, which, C, alpha=0.25, workspace=ws)
    torch.cuda.synchronize()
    err = (C[:, :N].cpu().double() - ref).abs().max().item()
    assert err <= 1e-5 * ref.abs().max().item() + 1e-6 * math.sqrt(Kd), (layout, which, err, ref.abs().max().item())


CONTINUOUS = ["gaussian", "softplus gaussian", "log-normal", "gamma", "bernoulli", "lomax",
              "exponentially_modified_gaussian"]


@pytest.mark.parametrize("kind", CONTINUOUS)
@pytest.mark.parametrize("M,G,tile", [(37, 203, 1), (64, 2052, 2)])
def test_continuous_likelihood_fwd_bwd_and_moments(kind, M, G, tile):
    """csrc/continuous.cu against the oracle (fp64): log p per row, gradient w.r.t. the head
    pre-activations (activation, clip and density in one), moments."""
    from scvae_b200 import kernels as K
    rng = numpy.random.RandomState(5)
    heads = O.LIKELIHOODS[kind]
    P = len(heads)
    B = M // tile if M % tile == 0 else M
    tile = M // B
    Gn = (G + 3) & ~3
    a = (rng.randn(M, P, G) * 1.5).astype(numpy.float32)
    if kind == "gaussian":
        a[:, 1, ::17] = 4.0            # log_sigma beyond its clip
    if kind == "lomax":
        a[:, 0, ::19] = 11.0           # log_concentration beyond its clip
    if kind == "bernoulli":
        t = (rng.rand(B, G) < 0.3).astype(numpy.float32)
    elif kind in ("gaussian", "softplus gaussian", "exponentially_modified_gaussian"):
        t = (rng.randn(B, G) * 3).astype(numpy.float32)
    else:
        t = (rng.gamma(2.0, 1.5, (B, G)) + 0.05).astype(numpy.float32)
    go = (-(0.5 + rng.rand(M)) / M).astype(numpy.float32)
    dev = _dev()
    A = torch.zeros(M, P * Gn, device=dev)
    for h in range(P):
        A[:, h * Gn:h * Gn + G] = torch.tensor(a[:, h])
    T = torch.zeros(B, Gn, device=dev)
    T[:, :G] = torch.tensor(t)
    logp = torch.zeros(M, device=dev)
    dA = torch.zeros(M, P * Gn, device=dev)
    K.continuous_likelihood(K.LIKELIHOOD_KINDS[kind], T, A, Gn, M, G, logp=logp,
                            go=torch.tensor(go).to(dev), da=dA)
    logp_f = torch.zeros(M, device=dev)
    K.continuous_likelihood(K.LIKELIHOOD_KINDS[kind], T, A, Gn, M, G, logp=logp_f)
    torch.cuda.synchronize()
    pre = [torch.tensor(a[:, h], dtype=torch.float64).requires_grad_(True) for h in range(P)]
    theta = {head: O._clip_head(pre[h], head) for h, head in enumerate(heads)}
    x64 = torch.tensor(t, dtype=torch.float64).repeat(tile, 1)
    lp = O.continuous_log_prob(kind, x64, theta).sum(dim=1)
    (lp * torch.tensor(go, dtype=torch.float64)).sum().backward()
    ref = lp.detach().numpy()
    assert numpy.isfinite(ref).all()
    tol = 2e-5 if kind != "exponentially_modified_gaussian" else 2e-4     # erfc in fp32
    assert numpy.abs(logp.cpu().numpy() - ref).max() <= tol * numpy.abs(ref).max() + 1e-3
    assert torch.allclose(logp, logp_f, rtol=2e-6, atol=1e-4)      # (two instantiations: fma contraction differs)
    for h in range(P):
        g = pre[h].grad
        err = (dA[:, h * Gn:h * Gn + G].cpu().double() - g).abs().max().item()
        assert err <= 5e-4 * g.abs().max().item() + 1e-9, (kind, heads[h], err, g.abs().max().item())
    # moments (RS = tile samples per cell)
    outs = [torch.zeros(B, Gn, device=dev) for _ in range(3)]
    K.continuous_moments(K.LIKELIHOOD_KINDS[kind], A, Gn, B, G, tile, 1, None, *outs)
    torch.cuda.synchronize()
    with torch.no_grad():
        m, v = O.continuous_moments(kind, theta)
    m = m.reshape(tile, B, G)
    v = v.reshape(tile, B, G)
    mean = m.mean(dim=0)
    vom = ((m - mean) ** 2).mean(dim=0)
    want = [mean, torch.sqrt(vom + v.mean(dim=0)), torch.sqrt(vom)]
    for got, w in zip(outs, want):
        got = got[:, :G].cpu().double()
        ok = torch.isfinite(w)
        assert torch.equal(torch.isfinite(got), ok)            # nan / inf where a Lomax moment does not exist
        scale = w[ok].abs().max().item() if ok.any() else 1.0
        # (fp32 with fast exponentials; the Lomax variance spans ten orders of magnitude)
        assert (got[ok] - w[ok]).abs().max().item() <= 1e-3 * scale + 1e-6


@pytest.mark.parametrize("feeder", ["batch", "device", "host", "hybrid"])
@pytest.mark.parametrize("cap,direct", [(200.0, True), (3000.0, False), (60000.0, False)])
def test_csr_densify_packed_matches_csr_densify(cap, direct, feeder):
    """The streamed wire format (packed row slabs, scvae_csr_densify_packed) against the CSR form:
    identical 16-bit minibatch and per-cell constants, counts with and without escapes (>= 255), a
    ragged last slab, slabs assembled by the feeder thread (scvae_pack_row_slab)."""
    import scipy.sparse
    from scvae_b200 import kernels as K
    from scvae_b200.hotloop import PackedStream
    rng = numpy.random.RandomState(1)
    N, G, B = 300, 1000, 128
    dense = numpy.minimum(_counts(rng, N, G, 0.85) * (1 + 50 * (rng.rand(N, G) < 0.01)), cap).astype(numpy.float32)
    dense[7] = 0
    csr = scipy.sparse.csr_matrix(dense)
    dev = _dev()
    stream = PackedStream(csr, dev, B, feeder=feeder, host_fraction=0.4)
    assert 2.0 < stream.bytes_per_nonzero < 3.2
    order = rng.permutation(N)
    stream.pack_epoch(order)
    indptr = torch.tensor(csr.indptr.astype(numpy.int64)).to(dev)
    indices = torch.tensor(csr.indices.astype(numpy.int32)).to(dev)
    values = torch.tensor(csr.data.astype(numpy.float32)).to(dev)
    ld16 = (G + 8) & ~7
    for k, slab in enumerate(stream.slabs):
        rows = slab["rows"]
        slot = stream.fetch(k % 2, k)
        torch.cuda.current_stream().wait_event(slot["ready"])
        x16 = torch.full((rows, ld16), 3.0, dtype=torch.float16, device=dev)
        t16 = None if direct else torch.full((rows, (G + 7) & ~7), 7, dtype=torch.int16, device=dev)
        rc = torch.zeros(rows, device=dev)
        split = slot.get("split", 0)
        if split:       # hybrid feeder: the pulled rows and the host-gathered rows arrive as two slabs
            assert 0 < split < rows
            K.csr_densify_packed(slot["buf"], split, G, row_const=rc[:split],
                                 t16=None if t16 is None else t16[:split], x16=x16[:split])
            K.csr_densify_packed(slot["buf_host"], rows - split, G, row_const=rc[split:],
                                 t16=None if t16 is None else t16[split:], x16=x16[split:])
        else:
            K.csr_densify_packed(slot["buf"], rows, G, row_const=rc, t16=t16, x16=x16)
        slot["free"].record()
        idx = torch.tensor(order[k * B:k * B + rows].astype(numpy.int64)).to(dev)
        x16_ref = torch.full((rows, ld16), 5.0, dtype=torch.float16, device=dev)
        t16_ref = None if direct else torch.full((rows, (G + 7) & ~7), 9, dtype=torch.int16, device=dev)
        rc_ref = torch.zeros(rows, device=dev)
        K.csr_densify(indptr, indices, values, idx, G, None, rc_ref, t16=t16_ref, x16=x16_ref)
        torch.cuda.synchronize()
        assert torch.equal(x16, x16_ref)
        if not direct:
            assert torch.equal(t16, t16_ref)
        assert torch.allclose(rc, rc_ref, rtol=2e-6, atol=1e-5)
