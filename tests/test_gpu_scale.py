"""Parity at the BASELINE.json shapes, through the code the bench times.

Every other GPU test gives a CTA of the fused heads kernel exactly one 64-gene tile; the
benchmarked configurations make it walk dozens (C2: 313 gene tiles, 35 per CTA; C3: 438 tiles):
weight-ring wrap-around, mbarrier phase flips, `A_EMPTY` / `A_STORED` waits and the accumulation
of the decoder gradient over many tiles in TMEM.  These tests run that regime against the fp64
oracle on identical operands:

  * the fused heads kernel alone at (4096 x 20 000, NB) and (512 x 28 000, ZINB), fp16 and
    uint16 targets, forward + backward and forward-only;
  * one full `TrainLoop.step` (CUDA graph, the object `bench.py` times) from a `ResidentCSR` and
    from a `StreamedCSR` slot at the C2 shape (B = 4096, G = 20 000, L = 50, H = [100], NB) and at
    the C3 shape (B = 512, G = 28 000, L = 100, ZINB): ELBO / ENRE / KL <= 1e-3 relative
    (north-star tolerance), per-cell latent means and log-likelihoods, raw gradients and the
    variables after clip + Adam (VAE:1026-1029, :2560-2770);
  * one GMVAE step with K = 20 clusters at G = 20 000 (C4 shape, cluster-chunked fused decoder).
"""
import math

import numpy
import pytest
import torch

from oracle import scvae_oracle as O

pytestmark = pytest.mark.gpu

ELBO_TOL = 1e-3          # BASELINE.json north_star: "ELBO within 1e-3 relative of the reference"


def _dev():
    return torch.device("cuda:0")


def _sparse_counts(rng, rows, G, density, cap):
    """10x-like sparse counts: ~density non-zero, values >= 1 with a geometric tail."""
    mask = rng.rand(rows, G) < (rng.rand(G) * 2.0 * density)
    vals = numpy.floor(1.0 - numpy.log(rng.rand(rows, G)) * 1.2)
    return numpy.minimum(mask * vals, cap).astype(numpy.float32)


def _oracle_logp(kind, t64, a64):
    heads = O.LIKELIHOODS[kind]
    theta = {h: O._clip_head(a, h) for h, a in zip(heads, a64)}
    return O.likelihood_log_prob(kind, t64, theta)


@pytest.mark.parametrize("kind,M,G,H,t_half", [
    ("negative binomial", 4096, 20000, 100, True),
    ("negative binomial", 4096, 20000, 100, False),
    ("zero-inflated negative binomial", 512, 28000, 100, False),
    ("zero-inflated negative binomial", 512, 28000, 100, True),
    ("poisson", 1024, 20000, 100, True),
], ids=["nb-c2-f16t", "nb-c2-u16t", "zinb-c3-u16t", "zinb-c3-f16t", "poisson-20k"])
def test_heads_fused_many_tiles_per_cta(kind, M, G, H, t_half):
    """heads_fused_kernel in its pipelined steady state (tiles_per_cta >> 1) vs fp64 on the
    same fp16-rounded operands: log p per cell, the fp16 head gradient, the decoder gradient."""
    from scvae_b200 import kernels as K
    rng = numpy.random.RandomState(17)
    P = len(O.LIKELIHOODS[kind])
    Gh = (G + 63) & ~63
    dev = _dev()
    d = numpy.abs(rng.randn(M, H)).astype(numpy.float32)
    d_aug = numpy.concatenate([d, numpy.ones((M, 1), numpy.float32)], axis=1)
    w = (rng.randn(P, G, H + 1) * (0.6 / math.sqrt(H))).astype(numpy.float32)
    t = _sparse_counts(rng, M, G, 0.07, 2000.0 if t_half else 60000.0)
    if not t_half:
        t[::7, ::501] = 40000.0            # beyond fp16: the uint16 encoding is the exact one
    go = (-(0.5 + rng.rand(M)) / M).astype(numpy.float32)
    scale = 2.0 ** round(math.log2(M / 16.0))
    d16 = torch.zeros(M, 128, dtype=torch.float16, device=dev)
    d16[:, :H + 1] = torch.tensor(d_aug).half()
    w16 = torch.zeros(P * Gh, 128, dtype=torch.float16, device=dev)
    for h in range(P):
        w16[h * Gh:h * Gh + G, :H + 1] = torch.tensor(w[h]).half()
    if t_half:
        t16 = torch.zeros(M, Gh, dtype=torch.float16, device=dev)
        t16[:, :G] = torch.tensor(t).half()
    else:
        t16 = torch.zeros(M, Gh, dtype=torch.int16, device=dev)
        K.f32_to_u16(torch.tensor(t).to(dev), G, t16)
    rc = torch.lgamma(1.0 + torch.tensor(t, dtype=torch.float64)).sum(dim=1).float().to(dev)
    da16 = torch.zeros(M, P * Gh, dtype=torch.float16, device=dev)
    dd = torch.full((M, 104), 5.0, device=dev)
    logp = torch.zeros(M, device=dev)
    ws = torch.zeros(K.heads_fused_workspace_floats(M, G), device=dev)
    # the regime under test: several gene tiles per CTA (the plan of heads_fused.cu:fused_plan)
    row_tiles, n_tiles = (M + 127) // 128, (G + 63) // 64
    target = max(1, min(n_tiles, (2 * 148 + row_tiles // 2) // row_tiles))
    assert (n_tiles + target - 1) // target >= 4
    K.heads_fused_bwd(K.LIKELIHOOD_KINDS[kind], d16, w16, Gh, t16, M, G, da16, dd, H, logp, ws,
                      row_const=rc, go=torch.tensor(go).to(dev), scale=scale)
    torch.cuda.synchronize()
    d64 = d16[:, :H + 1].cpu().double()
    w64 = [w16[h * Gh:h * Gh + G, :H + 1].cpu().double() for h in range(P)]
    a64 = [(d64 @ w64[h].t()).requires_grad_(True) for h in range(P)]
    lp = _oracle_logp(kind, torch.tensor(t, dtype=torch.float64), a64).sum(dim=1)
    (lp * torch.tensor(go, dtype=torch.float64)).sum().backward()
    ref = lp.detach().numpy()
    assert numpy.isfinite(ref).all()
    err = numpy.abs(logp.cpu().numpy() - ref).max()
    assert err <= 3e-5 * numpy.abs(ref).max() + 1e-4, ("log p", err, numpy.abs(ref).max())
    got_da = da16.cpu()
    for h in range(P):
        g = a64[h].grad
        e = (got_da[:, h * Gh:h * Gh + G].double() / scale - g).abs().max().item()
        assert e <= 2e-3 * g.abs().max().item(), (kind, "da", h, e, g.abs().max().item())
        if Gh > G:
            assert got_da[:, h * Gh + G:(h + 1) * Gh].abs().max().item() == 0
    dd_ref = sum(a64[h].grad @ w64[h] for h in range(P))[:, :H]
    e = (dd[:, :H].cpu().double() - dd_ref).abs().max().item()
    assert e <= 3e-3 * dd_ref.abs().max().item(), (kind, "dd", e, dd_ref.abs().max().item())
    logp_f = torch.zeros(M, device=dev)
    K.heads_fused_fwd(K.LIKELIHOOD_KINDS[kind], d16, w16, Gh, t16, M, G, logp_f, ws, row_const=rc)
    torch.cuda.synchronize()
    err = numpy.abs(logp_f.cpu().numpy() - ref).max()
    assert err <= 3e-5 * numpy.abs(ref).max() + 1e-4, ("forward-only log p", err)


def _rel(a, b):
    a = numpy.asarray(a, dtype=numpy.float64)
    b = numpy.asarray(b, dtype=numpy.float64)
    return numpy.abs(a - b).max() / (numpy.abs(b).max() + 1e-30)


STEP_CASES = [
    # name, B, G, L, likelihood, count cap
    ("c2-nb-b4096", 4096, 20000, 50, "negative binomial", 500.0),
    ("c3-zinb-b512", 512, 28000, 100, "zero-inflated negative binomial", 500.0),
    ("c2-nb-b1024-large-counts", 1024, 20000, 50, "negative binomial", 30000.0),
]


@pytest.mark.parametrize("source", ["resident", "streamed", "packed"])
@pytest.mark.parametrize("case", STEP_CASES, ids=[c[0] for c in STEP_CASES])
def test_train_loop_step_at_baseline_shape(case, source):
    """`TrainLoop.step` (CUDA-graph replay of densify -> noise -> forward -> backward -> clip +
    Adam, i.e. what bench.py times as `value` / `e2e`) against `O.train_step` on the same rows,
    weights and noise."""
    import scipy.sparse
    from scvae_b200.engine import VAEEngine
    from scvae_b200.hotloop import ResidentCSR, StreamedCSR, TrainLoop
    name, B, G, L, lik, cap = case
    dev = _dev()
    rng = numpy.random.RandomState(23)
    N = 2 * B
    x_all = _sparse_counts(rng, N, G, 0.07, cap)
    csr = scipy.sparse.csr_matrix(x_all)
    eng = VAEEngine(G, L, [100], lik, device=dev, seed=4)
    # a state a few hundred steps into training: non-trivial biases, BN betas and Adam slots
    gen = torch.Generator().manual_seed(9)
    params = eng.export_parameters()
    for k in params:
        if k.endswith("biases") or k.endswith("beta"):
            params[k] = torch.randn(params[k].shape, generator=gen) * 0.1
    eng.import_parameters(params)
    loop = TrainLoop(eng, B, seed=31, use_graph=True)
    lr, w = 1e-3, 0.8
    if source == "resident":
        data = ResidentCSR(csr, dev)
        rows = torch.from_numpy(rng.permutation(N)[:B].astype(numpy.int64)).to(dev)
        loop.rows.copy_(rows)
        src = data
        rows_host = rows.cpu().numpy()
    elif source == "packed":
        from scvae_b200.hotloop import PackedStream
        stream = PackedStream(csr, dev, B)
        order = rng.permutation(N)
        stream.pack_epoch(order)
        stream.fetch(0, 0)                              # (slabs are taken in epoch order)
        src = stream.fetch(1, 1)                        # the second slab of the epoch
        torch.cuda.current_stream().wait_event(src["ready"])
        rows_host = order[B:2 * B]
    else:
        stream = StreamedCSR(csr, dev, B)
        src = stream.fetch(0, B // 2, B // 2 + B)       # a slab that does not start at row 0
        torch.cuda.current_stream().wait_event(src["ready"])
        rows_host = numpy.arange(B // 2, B // 2 + B)
    before = {k: v.double() for k, v in eng.export_parameters().items()}
    bound = loop.step(src, lr, w)
    torch.cuda.synchronize()
    plan = loop.plan
    assert plan.fused_done, "the benchmarked (fused 16-bit) path was not taken"
    bound = bound.cpu().numpy()
    eps = plan.eps.cpu().double().reshape(1, B, L)
    x = torch.tensor(x_all[rows_host], dtype=torch.float64)

    cfg = O.VAEConfig(G, L, [100], lik)
    state = O.AdamState(before)
    ref = {k: v.clone() for k, v in before.items()}
    out, grads = O.train_step(cfg, ref, state, x, x, eps, lr, warm_up_weight=w)
    for i, key in enumerate(["lower_bound", "lower_bound_weighted", "reconstruction_error",
                             "kl_divergence"]):
        r = out[key].item()
        assert abs(bound[i] - r) <= ELBO_TOL * abs(r), (key, bound[i], r)
    assert _rel(plan.PH[:, :L].cpu(), out["q_z_mean"]) <= 1e-3
    assert _rel(plan.logp.cpu(), out["log_p_x_given_z"].reshape(-1)) <= 1e-4
    got = eng.export_gradients()
    gmax = max(g.abs().max().item() for g in grads.values())
    for k, g in grads.items():
        err = (got[k].double() - g).abs().max().item()
        print("gradient error", name, source, k, "max|err| %.3e  max|g| %.3e  ratio %.2e"
              % (err, g.abs().max().item(), err / (g.abs().max().item() + 1e-30)))      # (shown with -s)
        assert err <= 1e-2 * g.abs().max().item() + 1e-4 * gmax, (k, err, g.abs().max().item())
    new = eng.export_parameters()
    noise = 3e-2 * gmax
    for k, v in ref.items():
        diff = (new[k].double() - v).abs()
        if k in grads:
            diff = diff * (grads[k].abs() > noise)
        rtol = 1e-4 if "moving" in k else 1e-5
        assert diff.max().item() <= rtol * max(v.abs().max().item(), 1.0), (k, diff.max().item())
    # a second replay of the same graph on other rows keeps working (ring / phase state is
    # per launch) and stays finite
    if source == "resident":
        loop.rows.copy_(torch.from_numpy(rng.permutation(N)[:B].astype(numpy.int64)).to(dev))
    b2 = loop.step(src, lr, w).cpu().numpy()
    assert numpy.isfinite(b2).all()


def test_gmvae_step_k20_at_20k_genes():
    """C4 shape: GMVAE, K = 20 clusters, G = 20 000, L = 50, NB: the cluster passes go through
    the fused heads kernel in chunks (targets tile over the cluster rows)."""
    from scvae_b200.gmvae_engine import GMVAEEngine
    G, L, Kc, B = 20000, 50, 20, 128
    lik = "negative binomial"
    cfg = O.GMVAEConfig(G, L, Kc, [100], lik, 1, 1, True)
    params = O.gmvae_init_params(cfg, seed=2, dtype=torch.float64)
    gen = torch.Generator().manual_seed(5)
    for k in params:
        if k.endswith("biases") or k.endswith("beta") or k.endswith("LOGITS"):
            params[k] = torch.randn(params[k].shape, generator=gen, dtype=torch.float64) * 0.1
    rng = numpy.random.RandomState(3)
    x = _sparse_counts(rng, B, G, 0.07, 500.0)
    eps = torch.randn(Kc, 1, B, L, generator=gen, dtype=torch.float64)
    # head_buffer_bytes small enough to force several decoder chunks
    eng = GMVAEEngine(G, L, Kc, [100], lik, True, 1.0, "uniform", None, 0.0, device="cuda:0",
                      tensor_cores=True, head_buffer_bytes=256 << 20)
    eng.import_parameters(params)
    plan = eng._plan(B, 1)
    eng.set_batch_dense(plan, torch.tensor(x).cuda())
    plan.eps.copy_(eps.reshape(-1, L).float())
    state = O.AdamState(params)
    ref = {k: v.clone() for k, v in params.items()}
    x64 = torch.tensor(x, dtype=torch.float64)
    out, grads = O.train_step(cfg, ref, state, x64, x64, eps, 1e-3, warm_up_weight=1.0)
    bound = eng.train_step(plan, 1, 1, 1e-3, warm_up_weight=1.0).cpu().numpy()
    torch.cuda.synchronize()
    assert plan.fused_done and plan.chunk < Kc
    names = ["lower_bound", "lower_bound_weighted", "reconstruction_error", "kl_divergence_z",
             "kl_divergence_y"]
    for i, n in enumerate(names):
        assert abs(bound[i] - out[n].item()) <= ELBO_TOL * abs(out[n].item()) + 1e-5, \
            (n, bound[i], out[n].item())
    lp_ref = out["log_p_x_given_z"].reshape(-1)
    assert ((plan.logp.cpu().double() - lp_ref).abs().max() / lp_ref.abs().max()).item() <= 1e-4
    got = eng.export_gradients()
    gmax = max(g.abs().max().item() for g in grads.values())
    for k, g in grads.items():
        err = (got[k].double() - g).abs().max().item()
        assert err <= 1e-2 * g.abs().max().item() + 1e-4 * gmax, (k, err, g.abs().max().item())


def test_resident_csr_beyond_two_to_the_31_non_zeros():
    """C3 at full size holds 2.5 G non-zeros: the CSR offsets are 64-bit end to end.  A synthetic
    matrix with > 2^31 stored entries (1400 per row, built on the device: 17.6 GB) is densified and
    its row constants computed for rows on both sides of the 2^31 boundary."""
    from scvae_b200 import kernels as K
    dev = _dev()
    free, _ = torch.cuda.mem_get_info(dev)
    if free < 40 << 30:
        pytest.skip("needs 40 GB of free device memory")
    G, per_row = 20000, 1400
    n_rows = (1 << 31) // per_row + 4096              # ~1.54 M rows, nnz = 2.153 G > 2^31
    nnz = n_rows * per_row
    assert nnz > (1 << 31)
    indptr = torch.arange(n_rows + 1, dtype=torch.int64, device=dev) * per_row
    indices = torch.empty(nnz, dtype=torch.int32, device=dev)
    values = torch.empty(nnz, dtype=torch.float32, device=dev)
    chunk = 1 << 27
    for s in range(0, nnz, chunk):
        i = torch.arange(s, min(s + chunk, nnz), dtype=torch.int64, device=dev)
        pos, row = i % per_row, i // per_row
        indices[s:s + i.numel()] = (pos * 14 + row % 14).to(torch.int32)       # sorted within a row, < 19 600
        values[s:s + i.numel()] = (1 + (i % 3)).to(torch.float32)
        del i, pos, row
    rows = torch.tensor([0, 5, (1 << 31) // per_row - 1, (1 << 31) // per_row, (1 << 31) // per_row + 1,
                         n_rows - 1], dtype=torch.int64, device=dev)
    B = rows.numel()
    ld16 = (G + 8) & ~7
    x16 = torch.zeros(B, ld16, dtype=torch.float16, device=dev)
    x32 = torch.zeros(B, (G + 4) & ~3, dtype=torch.float32, device=dev)
    rc = torch.zeros(B, device=dev)
    K.csr_densify(indptr, indices, values, rows, G, None, rc, x16=x16)          # 16-bit form
    K.csr_densify(indptr, indices, values, rows, G, x32, None)                  # general form
    rc_all = torch.zeros(n_rows, device=dev)
    K.csr_row_constants(indptr, values, rc_all)
    torch.cuda.synchronize()
    for b, r in enumerate(rows.tolist()):
        want = torch.zeros(G, dtype=torch.float64)
        i = torch.arange(r * per_row, (r + 1) * per_row, dtype=torch.int64)
        want[(i % per_row) * 14 + r % 14] = (1 + (i % 3)).double()
        assert torch.equal(x16[b, :G].cpu().double(), want), r
        assert torch.equal(x32[b, :G].cpu().double(), want), r
        ref = torch.lgamma(1.0 + want).sum().item()
        assert abs(rc[b].item() - ref) <= 1e-4 * ref and abs(rc_all[r].item() - ref) <= 1e-4 * ref, r
