"""Host logic of the VAE engine on the CPU: buffer layout, launch sequence, gradient routing and
optimiser wiring of ``scvae_b200.engine.VAEEngine`` with every kernel wrapper replaced by its CPU
stand-in (``tests/kernel_standins.py``), held to the golden vectors recorded from the reference's
own graph code.  The very same test bodies run against the real kernels on the GPU
(``tests/test_zz_gpu_reference_graph.py``); here they prove that what surrounds the kernels --
and only that -- reproduces the reference.  The product itself has no CPU path: the stand-ins
are injected by monkeypatching inside this test module only.
"""
import pytest
import torch

import kernel_standins
import test_zz_gpu_reference_graph as Z


@pytest.fixture
def engine_on_cpu(monkeypatch):
    import scvae_b200.engine as E
    monkeypatch.setattr(E, "K", kernel_standins)
    original = E.VAEEngine.__init__

    def init(self, *args, **kwargs):
        kwargs["device"] = "cpu"
        original(self, *args, **kwargs)

    monkeypatch.setattr(E.VAEEngine, "__init__", init)
    import scvae_b200.gmvae_engine as GE
    monkeypatch.setattr(GE, "K", kernel_standins)
    gm_original = GE.GMVAEEngine.__init__

    def gm_init(self, *args, **kwargs):
        kwargs["device"] = "cpu"
        gm_original(self, *args, **kwargs)

    monkeypatch.setattr(GE.GMVAEEngine, "__init__", gm_init)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    del kernel_standins.launches[:]
    return kernel_standins.launches


@pytest.mark.parametrize("name", Z.VAE_TRAIN + ["vae_nb_sampled_kl_train",
                                                "vae_nb_unit_variance_train"])
def test_vae_engine_training_step_host_logic(engine_on_cpu, name):
    Z.test_vae_training_step_matches_reference_graph(name)
    launches = engine_on_cpu
    # one GPU: the optimiser launch advances the step counter itself (last CTA done)
    assert launches[-1] == "adam_clip_step" and launches.count("adam_clip_step") == 1
    assert "step_advance" not in launches
    sampled = name in ("vae_nb_sampled_kl_train", "vae_nb_unit_variance_train")
    assert ("gaussian_sampled_kl_bwd" in launches) == sampled
    assert ("gaussian_latent_bwd" in launches) == (not sampled)
    assert ("vae_bound_rows" in launches) == sampled


@pytest.mark.parametrize("name", Z.DROPOUT_CASES[1:])
def test_vae_engine_dropout_host_logic_more_sites(engine_on_cpu, name):
    Z.test_vae_dropout_training_step_matches_reference_graph(name)
    assert "fill_normal" not in engine_on_cpu


def test_vae_engine_dropout_host_logic(engine_on_cpu):
    """Dropout with the masks the reference-graph run recorded: six sites, one product per
    posterior parameter and per likelihood head, masked gradient routing."""
    Z.test_vae_dropout_training_step_matches_reference_graph(Z.DROPOUT_CASES[0])
    launches = engine_on_cpu
    assert launches.count("dropout_fwd") == 6       # ENCODER/1, MU, LOG_SIGMA, DECODER/1, P, LOG_R
    # two heads into d(decoder output), DECODER/1 in place, two posterior heads into d(encoder
    # output); ENCODER/1 masks the data itself (no input gradient)
    assert launches.count("dropout_bwd") == 5
    assert "fill_normal" not in launches            # masks injected, nothing drawn


@pytest.mark.parametrize("name", Z.VAE_EVAL)
def test_vae_engine_evaluation_host_logic(engine_on_cpu, name):
    Z.test_vae_evaluation_matches_reference_graph(name)
    assert "adam_clip_step" not in engine_on_cpu


@pytest.mark.parametrize("name", ["vae_nb_sampled_kl_eval_deterministic",
                                  "vae_nb_sampled_kl_eval_iw"])
def test_vae_engine_sampled_kl_evaluation_host_logic(engine_on_cpu, name):
    Z.test_vae_sampled_kl_evaluation_matches_reference_graph(name)
    assert "gaussian_sampled_kl" in engine_on_cpu and "vae_bound" not in engine_on_cpu


@pytest.mark.parametrize("name", Z.GMVAE_TRAIN)
def test_gmvae_engine_training_step_host_logic(engine_on_cpu, name):
    Z.test_gmvae_training_step_matches_reference_graph(name)
    assert engine_on_cpu.count("gmvae_bound") == 1


@pytest.mark.parametrize("name", Z.GMVAE_EVAL)
def test_gmvae_engine_evaluation_host_logic(engine_on_cpu, name):
    Z.test_gmvae_evaluation_matches_reference_graph(name)
    assert "adam_clip_step" not in engine_on_cpu


def test_gmvae_engine_constrained_poisson_host_logic(engine_on_cpu):
    Z.test_gmvae_constrained_poisson_matches_reference_graph()
    assert "constrained_poisson_mixture_moments" in engine_on_cpu
    assert "likelihood_fwd" not in engine_on_cpu and "likelihood_bwd" not in engine_on_cpu


def test_product_kernels_module_is_untouched_outside_the_fixture():
    import scvae_b200.engine as E
    import scvae_b200.gmvae_engine as GE
    import scvae_b200.kernels as K
    assert E.K is K and GE.K is K


def test_vae_engine_draws_its_own_dropout_masks(engine_on_cpu):
    """Without injected masks every site draws its own noise each step (one generator call per
    site, keyed by site and optimiser step); evaluation passes do not drop anything."""
    import numpy
    from scvae_b200.engine import VAEEngine
    from oracle import scvae_oracle as O
    G, L, B = 40, 3, 64
    eng = VAEEngine(G, L, [16, 8], "negative binomial", tensor_cores=False,
                    dropout_keep_probabilities=[0.8, 0.9, 0.7])
    x, _ = O.synthetic_counts(B, G, n_types=3, seed=4, target_zero_fraction=0.8)
    plan = eng._plan(B, 1)
    eng.set_batch_dense(plan, torch.tensor(numpy.minimum(x, 30.0), dtype=torch.float32))
    plan.eps.normal_(generator=torch.Generator().manual_seed(0))
    bound = eng.train_step(plan, 1, 1, 1e-3)
    assert torch.isfinite(bound).all()
    launches = list(engine_on_cpu)
    sites = ["ENCODER/1", "ENCODER/2", "POSTERIOR/MU", "POSTERIOR/LOG_SIGMA", "DECODER/2",
             "DECODER/1", "X_TILDE/P", "X_TILDE/LOG_R"]
    assert list(plan.drop) == sites
    assert launches.count("fill_normal") == len(sites) == launches.count("dropout_fwd")
    assert len({st.seed for st in plan.drop.values()}) == len(sites)
    for site, keep in (("ENCODER/1", 0.9), ("ENCODER/2", 0.8), ("DECODER/2", 0.7)):
        st = plan.drop[site]
        kept = (st.noise < st.threshold).float().mean().item()
        assert abs(kept - keep) < 4 * (keep * (1 - keep) / st.noise.numel()) ** 0.5, (site, kept)
    # the dropped copy: kept entries scaled by 1 / keep, the ones column intact
    st = plan.drop["ENCODER/1"]
    mask = (st.noise < st.threshold).float()
    assert torch.allclose(st.copy[:, :G], plan.X[:, :G] * mask / 0.9)
    assert torch.equal(st.copy[:, G], torch.ones(B))
    del engine_on_cpu[:]
    eng.forward(plan, False, 1, 1, 1.0)
    assert "dropout_fwd" not in engine_on_cpu and "fill_normal" not in engine_on_cpu


@pytest.mark.parametrize("seed,G,L,hidden,lik,R,S,B,extras,k_max", [
    (0, 37, 5, [7, 5], "zero-inflated negative binomial", 1, 1, 9, dict(number_of_batches=3), 0),
    (1, 41, 2, [6], "negative binomial", 2, 1, 5, dict(count_sum_feature=True), 2),
    (2, 30, 3, [9, 4, 6], "poisson", 1, 3, 7, dict(number_of_batches=2, count_sum_feature=True), 0),
    (3, 26, 4, [5], "zero-inflated poisson", 1, 1, 11, dict(generative_architecture="LFM"), 0),
])
def test_vae_engine_dropout_odd_shapes_against_the_oracle(engine_on_cpu, seed, G, L, hidden, lik,
                                                          R, S, B, extras, k_max):
    """Dropout wiring at sizes that are not multiples of four (padded leading dimensions, the
    ones column between z and the decoder extras, head blocks behind the P_K block), against
    the oracle with the same injected masks: ELBO, every gradient, the clip + Adam step."""
    import numpy
    from oracle import scvae_oracle as O
    from scvae_b200.engine import VAEEngine
    keep = [0.8, 0.9, 0.7]
    cfg = O.VAEConfig(G, L, hidden, lik, "gaussian", R, S, True, True, kl_weight=0.9,
                      number_of_reconstruction_classes=k_max, dropout_keep_probabilities=keep,
                      **extras)
    params = O.vae_init_params(cfg, seed=seed, dtype=torch.float64)
    gen = torch.Generator().manual_seed(100 + seed)
    for k in params:
        if k.endswith("biases") or k.endswith("beta"):
            params[k] = torch.randn(params[k].shape, generator=gen, dtype=torch.float64) * 0.1
    x = torch.tensor(numpy.minimum(O.synthetic_counts(B, G, n_types=3, seed=seed)[0], 20.0),
                     dtype=torch.float64)
    eps = torch.randn(R * S, B, L, generator=gen, dtype=torch.float64)
    feats = {}
    if cfg.number_of_batches:
        feats["batch_indices"] = torch.randint(0, cfg.number_of_batches, (B, 1), generator=gen)
    if cfg.count_sum_feature:
        feats["count_sum_feature"] = torch.rand(B, 1, generator=gen, dtype=torch.float64)
    dropout = {"generator": torch.Generator().manual_seed(7 + seed)}      # masks drawn once ...
    state = O.AdamState(params)
    reference = {k: v.clone() for k, v in params.items()}
    out, grads = O.train_step(cfg, reference, state, x, x, eps, 1e-3, warm_up_weight=0.6,
                              dropout=dropout, **feats)
    eng = VAEEngine(G, L, hidden, lik, "gaussian", True, kl_weight=0.9, tensor_cores=False,
                    number_of_reconstruction_classes=k_max, dropout_keep_probabilities=keep,
                    **extras)
    eng.import_parameters(params)
    plan = eng._plan(B, R * S)
    eng.set_batch_dense(plan, x.float())
    eng.set_batch_features(plan, feats.get("batch_indices"),
                           feats["count_sum_feature"].float() if "count_sum_feature" in feats
                           else None)
    plan.eps.copy_(eps.reshape(R * S * B, L).float())
    eng.inject_dropout_masks(plan, dropout["masks"])                      # ... and shared
    bound = eng.train_step(plan, R, S, 1e-3, warm_up_weight=0.6)
    assert sorted(plan.drop) == sorted(dropout["masks"])
    assert abs(bound[0].item() - out["lower_bound"].item()) <= 5e-5 * abs(out["lower_bound"].item())
    got = eng.export_gradients()
    gmax = max(g.abs().max().item() for g in grads.values() if g is not None)
    for key, g in grads.items():
        g = g if g is not None else torch.zeros_like(params[key])
        error = (got[key].double() - g).abs().max().item()
        assert error <= 2e-4 * g.abs().max().item() + 1e-5 * gmax, (key, error)
    new = eng.export_parameters()
    for key, value in reference.items():
        if "moving" in key:
            continue
        difference = (new[key].double() - value).abs()
        if key in grads and grads[key] is not None:
            difference = difference * (grads[key].abs() > 1e-3 * gmax)
        assert difference.max().item() <= 1e-5 * max(value.abs().max().item(), 1.0), key


@pytest.mark.parametrize("head_buffer_bytes", [4 << 30, 20000])
def test_gmvae_engine_constrained_poisson_chunked_against_the_oracle(engine_on_cpu,
                                                                     head_buffer_bytes):
    """Constrained Poisson through the cluster-chunked decoder (each chunk writes its slice of
    the row log-sum-exps) at sizes that are not multiples of four."""
    import numpy
    from oracle import scvae_oracle as O
    from scvae_b200.gmvae_engine import GMVAEEngine
    G, L, Kc, hidden, B, R, S = 37, 3, 4, [7, 5], 9, 1, 2
    cfg = O.GMVAEConfig(G, L, Kc, hidden, "constrained poisson", R, S, True, kl_weight=0.8)
    params = O.gmvae_init_params(cfg, seed=3, dtype=torch.float64)
    gen = torch.Generator().manual_seed(4)
    for k in params:
        if k.endswith("biases") or k.endswith("beta"):
            params[k] = torch.randn(params[k].shape, generator=gen, dtype=torch.float64) * 0.2
    x = torch.tensor(numpy.minimum(O.synthetic_counts(B, G, n_types=3, seed=5)[0], 40.0),
                     dtype=torch.float64)
    eps = torch.randn(Kc, R * S, B, L, generator=gen, dtype=torch.float64)
    count_sum = x.sum(dim=1, keepdim=True)
    state = O.AdamState(params)
    reference = {k: v.clone() for k, v in params.items()}
    out, grads = O.train_step(cfg, reference, state, x, x, eps, 1e-3, warm_up_weight=0.7,
                              count_sum=count_sum)
    eng = GMVAEEngine(G, L, Kc, hidden, "constrained poisson", True, 0.8, "uniform", None, 0.0,
                      tensor_cores=False, head_buffer_bytes=head_buffer_bytes)
    eng.import_parameters(params)
    plan = eng._plan(B, R * S)
    assert (plan.chunk < Kc) == (head_buffer_bytes < 1 << 20)
    eng.set_batch_dense(plan, x.float())
    eng.set_batch_count_sum_parameter(plan, count_sum.float())
    plan.eps.copy_(eps.reshape(-1, L).float())
    bound = eng.train_step(plan, R, S, 1e-3, warm_up_weight=0.7)
    for i, key in enumerate(["lower_bound", "lower_bound_weighted", "reconstruction_error"]):
        assert abs(bound[i].item() - out[key].item()) <= 5e-5 * abs(out[key].item()), key
    # every row's log-sum-exp landed in its slot, whatever the chunking
    a_rows = torch.logsumexp(plan.A[:plan.chunk * R * S * B, :G].double(), dim=1)
    first = plan.chunk * R * S * B
    last_chunk_rows = (Kc - (Kc - 1) // plan.chunk * plan.chunk) * R * S * B
    assert torch.allclose(plan.lse_all[-last_chunk_rows:].double(), a_rows[:last_chunk_rows],
                          rtol=1e-5) or plan.chunk == Kc
    if plan.chunk == Kc:
        assert torch.allclose(plan.lse_all.double(), a_rows[:first], rtol=1e-5)
    got = eng.export_gradients()
    gmax = max(g.abs().max().item() for g in grads.values())
    for key, g in grads.items():
        error = (got[key].double() - g).abs().max().item()
        assert error <= 3e-4 * g.abs().max().item() + 1e-5 * gmax, (key, error)


@pytest.mark.parametrize("head_buffer_bytes,keep,lik,k_max", [
    (4 << 30, [0.8, 0.9, 0.7, 0.6], "negative binomial", 0),
    (20000, [0.8, 0.9, 0.7, 0.6], "zero-inflated negative binomial", 0),    # chunked decoder
    (20000, [0.7, False, 0.8], "poisson", 2),              # no x / y dropout: the shared x W_x path
])
def test_gmvae_engine_dropout_against_the_oracle(engine_on_cpu, head_buffer_bytes, keep, lik, k_max):
    """GMVAE dropout (GMVAE:276-296, :3036-3040): one mask per site AND per cluster build, the
    first q(z|x,y) layer on K dropped copies of [x | e_k], the p(z|y) heads behind a dropped
    one-hot, decoder sites sliced by the cluster chunks -- odd shapes, masks shared with the
    oracle, with and without chunking."""
    import numpy
    from oracle import scvae_oracle as O
    from scvae_b200.gmvae_engine import GMVAEEngine
    G, L, Kc, hidden, B, R, S = 37, 3, 4, [7, 5], 9, 1, 2
    cfg = O.GMVAEConfig(G, L, Kc, hidden, lik, R, S, True, kl_weight=0.8,
                        number_of_reconstruction_classes=k_max, dropout_keep_probabilities=keep)
    params = O.gmvae_init_params(cfg, seed=3, dtype=torch.float64)
    gen = torch.Generator().manual_seed(4)
    for k in params:
        if k.endswith("biases") or k.endswith("beta"):
            params[k] = torch.randn(params[k].shape, generator=gen, dtype=torch.float64) * 0.2
    x = torch.tensor(numpy.minimum(O.synthetic_counts(B, G, n_types=3, seed=5)[0], 40.0),
                     dtype=torch.float64)
    eps = torch.randn(Kc, R * S, B, L, generator=gen, dtype=torch.float64)
    state = O.AdamState(params)
    reference = {k: v.clone() for k, v in params.items()}
    dropout = {"generator": torch.Generator().manual_seed(11)}
    out, grads = O.train_step(cfg, reference, state, x, x, eps, 1e-3, warm_up_weight=0.7,
                              dropout=dropout)
    eng = GMVAEEngine(G, L, Kc, hidden, lik, True, 0.8, "uniform", None, 0.0, tensor_cores=False,
                      head_buffer_bytes=head_buffer_bytes, number_of_reconstruction_classes=k_max,
                      dropout_keep_probabilities=keep)
    eng.import_parameters(params)
    plan = eng._plan(B, R * S)
    assert (plan.chunk < Kc) == (head_buffer_bytes < 1 << 20)
    eng.inject_dropout_masks(plan, dropout["masks"])
    eng.set_batch_dense(plan, x.float())
    plan.eps.copy_(eps.reshape(-1, L).float())
    bound = eng.train_step(plan, R, S, 1e-3, warm_up_weight=0.7)
    for i, key in enumerate(["lower_bound", "lower_bound_weighted", "reconstruction_error"]):
        assert abs(bound[i].item() - out[key].item()) <= 5e-5 * abs(out[key].item()), key
    # every mask the oracle drew was consumed: site names + one per cluster build
    assert sorted(k for k in dropout["masks"] if "#" not in k) == sorted(plan.drop)
    got = eng.export_gradients()
    gmax = max(g.abs().max().item() for g in grads.values())
    for key, g in grads.items():
        error = (got[key].double() - g).abs().max().item()
        assert error <= 3e-4 * g.abs().max().item() + 1e-5 * gmax, (key, error)
    # evaluation never drops
    eng.forward(plan, False, R, S, 1.0)
    now = {k: v.double() for k, v in eng.export_parameters().items()}
    ref_eval = O.gmvae_forward(cfg, now, x, x, eps, is_training=False)
    assert abs(plan.bound[0].item() - ref_eval["lower_bound"].item()) <= 5e-5 * abs(
        ref_eval["lower_bound"].item())


@pytest.mark.parametrize("head_buffer_bytes,lik,R,S,prior", [
    (4 << 30, "negative binomial", 1, 2, "uniform"),
    (20000, "poisson", 2, 1, "learn"),                    # chunked decoder, learnt prior
])
def test_gmvae_engine_full_covariance_mixture_against_the_oracle(engine_on_cpu, head_buffer_bytes, lik,
                                                                 R, S, prior):
    """`-q "full-covariance gaussian mixture"` (f4; DU:75-93, multivariate_normal.py:90-150):
    multivariate-Gaussian q(z|x,y) / p(z|y) heads with fill_triangular scale matrices through the
    engine -- bound terms, gradients of every variable (incl. the L (L + 1) / 2 scale heads of both
    distributions), post-Adam variables, evaluation -- against the oracle at odd shapes."""
    import numpy
    from oracle import scvae_oracle as O
    from scvae_b200.gmvae_engine import GMVAEEngine
    G, L, Kc, hidden, B = 37, 3, 4, [7, 5], 9
    name = "full-covariance gaussian mixture"
    cfg = O.GMVAEConfig(G, L, Kc, hidden, lik, R, S, True, kl_weight=0.8,
                        prior_probabilities_method=prior, latent_distribution=name)
    params = O.gmvae_init_params(cfg, seed=3, dtype=torch.float64)
    gen = torch.Generator().manual_seed(4)
    for k in params:
        if k.endswith("biases") or k.endswith("beta") or k == "Y/P/LOGITS":
            params[k] = torch.randn(params[k].shape, generator=gen, dtype=torch.float64) * 0.2
    assert params["Z/Q/MULTIVARIATE_GAUSSIAN/SCALES/DENSE/weights"].shape == (5, 6)
    x = torch.tensor(numpy.minimum(O.synthetic_counts(B, G, n_types=3, seed=5)[0], 40.0),
                     dtype=torch.float64)
    eps = torch.randn(Kc, R * S, B, L, generator=gen, dtype=torch.float64)
    state = O.AdamState(params)
    reference = {k: v.clone() for k, v in params.items()}
    out, grads = O.train_step(cfg, reference, state, x, x, eps, 1e-3, warm_up_weight=0.7)
    eng = GMVAEEngine(G, L, Kc, hidden, lik, True, 0.8, prior, None, 0.0, tensor_cores=False,
                      head_buffer_bytes=head_buffer_bytes, latent_distribution=name)
    eng.import_parameters(params)
    assert sorted(eng.export_parameters()) == sorted(params)          # same variables, same names
    plan = eng._plan(B, R * S)
    eng.set_batch_dense(plan, x.float())
    plan.eps.copy_(eps.reshape(-1, L).float())
    bound = eng.train_step(plan, R, S, 1e-3, warm_up_weight=0.7)
    for i, key in enumerate(["lower_bound", "lower_bound_weighted", "reconstruction_error",
                             "kl_divergence_z", "kl_divergence_y"]):
        assert abs(bound[i].item() - out[key].item()) <= 5e-5 * abs(out[key].item()) + 1e-6, key
    got = eng.export_gradients()
    gmax = max(g.abs().max().item() for g in grads.values())
    for key, g in grads.items():
        error = (got[key].double() - g).abs().max().item()
        assert error <= 3e-4 * g.abs().max().item() + 1e-5 * gmax, (key, error)
    eng.forward(plan, False, R, S, 1.0)
    now = {k: v.double() for k, v in eng.export_parameters().items()}
    ref_eval = O.gmvae_forward(cfg, now, x, x, eps, is_training=False)
    assert abs(plan.bound[0].item() - ref_eval["lower_bound"].item()) <= 5e-5 * abs(
        ref_eval["lower_bound"].item())
    z_mean = eng.z_mean(plan)
    assert (z_mean.double() - ref_eval["z_mean"]).abs().max().item() <= 1e-4
