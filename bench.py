#!/usr/bin/env python
"""Benchmark of the scVAE training hot path (BASELINE.json metric):
cells/sec of VAE training, negative-binomial likelihood, 20 k genes, on N B200s.

    python bench.py --gpus N --steps K --warmup W            # this repository (CUDA path)
    python bench.py --impl reference --steps K --warmup W    # reference CPU path (restated)

A "step" is one pass of the hot path over one minibatch: CSR gather/densify -> encoder ->
reparameterise -> decoder -> NB log-likelihood -> backward -> [all-reduce] -> clip + Adam,
exactly what the reference runs per ``session.run([optimiser, lower_bound])``
(scvae/models/variational_autoencoder.py:987-1029).  Workload = BASELINE.json configs[1]:
68 000 cells x 20 000 genes (10x-PBMC-shaped synthetic, ~7 % non-zero), NB, latent 50,
hidden [100] (reference default), R = S = 1.  One JSON line is printed by rank 0.
"""

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cells/sec VAE training (NB, 20k genes)"
UNIT = "cells/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=68000)
    ap.add_argument("--genes", type=int, default=20000)
    ap.add_argument("--latent", type=int, default=50)
    ap.add_argument("--hidden", type=int, nargs="*", default=[100])
    ap.add_argument("--likelihood", default="negative binomial")
    ap.add_argument("--minibatch", type=int, default=4096, help="cells per step per GPU")
    ap.add_argument("--density", type=float, default=0.07)
    ap.add_argument("--exchange", default=None, choices=["p2p", "nccl"],
                    help="N > 1: fused peer-memory exchange + optimiser (default) or NCCL all-reduce")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--feeder", default=None, choices=["host", "device", "batch", "hybrid"],
                    help="hotloop.PackedStream feeder of the e2e leg (default: the stream's own choice)")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the C3 / C4 shaped extra configurations (BASELINE.json configs[2..3])")
    ap.add_argument("--no-parity", action="store_true",
                    help="skip the untimed oracle check of one timed-path minibatch")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# synthetic data: ZINB-flavoured sparse counts, built as CSR
# ---------------------------------------------------------------------------------------------
def make_csr(n_cells, n_genes, density, seed, device=None):
    """Returns scipy CSR (host).  Generated on the GPU when one is given (fast), else numpy."""
    import scipy.sparse
    if device is not None:
        gen = torch.Generator(device=device).manual_seed(seed)
        gene_rate = torch.rand(n_genes, generator=gen, device=device) * 2.0 * density
        indptr = [torch.zeros(1, dtype=torch.int64, device=device)]
        cols, vals = [], []
        chunk = 4096
        total = 0
        for s in range(0, n_cells, chunk):
            rows = min(chunk, n_cells - s)
            mask = torch.rand(rows, n_genes, generator=gen, device=device) < gene_rate
            nz = mask.nonzero()
            counts = torch.bincount(nz[:, 0], minlength=rows)
            indptr.append(total + torch.cumsum(counts, 0))
            total += int(nz.shape[0])
            cols.append(nz[:, 1].to(torch.int32))
            # geometric-ish count values >= 1 (mostly 1-3, heavy tail)
            u = torch.rand(nz.shape[0], generator=gen, device=device)
            vals.append(torch.floor(1.0 - torch.log(u) * 1.2).clamp_(1, 500))
        indptr = torch.cat(indptr).cpu().numpy()
        indices = torch.cat(cols).cpu().numpy()
        data = torch.cat(vals).cpu().numpy().astype(numpy.float32)
    else:
        rng = numpy.random.RandomState(seed)
        gene_rate = rng.rand(n_genes) * 2.0 * density
        indptr = [0]
        cols, vals = [], []
        for s in range(0, n_cells, 1024):
            rows = min(1024, n_cells - s)
            mask = rng.rand(rows, n_genes) < gene_rate
            r, c = numpy.nonzero(mask)
            indptr.extend((indptr[-1] + numpy.cumsum(numpy.bincount(r, minlength=rows))).tolist())
            cols.append(c.astype(numpy.int32))
            vals.append(numpy.clip(numpy.floor(1.0 - numpy.log(rng.rand(len(c))) * 1.2), 1, 500))
        indptr = numpy.asarray(indptr, dtype=numpy.int64)
        indices = numpy.concatenate(cols)
        data = numpy.concatenate(vals).astype(numpy.float32)
    return scipy.sparse.csr_matrix((data, indices, indptr), shape=(n_cells, n_genes))


# ---------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self.proc = None
        self.marks = []

    def mark(self):
        """Bracket the timed region: samples between the first and the last mark are reported
        (the sampler itself starts before the warm-up, so short regions still see samples)."""
        self.marks.append(time.perf_counter())

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "5"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            # nvidia-smi needs a moment before its first line: wait for it (bounded), so that a
            # timed region of ~0.1 s is not over before the sampler has started printing
            t0 = time.perf_counter()
            while not self.samples and time.perf_counter() - t0 < 5.0 and self.proc.poll() is None:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        picked = self.samples
        if len(self.marks) >= 2:
            # a sample describes the interval before it was printed: keep those printed inside the
            # timed region or right behind it; fall back to the samples under load (warm-up
            # included) when the region was shorter than one sampling period
            lo, hi = self.marks[0], self.marks[-1] + 0.006
            inside = [s for s in self.samples if lo <= s[0] <= hi]
            picked = inside if inside else [s for s in self.samples if s[0] <= hi][-4:]
        for _, s in picked:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(numpy.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# reference CPU path (restated: the oracle, see oracle/scvae_oracle.py header)
# ---------------------------------------------------------------------------------------------
def cpu_reference_run(args, csr, steps, warmup, budget_s):
    """Times the restated reference training step on the host cores: per step
    ``x[idx].toarray()`` twice (VAE:997-998), forward, autograd backward, clip, Adam."""
    from oracle import scvae_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.VAEConfig(args.genes, args.latent, args.hidden, args.likelihood)
    params = O.vae_init_params(cfg, seed=0, dtype=torch.float32)
    state = O.AdamState(params)
    rng = numpy.random.RandomState(2)
    gen = torch.Generator().manual_seed(1)
    B = min(args.minibatch, csr.shape[0])
    n = csr.shape[0]

    def one(b):
        idx = rng.randint(0, n, size=b)
        x = torch.from_numpy(csr[idx].toarray())
        t = torch.from_numpy(csr[idx].toarray())
        eps = torch.randn(1, b, args.latent, generator=gen)
        out, _ = O.train_step(cfg, params, state, x, t, eps, 1e-4)
        return float(out["lower_bound"])

    t0 = time.perf_counter()
    one(B)
    first = time.perf_counter() - t0
    # bound the whole run: shrink the per-step sample if needed
    if first * (steps + warmup) > budget_s:
        B = max(256, int(B * budget_s / (first * (steps + warmup))))
    for _ in range(max(warmup - 1, 0)):
        one(B)
    t0 = time.perf_counter()
    for _ in range(steps):
        one(B)
    dt = time.perf_counter() - t0
    return {"value": steps * B / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "{} timed minibatches of {} cells (+{} warm-up) of the same workload; "
                      "restated reference CPU path (PyTorch-CPU fp32 oracle; TF 1.15 is not "
                      "installable, SURVEY 8c)".format(steps, B, warmup)}, dt / steps * 1e3, B


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    csr = make_csr(min(args.cells, 16384), args.genes, args.density, seed=60)
    base, ms, B = cpu_reference_run(args, csr, args.steps, max(args.warmup, 1), budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, B), "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, minibatch):
    return {
        "workload": "C2: VAE, {} cells x {} genes synthetic (10x-PBMC-shaped, {:.0f}% non-zero), "
                    "{}, latent {}, hidden {}, R=S=1".format(
                        args.cells, args.genes, args.density * 100, args.likelihood, args.latent,
                        args.hidden),
        "minibatch_per_gpu": minibatch,
        "l2": "per-step working set (~2.6 GB of (cells x genes) fp32 tensors at B=4096) exceeds "
              "the 126 MB L2; minibatch rows change every step",
    }


# ---------------------------------------------------------------------------------------------
# parity of the benchmarked path against the oracle (checker only, outside every timed region)
# ---------------------------------------------------------------------------------------------
def parity_check(args, eng, loop, data, csr, rows, compare=True):
    """Replays the captured training step (the object the timed region replays) on ``rows`` and
    compares ELBO / reconstruction error / KL / per-cell log p / per-cell latent means with
    ``oracle.train_step`` (fp64) on the same rows, variables and reparameterisation noise.
    With N > 1 ranks the step includes the gradient exchange; the quantities compared are this
    rank's forward results, which do not depend on it."""
    from oracle import scvae_oracle as O
    B, L = rows.numel(), args.latent
    before = {k: v.double() for k, v in eng.export_parameters().items()}
    loop.rows.copy_(rows)
    bound = loop.step(data, 1e-4, 1.0)
    torch.cuda.synchronize()
    plan = loop.plan
    if compare is False:          # ranks > 0: they only take part in the step's gradient exchange
        return None
    bound = bound.cpu().numpy().astype(numpy.float64)
    eps = plan.eps.cpu().double().reshape(1, B, L)
    x = torch.from_numpy(csr[rows.cpu().numpy()].toarray()).double()
    cfg = O.VAEConfig(args.genes, L, args.hidden, args.likelihood)
    torch.set_num_threads(os.cpu_count() or 1)
    out = O.vae_forward(cfg, before, x, x, eps, is_training=True)
    names = ["lower_bound", "lower_bound_weighted", "reconstruction_error", "kl_divergence"]
    rel = {n: abs(bound[i] - out[n].item()) / abs(out[n].item()) for i, n in enumerate(names)}
    lp_ref = out["log_p_x_given_z"].reshape(-1)
    lp_rel = ((plan.logp.cpu().double() - lp_ref).abs().max() / lp_ref.abs().max()).item()
    mu_ref = out["q_z_mean"]
    mu_rel = ((plan.PH[:, :L].cpu().double() - mu_ref).abs().max() / mu_ref.abs().max()).item()
    tol = 1e-3
    return {"checked": "one CUDA-graph replay of the timed training step vs oracle.vae_forward "
                       "(fp64) on the same {} rows, variables and noise".format(B),
            "elbo_gpu": bound[0], "elbo_oracle": out["lower_bound"].item(),
            "rel_err": {"elbo": rel["lower_bound"], "enre": rel["reconstruction_error"],
                        "kl": rel["kl_divergence"], "per_cell_log_p_max": lp_rel,
                        "per_cell_latent_mean_max": mu_rel},
            "tolerance": tol,
            "pass": bool(max(rel.values()) <= tol and lp_rel <= tol and mu_rel <= tol),
            "fused_16bit_path": bool(plan.fused_done)}


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[2] and configs[3] at their stated shapes (not the headline metric): a short
# timed run each, with the same oracle check as the headline configuration
# ---------------------------------------------------------------------------------------------
def _time_steps(step, steps, warmup):
    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def _timed_kernels(K, names, run, repeats):
    """Per-call CUDA-event time of the named C-ABI wrappers over `repeats` eager runs."""
    evs = {k: [] for k in names}
    originals = {k: getattr(K, k) for k in names}

    def wrap(name):
        def timed(*a, **kw):
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record()
            originals[name](*a, **kw)
            e_.record()
            evs[name].append((s_, e_))
        return timed
    for k in names:
        setattr(K, k, wrap(k))
    try:
        for i in range(repeats):
            run(i)
        torch.cuda.synchronize()
    finally:
        for k in names:
            setattr(K, k, originals[k])
    return {k: sum(s_.elapsed_time(e_) for s_, e_ in v) / repeats for k, v in evs.items() if v}


def evaluate_reconstruction(args, dev, csr):
    """f2 (VAE:1939-2052): one `evaluate` pass with the reconstruction at the C2 shape, on a
    16 384-cell shard (a 1.3 GB fp32 result in pinned memory): forward with the evaluation-mode
    graph, moments kernel straight into the staging buffers of hotloop.ReconstructionSink, D2H of
    p_x_mean behind the next minibatch.  cells/s INCLUDING the device -> host copy."""
    from scvae_b200.variational_autoencoder import VariationalAutoencoder
    from scvae_b200.hotloop import ResidentCSR
    n = min(16384, csr.shape[0])
    sub = csr[:n]
    B = min(args.minibatch, n)
    model = VariationalAutoencoder(feature_size=args.genes, latent_size=args.latent,
                                   hidden_sizes=list(args.hidden), reconstruction_distribution=args.likelihood,
                                   log_directory="/tmp/scvae_b200_bench_log", seed=0)
    engine = model._get_engine()
    data = ResidentCSR(sub, dev)
    out = {}
    for dtype in ("float32", "float16"):
        best = None
        for rep in range(2):                 # the first pass allocates plans and pins the result
            collect, sink, _, _ = model._reconstruction_collector(engine, n, B, 1, 1, True, [], dtype)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            model._evaluate_pass(engine, data, data, B, 1, 1, deterministic=True, on_batch=collect)
            values = sink.finish()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        out[dtype] = {"cells_per_s": n / best, "seconds": best,
                      "d2h_bytes": int(sink.bytes_copied), "d2h_gb_per_s": sink.bytes_copied / best / 1e9,
                      "finite": bool(numpy.isfinite(values[::97]).all())}
    out["what"] = ("VariationalAutoencoder._evaluate_pass + ReconstructionSink on {} cells x {} genes, "
                   "minibatch {}, wall clock of the whole pass incl. the D2H of p_x_mean "
                   "(second pass; the first allocates)").format(n, args.genes, B)
    return out


def extra_configs(dev):
    """C3 shape: VAE, zero-inflated NB, 28 000 genes, latent 100 (a 32 768-cell shard of the 1.3 M
    cells: a step touches one minibatch, so the per-step work is that of the full matrix);
    C4 shape: GMVAE, K = 20 clusters, 20 000 genes, NB, latent 50.  Each: cells/s of CUDA-graph
    replayed training steps, the dominant kernel, and the oracle check of one replayed step."""
    from oracle import scvae_oracle as O
    from scvae_b200 import kernels as K
    from scvae_b200.engine import VAEEngine
    from scvae_b200.gmvae_engine import GMVAEEngine
    from scvae_b200.hotloop import ResidentCSR, TrainLoop
    out = []
    heavy = ["heads_fused_bwd", "gemm_f16", "gemm_f16_split", "gemm", "csr_densify", "adam_clip_step",
             "vae_mid_fwd", "vae_mid_bwd", "likelihood_bwd"]

    def one(name, make_engine, cfg, forward, n_cells, G, L, B, parity_B, steps, bound_names):
        csr = make_csr(n_cells, G, 0.07, seed=61, device=dev)
        data = ResidentCSR(csr, dev)
        eng = make_engine()
        loop = TrainLoop(eng, B, seed=5, use_graph=True)
        n_batches = n_cells // B
        perm = torch.from_numpy(numpy.random.RandomState(3).permutation(n_cells)).to(dev)

        def step(i):
            b = i % n_batches
            loop.rows.copy_(perm[b * B:(b + 1) * B])
            return loop.step(data, 1e-4, 1.0)
        ms = _time_steps(step, steps, 3)
        loop.use_graph = False
        overlap, eng.overlap_streams = eng.overlap_streams, False
        per = _timed_kernels(K, heavy, step, 3)
        eng.overlap_streams = overlap
        loop.use_graph = True
        top = max(per, key=per.get)
        entry = {"workload": name, "minibatch": B, "steps": steps, "ms_per_step": ms,
                 "value": B / (ms * 1e-3), "unit": UNIT,
                 "dominant_kernel": {"entry_point": top, "ms_per_step": round(per[top], 4),
                                     "share_of_step": round(per[top] / ms, 3)},
                 "eager_ms_per_step_by_kernel": {k: round(v, 4) for k, v in per.items()}}
        # oracle check of one replayed step on a smaller minibatch of the same data and model
        ploop = TrainLoop(eng, parity_B, seed=9, use_graph=True)
        rows = perm[:parity_B]
        ploop.rows.copy_(rows)
        before = {k: v.double() for k, v in eng.export_parameters().items()}
        bound = ploop.step(data, 1e-4, 1.0)
        torch.cuda.synchronize()
        bound = bound.cpu().numpy().astype(numpy.float64)
        plan = ploop.plan
        x = torch.from_numpy(csr[rows.cpu().numpy()].toarray()).double()
        eps = plan.eps.cpu().double()
        ref = forward(cfg, before, x, eps, parity_B)
        rel = {n: abs(bound[i] - ref[n].item()) / (abs(ref[n].item()) + 1e-30)
               for i, n in enumerate(bound_names)}
        lp_ref = ref["log_p_x_given_z"].reshape(-1)
        lp_rel = ((plan.logp.cpu().double()[:lp_ref.numel()] - lp_ref).abs().max()
                  / lp_ref.abs().max()).item()
        entry["parity"] = {"minibatch": parity_B, "rel_err": dict(rel, per_cell_log_p_max=lp_rel),
                           "tolerance": 1e-3,
                           "pass": bool(max(rel.values()) <= 1e-3 and lp_rel <= 1e-3),
                           "fused_16bit_path": bool(plan.fused_done)}
        out.append(entry)
        del loop, ploop, eng, data
        torch.cuda.empty_cache()

    G3, L3 = 28000, 100
    lik3 = "zero-inflated negative binomial"
    cfg3 = O.VAEConfig(G3, L3, [100], lik3)
    one("C3 shape: VAE, 1.3M x 28000 genes synthetic (32768-cell shard resident), zero-inflated "
        "negative binomial, latent 100, hidden [100], R=S=1",
        lambda: VAEEngine(G3, L3, [100], lik3, device=dev, seed=0), cfg3,
        lambda cfg, prm, x, eps, b: O.vae_forward(cfg, prm, x, x, eps.reshape(1, b, L3), is_training=True),
        32768, G3, L3, 4096, 1024, 30,
        ["lower_bound", "lower_bound_weighted", "reconstruction_error", "kl_divergence"])
    # the headline configuration at two other minibatch sizes: twice the benchmarked one (the fixed
    # per-step latencies amortise further) and the reference's default of 100 cells (pure latency)
    G2, L2 = 20000, 50
    cfg2 = O.VAEConfig(G2, L2, [100], "negative binomial")
    for B2, pB2, steps2 in ((8192, 512, 30), (100, 100, 200)):
        one("C2 shape at minibatch {}: VAE, 68000 x 20000 genes synthetic (16384-cell shard resident), "
            "negative binomial, latent 50, hidden [100], R=S=1".format(B2),
            lambda: VAEEngine(G2, L2, [100], "negative binomial", device=dev, seed=0), cfg2,
            lambda cfg, prm, x, eps, b: O.vae_forward(cfg, prm, x, x, eps.reshape(1, b, L2), is_training=True),
            16384, G2, L2, B2, pB2, steps2,
            ["lower_bound", "lower_bound_weighted", "reconstruction_error", "kl_divergence"])
    G4, L4, K4 = 20000, 50, 20
    cfg4 = O.GMVAEConfig(G4, L4, K4, [100], "negative binomial", 1, 1, True)
    one("C4 shape: GMVAE, K=20 clusters, 68000 x 20000 genes synthetic (16384-cell shard resident), "
        "negative binomial, latent 50, hidden [100], R=S=1",
        lambda: GMVAEEngine(G4, L4, K4, [100], "negative binomial", device=dev, seed=0), cfg4,
        lambda cfg, prm, x, eps, b: O.gmvae_forward(cfg, prm, x, x, eps.reshape(K4, 1, b, L4), is_training=True),
        16384, G4, L4, 1024, 128, 10,
        ["lower_bound", "lower_bound_weighted", "reconstruction_error", "kl_divergence_z",
         "kl_divergence_y"])
    return out


# ---------------------------------------------------------------------------------------------
# this repository's arm
# ---------------------------------------------------------------------------------------------
def measure_blocks(timed_block, world, device, budget_ms=250.0, most=25):
    """``timed_block(b)`` -> milliseconds of block ``b`` (already the maximum over the ranks).  Block
    0 is the contract's region; while the measurement is short, further blocks follow until about
    ``budget_ms`` in all (at most ``most`` blocks).  Every rank runs the SAME number of blocks:
    rank 0's count is broadcast (the blocks contain barriers and collectives)."""
    import torch.distributed as dist
    blocks = [timed_block(0)]
    more = torch.tensor([max(0, min(most, int(budget_ms / max(blocks[0], 1e-3))) - 1)],
                        dtype=torch.int64, device=device)
    if world > 1:
        dist.broadcast(more, src=0)
    for b in range(int(more.item())):
        blocks.append(timed_block(b + 1))
    return blocks


def _dbg(msg):
    if os.environ.get("SCVAE_BENCH_DEBUG"):
        print("[bench rank {}] {}".format(os.environ.get("RANK", "0"), msg), file=sys.stderr,
              flush=True)


def run_b200(args):
    import torch.distributed as dist
    from scvae_b200 import _lib
    from scvae_b200.engine import VAEEngine
    from scvae_b200.hotloop import PackedStream, ResidentCSR, TrainLoop
    from scvae_b200 import kernels as K

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the scVAE hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    # ---- data (each rank owns its own shard of `cells` cells: weak scaling) ---------------
    csr = make_csr(args.cells, args.genes, args.density, seed=60 + rank, device=dev)
    B = min(args.minibatch, args.cells)
    data = ResidentCSR(csr, dev)
    eng = VAEEngine(args.genes, args.latent, args.hidden, args.likelihood, device=dev, seed=0)
    exchange = "none"
    if world > 1:
        from scvae_b200 import distributed as D
        D.attach(eng, exchange=args.exchange)
        exchange = "p2p (fused reduce-scatter + Adam + all-gather over NVLink peer memory)" \
            if eng._peer is not None else "nccl all-reduce"
    loop = TrainLoop(eng, B, seed=1 + rank, use_graph=not args.no_graph)
    n_batches = args.cells // B
    perm = torch.from_numpy(numpy.random.RandomState(2).permutation(args.cells)).to(dev)

    def step(i):
        b = i % n_batches
        loop.rows.copy_(perm[b * B:(b + 1) * B])
        return loop.step(data, 1e-4, 1.0)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    _dbg("data + engine ready")
    # count our kernel launches per step (eager, outside the timed region)
    step(0)
    _dbg("first step (graph captured) done")
    torch.cuda.synchronize()
    l0 = lib.scvae_launch_count()
    loop.use_graph, saved = False, loop.use_graph
    step(0)
    torch.cuda.synchronize()
    launches_per_step = lib.scvae_launch_count() - l0
    loop.use_graph = saved

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        step(i)
    sync_all()
    _dbg("warm-up done")
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_block(first):
        """EXACTLY `steps` steps between two device events, barrier + synchronize on both sides,
        the maximum over the ranks."""
        sync_all()
        e0.record()
        for i in range(args.steps):
            step(first + i)
        e1.record()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # The contract's region: `steps` steps behind the warm-up.  A region of 20 steps lasts 10 ms, so
    # while the measurement is short further blocks of exactly `steps` steps follow (about 0.25 s in
    # all) and the MEDIAN block is reported; every block is listed in `timed_blocks`.
    blocks = measure_blocks(lambda b: timed_block(args.warmup + b * args.steps), world, dev)
    sampler.mark()
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = float(numpy.median(blocks))
    _dbg("timed region done")
    bound = loop.plan.bound.cpu().tolist()
    value = args.steps * B * world / (elapsed_ms * 1e-3)

    # ---- dominant kernel, timed live with CUDA events (eager steps, same workload) ---------
    # every C-ABI entry point is bracketed with events for a few eager steps; the family with
    # the largest share of the step is the roofline kernel
    P = eng.P
    cells_genes = float(B) * args.genes
    families = {
        # name: (algorithmic bytes per launch, description)
        "heads_fused_bwd": ((2 + 4 * 2 * P) * cells_genes,
                            "heads_fused_kernel<NB> (heads GEMM + NB log-prob + gradient + dgrad, "
                            "one kernel)"),
        "likelihood_bwd": ((1 + 2 * P) * 4.0 * cells_genes,
                           "likelihood_kernel<NB, BWD> (fused log-prob + gradient)"),
        "csr_densify": (None, "csr_densify_kernel"),
        "gemm": (None, "gemm_tc_kernel (tf32)"),
        "gemm_f16": (None, "gemm_tc_kernel (fp16)"),
        "adam_clip_step": (7 * 4.0 * eng.store.total, "adam_clip_kernel"),
        "gemm_f16_split": (None, "gemm_tc_kernel (fp16, split operand: first layer forward / weight gradient)"),
        "vae_mid_fwd": (None, "vae_mid_fwd_kernel (hidden layers, posterior, sample, KL)"),
        "vae_mid_bwd": (None, "vae_mid_bwd_kernel (their backward + the bound)"),
    }
    evs = {k: [] for k in families}
    originals = {k: getattr(K, k) for k in families}

    def wrap(name):
        def timed(*a, **kw):
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record()
            originals[name](*a, **kw)
            e_.record()
            evs[name].append((s_, e_))
        return timed
    for k in families:
        setattr(K, k, wrap(k))
    loop.use_graph = False
    overlap_saved, eng.overlap_streams = eng.overlap_streams, False   # serial launches: clean per-kernel times
    n_eager = min(args.steps, 10)
    for i in range(n_eager):
        step(args.warmup + args.steps + i)
    torch.cuda.synchronize()
    for k in families:
        setattr(K, k, originals[k])
    eng.overlap_streams = overlap_saved
    loop.use_graph = saved
    per_step_ms = {k: sum(s_.elapsed_time(e_) for s_, e_ in v) / n_eager for k, v in evs.items() if v}
    launches = {k: len(v) / n_eager for k, v in evs.items() if v}
    candidates = [k for k in per_step_ms if families[k][0] is not None and k != "adam_clip_step"]
    top = max(candidates, key=lambda k: per_step_ms[k])
    lik_ms = per_step_ms[top] / launches[top]
    # Algorithmic bytes per (cell, gene) of the likelihood stage, SURVEY 8d: targets (2 B when
    # stored in 16 bits, else 4) + P fp32 parameters read + P fp32 gradients written.
    lik_bytes = families[top][0] + 8.0 * B
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = json.load(open(peaks_path))["hbm_gbs"]
        peak_src = "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = lik_bytes / (lik_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("minibatch") == B and tj.get("genes") == args.genes:
                traffic = tj.get(top + "_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"kernel": families[top][1], "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": lik_bytes,
                "launch_ms": lik_ms, "share_of_step": per_step_ms[top] / (elapsed_ms / args.steps),
                "dram_traffic_frac_of_peak": (traffic / (lik_ms * 1e-3) / 1e9 / peak) if traffic else None,
                "note": ("algorithmic bytes = SURVEY 8d per-(cell,gene) figure of the likelihood "
                         "stage (16-bit targets + P fp32 parameters + P fp32 gradients); the fused "
                         "kernel keeps the parameters on chip, so its real DRAM traffic (`traffic`) is "
                         "a third of that: frac is the rate an unfused HBM-bound implementation would "
                         "have to stream at to match it (it can exceed 1), dram_traffic_frac_of_peak "
                         "is the share of HBM peak the kernel really uses; its limiter is FP32/MUFU "
                         "instruction issue (profiles/)" if top == "heads_fused_bwd" else
                         "HBM-bound streaming kernel"),
                "eager_ms_per_step_by_kernel": {k: round(v, 4) for k, v in per_step_ms.items()}}
    _dbg("kernel timing done")

    # ---- forward-only evaluation pass (the reference runs one over the training set after every
    # epoch, VAE:1092-1150): 16-bit minibatch + forward-only fused heads, eager launches
    eval_pass = None
    try:
        n_eval = min(n_batches, 16)
        plan = loop.plan

        def eval_batch(b):
            eng.set_batch_csr(plan, data.indptr, data.indices, data.values, perm[b * B:(b + 1) * B],
                              u16_ok=data.u16_ok, f16_exact=data.f16_exact, train16=True,
                              row_const_all=data.row_const)
            K.fill_normal(plan.eps, 7, b)
            eng.forward(plan, False, 1, 1, 1.0, keep_heads=False)

        eval_batch(0)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for b in range(n_eval):
            eval_batch(b)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / n_eval
        eval_pass = {"value": B * world / (ms * 1e-3), "unit": UNIT, "ms_per_batch": ms,
                     "what": "forward-only evaluation pass per GPU x n_gpus (not part of `value`)"}
    except Exception as exc:      # informational only
        eval_pass = {"error": str(exc)}
    _dbg("evaluation pass timed")

    # ---- end to end: host CSR in pinned memory, per-step H2D of the row slab, D2H of ELBO ---
    e2e = None
    if not args.no_e2e:
        # the product path for host-resident data (VariationalAutoencoder.train(...,
        # data_residency="host")): the epoch's rows, in shuffled order, as packed slabs in pinned
        # memory (hotloop.PackedStream, ~2 bytes per non-zero); one host -> device copy per step
        stream = PackedStream(csr, dev, B, feeder=args.feeder)
        # shuffled epochs back to back, full minibatches only: the feeder thread assembles each
        # step's slab in pinned memory (scvae_pack_row_slab) WHILE the timed steps run
        full = n_batches * B
        need = args.warmup + args.steps + 1
        rs = numpy.random.RandomState(17 + rank)
        epochs = [perm.cpu().numpy()[:full]] + [rs.permutation(args.cells)[:full]
                                                for _ in range(-(-need // n_batches) - 1)]
        stream.pack_epoch(numpy.concatenate(epochs)[:need * B])
        compute = torch.cuda.current_stream()
        h2d = 0

        # The bound of every step is read back by the host, as `session.run` returns it; the copy
        # goes to pinned memory right behind the step and the host waits for it one step later
        # (the ELBO is only logged), so the read never drains the GPU queue.
        elbo_host = [torch.zeros(4, dtype=torch.float32).pin_memory() for _ in range(2)]
        elbo_done = [torch.cuda.Event(), torch.cuda.Event()]
        elbos = []

        def e2e_step(i, pending):
            nonlocal h2d
            slot = pending
            nxt = stream.fetch((i + 1) % 2, i + 1)
            compute.wait_event(slot["ready"])
            out = loop.step(slot, 1e-4, 1.0)
            slot["free"].record(compute)
            h2d += slot["bytes"]
            elbo_host[i % 2].copy_(out, non_blocking=True)     # D2H of the step's result
            elbo_done[i % 2].record(compute)
            if i > 0:
                elbo_done[(i - 1) % 2].synchronize()
                elbos.append(float(elbo_host[(i - 1) % 2][0]))
            return nxt, None

        sync_all()            # (ranks finish encoding their shards at different times)
        _dbg("e2e: streamed CSR ready")
        pending = stream.fetch(0, 0)
        for i in range(args.warmup):
            pending, _ = e2e_step(i, pending)
        sync_all()
        h2d = 0
        t0 = time.perf_counter()
        for i in range(args.warmup, args.warmup + args.steps):
            pending, _ = e2e_step(i, pending)
        torch.cuda.synchronize()
        elbos.append(float(elbo_host[(args.warmup + args.steps - 1) % 2][0]))   # last step's result
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        stream.close()
        e2e = {"value": args.steps * B * world / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(h2d / args.steps), "d2h_bytes_per_step": 16,
               "path": "hotloop.PackedStream (the feeder of train(..., data_residency='host')): the "
                       "matrix stays in host memory as per-row strings (encoded once per data set); "
                       "every step a feeder thread gathers the shuffled minibatch's strings into a "
                       "pinned slab (scvae_pack_row_slab, {} host threads, INSIDE the timed region) "
                       "-> ONE H2D per step (copy stream, double-buffered) -> "
                       "scvae_csr_densify_packed -> train step -> D2H of the bound (async to pinned "
                       "memory, read one step later)".format(stream.pack_threads),
               "feeder": stream.feeder,
               "host_pack_ms_per_step": round(1e3 * stream.pack_seconds / max(need, 1), 3),
               "bytes_per_nonzero": round(stream.bytes_per_nonzero, 3),
               "encode_once_seconds": round(stream.encode_seconds, 2),
               "encode_note": "the per-row strings are encoded once per data set (numpy), outside "
                              "the timed steps; the per-step gather, shuffle included, is inside"}

    # ---- parity of the timed path (untimed): one more replay of the SAME captured step on the
    # rows of the first timed minibatch, checked against the oracle on those rows / weights / noise
    parity = None
    if not args.no_parity:
        # (every rank replays the step -- it contains the gradient exchange; rank 0 compares)
        try:
            parity = parity_check(args, eng, loop, data, csr, perm[args.warmup % n_batches * B:
                                                                   (args.warmup % n_batches + 1) * B],
                                  compare=rank == 0)
        except Exception as exc:      # reported, never silently dropped
            parity = {"error": repr(exc)}
    replicas = None
    if world > 1:
        # replica identity after the timed steps: every rank must hold the same variables
        # (max |theta_r - theta_0| over ranks and entries, via one max- and one min-all-reduce) and
        # step counter.  The Adam slots are compared after PeerExchange.gather_slots: between
        # checkpoints each rank keeps only its own slice of them live (the fused exchange shards
        # the optimiser state).
        torch.cuda.synchronize()

        def spread(buf):
            hi, lo = buf.clone(), buf.clone()
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            return float((hi - lo).abs().max().item())
        worst = spread(eng.store.param)
        peer = getattr(eng, "_peer", None)
        if peer is not None and getattr(eng, "_last_ranges", None):
            peer.gather_slots(eng._last_ranges)
        worst_slots = max(spread(eng.store.m), spread(eng.store.v))
        steps_spread = spread(eng.store.step.clone().float())
        failed = bool(peer is not None and peer.timed_out())
        replicas = {"identical": bool(worst == 0.0 and worst_slots == 0.0 and steps_spread == 0.0
                                      and not failed),
                    "max_abs_difference_variables": worst,
                    "max_abs_difference_adam_slots": worst_slots,
                    "exchange_timed_out": failed}
        dist.barrier()

    extra = None
    evaluate_block = None
    if rank == 0 and world == 1 and not args.no_extra:
        # free the headline configuration's buffers first
        try:
            extra = extra_configs(dev)
        except Exception as exc:
            extra = {"error": repr(exc)}
        try:
            evaluate_block = evaluate_reconstruction(args, dev, csr)
        except Exception as exc:
            evaluate_block = {"error": repr(exc)}

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sub = csr[:min(args.cells, 16384)]
        cpu_base, _, _ = cpu_reference_run(args, sub, steps=2, warmup=1, budget_s=40.0)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "timed_blocks": {"steps_per_block": args.steps, "ms": [round(b, 4) for b in blocks],
                             "reported": "median block (the first one is the region right behind "
                                         "the warm-up)"},
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (f16 tensor-core operands, f32 accumulate)",
            "data": "synthetic", "config": workload_config(args, B), "clocks": clocks,
            "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps),
            "roofline": roofline, "cpu_baseline": cpu_base,
            "evaluation_pass": eval_pass,
            "lower_bound_last_step": bound[0],
            "cuda_graph": bool(saved),
        }
        line["gradient_exchange"] = exchange     # (not in `config`: both arms name one workload)
        line["parity"] = parity
        line["replicas_identical"] = None if replicas is None else replicas["identical"]
        line["replicas"] = replicas
        line["extra_configs"] = extra
        line["evaluate_reconstruction"] = evaluate_block
        print(json.dumps(line), flush=True)
    if world > 1:
        # CUDA graphs that captured NCCL kernels must die before the communicator does; guard
        # the teardown with a watchdog so a stuck destroy can never hang the launcher
        torch.cuda.synchronize()
        dist.barrier()
        loop._graphs = {}
        loop._graph = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        watchdog = threading.Timer(15.0, lambda: os._exit(0))
        watchdog.daemon = True
        watchdog.start()
        dist.destroy_process_group()
        watchdog.cancel()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
