"""world_size-2 gloo tests of the data-parallel host logic (no GPU): sharding, permutation
broadcast and the exactness of 'sum of per-shard mean-gradients / W == global-batch gradient'."""
import os
import sys

import numpy
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, results):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from oracle import scvae_oracle as O
    from scvae_b200 import distributed as D
    torch.set_num_threads(1)
    r, w = D.initialise_from_environment(backend="gloo")
    assert (r, w) == (rank, world) and D.is_active()
    perm = D.broadcast_permutation(40, numpy.random.RandomState(100 + rank))
    # every rank must hold rank 0's permutation
    expected = numpy.random.RandomState(100).permutation(40)
    assert numpy.array_equal(perm.numpy(), expected)

    cfg = O.VAEConfig(12, 3, [6], "negative binomial", minibatch_normalisation=False)
    params = O.vae_init_params(cfg, seed=0, dtype=torch.float64)
    x = torch.tensor(O.synthetic_counts(40, 12, seed=1)[0], dtype=torch.float64).clamp(max=20)
    eps = torch.randn(1, 40, 3, generator=torch.Generator().manual_seed(2), dtype=torch.float64)
    B = 16
    lo, hi = D.shard_bounds(8, B, rank, world)
    assert hi - lo == B // world
    rows = perm[lo:hi]
    names = O.trainable_names(params)

    def grads(idx):
        leaves = {k: params[k].clone().requires_grad_(True) for k in names}
        local = {k: leaves.get(k, v) for k, v in params.items()}
        out = O.vae_forward(cfg, local, x[idx], x[idx], eps[:, idx], True)
        g = torch.autograd.grad(-out["lower_bound_weighted"], [leaves[k] for k in names])
        return torch.cat([t.reshape(-1) for t in g])

    flat = grads(rows)
    D.all_reduce_sum_(flat)
    flat /= world                                   # the optimiser kernel's grad_scale
    full = grads(perm[8:8 + B])
    results[rank] = (flat - full).abs().max().item() / full.abs().max().item()
    dist.destroy_process_group()


def test_two_rank_gradient_average_is_exact():
    port = 29500 + (os.getpid() % 2000)
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(_worker, args=(2, port, results), nprocs=2, join=True)
    assert len(results) == 2
    assert max(results.values()) < 1e-12


def test_shard_bounds_cover_minibatch():
    from scvae_b200 import distributed as D
    spans = [D.shard_bounds(100, 64, r, 4) for r in range(4)]
    assert spans == [(100, 116), (116, 132), (132, 148), (148, 164)]
    assert D.shard_bounds(0, 10, 1, 4) == (2, 4)     # remainder dropped: equal work per rank
    assert D.rank() == 0 and D.world_size() == 1 and not D.is_active()


def test_peer_exchange_slices_partition_every_range():
    """Ownership map of the fused peer exchange: the W slices of a flat range are disjoint,
    ordered, 16-byte aligned and cover it exactly (also when W does not divide the range)."""
    from scvae_b200.distributed import peer_slice_bounds
    for lo, hi in ((0, 4), (0, 6060288), (2023424, 6060288), (16, 16 + 4 * 37), (8, 8)):
        for world in (1, 2, 3, 4, 8, 16):
            cursor = lo
            for r in range(world):
                a, b = peer_slice_bounds(lo, hi, r, world)
                assert a == cursor and a <= b <= hi and (a - lo) % 4 == 0
                cursor = b
            assert cursor == hi


def _engine_worker(rank, world, port, results):
    """Two engine replicas (CPU, kernel stand-ins) train on their halves of one global minibatch
    through the engine's data-parallel wiring (all-reduce of the flat gradient buffer, 1/W scale
    inside the optimiser kernel, clip after the reduction)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import kernel_standins
    import scvae_b200.engine as E
    from oracle import scvae_oracle as O
    from scvae_b200 import distributed as D
    torch.set_num_threads(1)
    D.initialise_from_environment(backend="gloo")
    E.K = kernel_standins                       # this process only: test infrastructure
    G, L, H, B = 20, 3, [8, 6], 12
    cfg = O.VAEConfig(G, L, H, "zero-inflated negative binomial", minibatch_normalisation=False,
                      kl_weight=0.7)
    params = O.vae_init_params(cfg, seed=0, dtype=torch.float64)
    for k in params:
        if k.endswith("weights"):
            params[k] = params[k] * 0.3
    x = torch.tensor(O.synthetic_counts(B, G, seed=1)[0], dtype=torch.float64).clamp(max=6)
    eps = torch.randn(1, B, L, generator=torch.Generator().manual_seed(2), dtype=torch.float64)
    eng = E.VAEEngine(G, L, H, "zero-inflated negative binomial", "gaussian", False,
                      kl_weight=0.7, device="cpu", tensor_cores=False)
    eng.overlap_streams = False
    eng.import_parameters(params)
    eng.set_data_parallel(world, D.all_reduce_sum_)
    lo, hi = D.shard_bounds(0, B, rank, world)
    plan = eng._plan(hi - lo, 1)
    eng.set_batch_dense(plan, x[lo:hi].float())
    plan.eps.copy_(eps[0, lo:hi].float())
    for _ in range(2):                          # two steps: the replicas must stay identical
        eng.train_step(plan, 1, 1, 1e-3, warm_up_weight=0.5)
    got = eng.export_parameters()
    state = O.AdamState(params)
    for _ in range(2):                          # the same two steps on the whole minibatch
        _, grads = O.train_step(cfg, params, state, x, x, eps, 1e-3, warm_up_weight=0.5)
    worst = max((got[k].double() - v).abs().max().item() / max(v.abs().max().item(), 1.0)
                for k, v in params.items())
    flat = eng.store.param.clone()
    D.all_reduce_sum_(flat)
    results[rank] = (worst, (flat / world - eng.store.param).abs().max().item())
    dist.destroy_process_group()


def test_two_engine_replicas_match_the_global_minibatch_step():
    port = 31500 + (os.getpid() % 2000)
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(_engine_worker, args=(2, port, results), nprocs=2, join=True)
    assert len(results) == 2
    for worst, drift in results.values():
        assert worst < 2e-5          # fp32 engine state against the fp64 whole-minibatch step
        assert drift == 0.0          # replicas bit-identical after the exchange


def _sharded_train_worker(rank, world, port, results, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import scipy.sparse
    import shell_cpu_patch
    shell_cpu_patch.apply()
    from oracle import scvae_oracle as O
    from scvae_b200 import distributed as D
    from scvae_b200 import hotloop as H
    from scvae_b200.data_set import DataSet
    from scvae_b200.variational_autoencoder import VariationalAutoencoder
    torch.set_num_threads(1)
    D.initialise_from_environment(backend="gloo")
    x, labels = O.synthetic_counts(90, 24, n_types=3, seed=3, target_zero_fraction=0.7)
    ds = DataSet("toy", values=scipy.sparse.csr_matrix(numpy.minimum(x, 30.0)), labels=labels.astype(str))
    uploaded = []
    original = H.ResidentCSR.__init__

    def spy(self, matrix, device):
        original(self, matrix, device)
        uploaded.append(self.shape[0])
    H.ResidentCSR.__init__ = spy
    model = VariationalAutoencoder(feature_size=24, latent_size=3, hidden_sizes=[8],
                                   reconstruction_distribution="poisson",
                                   log_directory=os.path.join(tmp, "rank{}".format(rank)), seed=2)
    model.train(ds, number_of_epochs=2, minibatch_size=30, learning_rate=1e-2, shuffle_seed=4,
                data_sharding="rank")
    eng = model._get_engine()
    flat = eng.store.param.clone()
    D.all_reduce_sum_(flat)
    from scvae_b200 import model_utilities as MU
    curve = MU.load_learning_curves(model, "training")["lower_bound"] if rank == 0 else [0.0, 0.0]
    results[rank] = (uploaded[0], (flat / world - eng.store.param).abs().max().item(),
                     [float(v) for v in curve])
    dist.destroy_process_group()


def test_rank_sharded_resident_training_data(tmp_path):
    """train(..., data_sharding="rank") on two gloo ranks: each rank uploads only its cells r::W,
    the replicas stay identical, and the logged training ELBO is the all-rank aggregate."""
    port = 33500 + (os.getpid() % 2000)
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(_sharded_train_worker, args=(2, port, results, str(tmp_path)), nprocs=2, join=True)
    assert len(results) == 2
    assert results[0][0] == 45 and results[1][0] == 45         # half of the 90 cells each
    for _, drift, _ in results.values():
        assert drift == 0.0
    curve = results[0][2]
    assert len(curve) == 2 and all(numpy.isfinite(curve))


def _blocks_worker(rank, world, port, results):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = []

    def timed_block(b):
        # as bench.py's: a collective inside every block, the maximum over the ranks returned.
        # The ranks see different local times (rank 1 is 4 x slower in block 0): the block count
        # must still agree, or the next collective would hang
        calls.append(b)
        local = (10.0 if rank == 0 else 40.0) + b
        t = torch.tensor([local])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    blocks = bench.measure_blocks(timed_block, world, torch.device("cpu"), budget_ms=250.0, most=25)
    results[rank] = (blocks, calls)
    dist.destroy_process_group()


def test_bench_blocks_agree_across_ranks():
    """bench.py times several blocks of K steps and reports the median: every rank must run the
    same number of blocks (each contains barriers and an all-reduce)."""
    port = 31500 + (os.getpid() % 2000)
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(_blocks_worker, args=(2, port, results), nprocs=2, join=True)
    assert len(results) == 2
    (b0, c0), (b1, c1) = results[0], results[1]
    assert b0 == b1 and c0 == c1
    assert len(b0) == 6 and b0[0] == 40.0            # 250 ms / 40 ms -> 6 blocks in all
    # a single process: no collective, the same arithmetic
    import bench
    assert len(bench.measure_blocks(lambda b: 100.0, 1, torch.device("cpu"))) == 2
    assert len(bench.measure_blocks(lambda b: 1000.0, 1, torch.device("cpu"))) == 1
    assert len(bench.measure_blocks(lambda b: 0.5, 1, torch.device("cpu"))) == 25
