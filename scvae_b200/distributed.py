"""Data-parallel plumbing: one process per GPU, cells sharded across ranks, ONE
flat fp32 sum all-reduce of the gradient buffer per optimiser step (NCCL over NVLink), then the
identical clip + Adam on every rank.  The reference has no distributed code at all; this is
the only exchange step the path needs, so no other collective exists on the data path.

Conventions
  * every rank sees the same global permutation (rank 0 broadcasts it once per epoch) and
    takes the ``rank``-th contiguous slice of each global minibatch, so a W-rank run visits
    the same global minibatches as a 1-rank run;
  * each rank computes the gradient of the mean over ITS slice; the all-reduce sums them and
    the optimiser kernel scales by 1/W (equal slices), i.e. the global-minibatch mean -- the
    elementwise clip is applied after the reduction, as in the reference (VAE:2751-2755);
  * batch-norm statistics are per rank (local batch statistics; rank 0's moving averages
    are the ones checkpointed).
"""

import os

import numpy
import torch
import torch.distributed as dist


def is_active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def initialise_from_environment(backend=None):
    """torchrun-style initialisation (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1 or (dist.is_available() and dist.is_initialized()):
        return rank(), world_size()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend)
    return dist.get_rank(), dist.get_world_size()


def shard_bounds(global_start, global_rows, rank_index, world):
    """[lo, hi) of this rank's slice of a global minibatch of ``global_rows`` rows; slices are
    equal (``global_rows // world``), a remainder is dropped so that every rank does the same
    amount of work (the gradient average stays exact)."""
    per_rank = global_rows // world
    lo = global_start + rank_index * per_rank
    return lo, lo + per_rank


def broadcast_permutation(n, random_state, device="cpu"):
    """Same shuffled order on every rank: rank 0 draws it, the others receive it."""
    if rank() == 0:
        perm = torch.from_numpy(numpy.asarray(random_state.permutation(n), dtype=numpy.int64))
    else:
        perm = torch.empty(n, dtype=torch.int64)
    if is_active():
        perm = perm.to(device)
        dist.broadcast(perm, src=0)
    return perm


def all_reduce_sum_(flat):
    """In-place sum of the flat gradient buffer over ranks (the single per-step collective)."""
    if is_active():
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def attach(engine):
    """Switch an engine to data-parallel mode if a process group is active."""
    if is_active():
        engine.set_data_parallel(world_size(), all_reduce_sum_)
    return engine


def broadcast_parameters(engine):
    """Make replicas bit-identical at start (rank 0's parameters, optimiser slots and step)."""
    if not is_active():
        return
    s = engine.store
    for buf in (s.param, s.m, s.v, s.step):
        dist.broadcast(buf, src=0)
    for layer in getattr(engine, "bn_layers", lambda: [])():
        dist.broadcast(layer.moving_mean, src=0)
        dist.broadcast(layer.moving_var, src=0)
