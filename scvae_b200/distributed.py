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


def peer_slice_bounds(lo, hi, rank_index, world):
    """[a, b) of the flat range [lo, hi) whose reduction, Adam update and parameter broadcast
    rank ``rank_index`` owns in ``scvae_dp_reduce_adam`` (128-bit granules, ceil-divided)."""
    n4 = (hi - lo) // 4
    per = (n4 + world - 1) // world
    a = min(rank_index * per, n4)
    return lo + 4 * a, lo + 4 * min(a + per, n4)


class PeerExchange:
    """Gradient exchange over peer memory, fused with the optimiser (``scvae_dp_reduce_adam``).

    The flat parameter and gradient buffers of the engine are moved into symmetric memory
    (``torch.distributed._symmetric_memory``: every rank maps every peer's buffer over NVLink);
    one kernel per rank and exchange channel then sums its 1/W slice of the gradient from all
    peers, applies clip + Adam to that slice and stores the new parameters into every replica.
    Only this rank's slice of the Adam slots is live; ``gather_slots`` rebuilds the full slots
    (checkpoints).  NCCL remains the transport for set-up collectives and for the plain
    all-reduce mode (``set_data_parallel(..., all_reduce)``)."""

    CHANNELS = 2           # main stream / side stream of the step engine
    FLAGS_PER_CHANNEL = 32

    def __init__(self, engine, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        store, dev = engine.store, engine.device
        n = store.total
        self.param = symm_mem.empty(n, dtype=torch.float32, device=dev)
        self.grad = symm_mem.empty(n, dtype=torch.float32, device=dev)
        self.flags = symm_mem.empty(self.CHANNELS * self.FLAGS_PER_CHANNEL, dtype=torch.int32,
                                    device=dev)
        self.param.copy_(store.param)
        self.grad.zero_()
        self.flags.zero_()
        self.h_param = symm_mem.rendezvous(self.param, group)
        self.h_grad = symm_mem.rendezvous(self.grad, group)
        self.h_flags = symm_mem.rendezvous(self.flags, group)
        self.param_ptrs = [int(x) for x in self.h_param.buffer_ptrs]
        self.grad_ptrs = [int(x) for x in self.h_grad.buffer_ptrs]
        self.flag_ptrs = [int(x) for x in self.h_flags.buffer_ptrs]
        self.ctl = torch.zeros(self.CHANNELS * 4, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)
        dist.barrier(group)                    # every rank's buffers are initialised
        engine.rebind_flat_buffers(self.param, self.grad)
        self.engine = engine

    def reduce_adam(self, lo, hi, learning_rate, channel, max_ctas=0, scalars=None):
        """Exchange + optimiser update of the flat range [lo, hi) on the current stream."""
        from . import kernels as K
        from .engine import ADAM_BETA1, ADAM_BETA2, ADAM_EPSILON, GRADIENT_CLIP
        s = self.engine.store
        fo = channel * self.FLAGS_PER_CHANNEL * 4
        K.dp_reduce_adam(self.world, self.rank, [p + 4 * lo for p in self.grad_ptrs],
                         [p + 4 * lo for p in self.param_ptrs], [p + fo for p in self.flag_ptrs],
                         s.m[lo:hi], s.v[lo:hi], hi - lo, s.step, learning_rate, ADAM_BETA1,
                         ADAM_BETA2, ADAM_EPSILON, GRADIENT_CLIP, 1.0 / self.world,
                         ctl=self.ctl[channel * 4:channel * 4 + 4], max_ctas=max_ctas,
                         scalars=scalars)

    def slice_bounds(self, lo, hi, rank_index):
        """[a, b) of the flat range [lo, hi) owned by ``rank_index`` (as in the kernel)."""
        return peer_slice_bounds(lo, hi, rank_index, self.world)

    def gather_slots(self, ranges):
        """Full Adam slots on every rank (each rank owns one slice per exchanged range)."""
        s = self.engine.store
        for buf in (s.m, s.v):
            full = torch.zeros_like(buf)
            for lo, hi in ranges:
                a, b = self.slice_bounds(lo, hi, self.rank)
                full[a:b] = buf[a:b]
            dist.all_reduce(full)
            buf.copy_(full)

    def timed_out(self):
        return bool(self.ctl.view(self.CHANNELS, 4)[:, 2].any().item())


def attach(engine, exchange=None):
    """Switch an engine to data-parallel mode if a process group is active.  ``exchange``:
    "p2p" (fused peer-memory exchange + optimiser), "nccl" (flat all-reduce, then the replicated
    optimiser) or None = p2p when symmetric memory is available on a CUDA engine, else nccl."""
    if not is_active():
        return engine
    exchange = exchange or os.environ.get("SCVAE_DP_EXCHANGE")
    engine.set_data_parallel(world_size(), all_reduce_sum_)
    want_p2p = exchange in (None, "p2p") and hasattr(engine, "set_peer_exchange") \
        and torch.device(engine.device).type == "cuda" and dist.get_backend() == "nccl"
    if want_p2p:
        try:
            engine.set_peer_exchange(PeerExchange(engine))
        except Exception as exc:          # symmetric memory unavailable: NCCL all-reduce mode
            if exchange == "p2p":
                raise
            if rank() == 0:
                import sys
                print("scvae_b200: peer-memory exchange unavailable ({}); using NCCL "
                      "all-reduce".format(exc), file=sys.stderr)
    return engine


def raise_if_exchange_failed(engine):
    """Checked once per epoch, before the checkpoint: ``scvae_dp_reduce_adam`` bounds its flag
    waits and records an expired one in its control block instead of hanging the GPU."""
    peer = getattr(engine, "_peer", None)
    if peer is None or not is_active():
        return
    failed = torch.tensor([1 if peer.timed_out() else 0], dtype=torch.int32, device=engine.device)
    dist.all_reduce(failed, op=dist.ReduceOp.MAX)      # every rank raises, none is left waiting
    if int(failed.item()):
        raise RuntimeError("scvae_b200: the peer-memory gradient exchange timed out on at least "
                           "one rank during this epoch (a rank stalled beyond the bounded flag "
                           "wait); replicas may have diverged -- aborting before the checkpoint")


def broadcast_parameters(engine):
    """Make replicas bit-identical at start (rank 0's parameters, optimiser slots and step)."""
    if not is_active():
        return
    s = engine.store
    if hasattr(engine, "invalidate_shadows"):
        engine.invalidate_shadows()
    for buf in (s.param, s.m, s.v, s.step):
        dist.broadcast(buf, src=0)
    for layer in getattr(engine, "bn_layers", lambda: [])():
        dist.broadcast(layer.moving_mean, src=0)
        dist.broadcast(layer.moving_var, src=0)
