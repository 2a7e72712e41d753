"""Data-parallel exchange check (run under torchrun, N >= 2 GPUs):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py

Trains the same small VAE for a few steps with the NCCL all-reduce + replicated Adam and with the
fused peer-memory exchange (scvae_dp_reduce_adam); parameters and Adam slots must agree, and all
replicas must be bit-identical."""
import os
import sys

import numpy
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scvae_b200 import distributed as D  # noqa: E402
from scvae_b200.engine import VAEEngine  # noqa: E402
from scvae_b200.hotloop import ResidentCSR, TrainLoop  # noqa: E402


def main():
    rank, world = D.initialise_from_environment("nccl")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    import scipy.sparse
    G, B, N = 2048, 256, 1024
    rng = numpy.random.RandomState(7 + rank)
    x = (rng.rand(N, G) < 0.08) * rng.randint(1, 9, size=(N, G))
    csr = scipy.sparse.csr_matrix(x.astype(numpy.float32))
    results = {}
    for mode in ("nccl", "p2p", "nccl#2", "p2p#2"):
        eng = VAEEngine(G, 16, [64], "negative binomial", device=dev, seed=0)
        D.attach(eng, exchange=mode.split("#")[0])
        assert (eng._peer is not None) == mode.startswith("p2p")
        D.broadcast_parameters(eng)
        loop = TrainLoop(eng, B, seed=1 + rank, use_graph=True)
        data = ResidentCSR(csr, dev)
        for i in range(6):
            loop.rows.copy_(torch.arange(B, device=dev) + (i % (N // B)) * B)
            bound = loop.step(data, 1e-3, 1.0)
            if i == 0:
                torch.cuda.synchronize()
                sd = eng.state_dict()
                results[mode] = {k: sd[k].clone() for k in ("param", "m", "v")}
        torch.cuda.synchronize()
        results[mode]["bound"] = float(bound[0].item())
        results[mode]["param6"] = eng.state_dict()["param"].clone()
        if eng._peer is not None:
            assert not eng._peer.timed_out(), "peer exchange timed out"
        # replicas identical?
        for k in ("param6",):
            t = results[mode][k].to(dev)
            ref = t.clone()
            dist.broadcast(ref, src=0)
            assert torch.equal(t, ref), "replicas differ in mode {} ({})".format(mode, k)
        loop._graphs = {}
        loop._graph = None
        del loop
    if rank == 0:
        for x, y in (("nccl", "nccl#2"), ("p2p", "p2p#2")):
            print("repeat", x, "max diff after 1 step:",
                  (results[x]["param"] - results[y]["param"]).abs().max().item(), "m:",
                  (results[x]["m"] - results[y]["m"]).abs().max().item())
        a, b = results["nccl"]["param"], results["p2p"]["param"]
        d = (a - b).abs()
        idx = torch.topk(d, 8).indices
        print("n =", a.numel(), "head offset", eng.store.offsets[eng.head.name + "/W"][0],
              "entries with |diff| > 1e-5:", int((d > 1e-5).sum()))
        for i in idx.tolist():
            print("  idx", i, "param", a[i].item(), b[i].item(), "m", results["nccl"]["m"][i].item(),
                  results["p2p"]["m"][i].item(), "v", results["nccl"]["v"][i].item(),
                  results["p2p"]["v"][i].item())
    for k in ("param", "m", "v"):
        a, b = results["nccl"][k], results["p2p"][k]
        err = (a - b).abs().max().item()
        scale = a.abs().max().item()
        if rank == 0:
            print("dp_check {}: max |nccl - p2p| = {:.3e} (scale {:.3e})".format(k, err, scale))
        # the step is bit-reproducible and a two-rank sum is commutative: the two exchange modes
        # agree exactly (observed: 0.0); a few ulp of slack for other NCCL reduction orders
        assert err <= 1e-6 * scale + 1e-12, (k, err, scale)
    # after 6 steps
    a, b = results["nccl"]["param6"], results["p2p"]["param6"]
    frac = ((a - b).abs() > 1e-4).float().mean().item()
    rel = abs(results["nccl"]["bound"] - results["p2p"]["bound"]) / abs(results["nccl"]["bound"])
    if rank == 0:
        print("dp_check 6 steps: {:.4%} of parameters differ by > 1e-4; ELBO {:.4f} vs {:.4f} "
              "(rel {:.2e})".format(frac, results["nccl"]["bound"], results["p2p"]["bound"], rel))
    assert frac < 0.01 and rel < 1e-4
    if rank == 0:
        print("dp_check OK: world", world)
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)


if __name__ == "__main__":
    main()
