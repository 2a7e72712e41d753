"""C4-shaped GMVAE training-step timing (development aid): K = 20 clusters, 20 000 genes, NB,
latent 50, hidden [100]; fused heads on / off.

    python tools/gmvae_bench.py [B]
"""
import sys

import torch

sys.path.insert(0, ".")
from scvae_b200.gmvae_engine import GMVAEEngine  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    G, L, Kc, H = 20000, 50, 20, [100]
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(0)
    rate = torch.rand(G, generator=gen, device=dev) * 0.14
    mask = torch.rand(B, G, generator=gen, device=dev) < rate
    vals = torch.floor(1.0 - torch.log(torch.rand(B, G, generator=gen, device=dev)) * 1.2).clamp_(1, 500)
    x = mask.float() * vals
    for fused in (True, False):
        eng = GMVAEEngine(G, L, Kc, H, "negative binomial", device=dev, tensor_cores=True)
        eng.fused_heads = fused
        p = eng._plan(B, 1)
        eng.set_batch_dense(p, x)
        torch.manual_seed(1)
        p.eps.normal_()
        for _ in range(2):
            eng.train_step(p, 1, 1, 1e-4)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        s.record()
        for _ in range(n):
            eng.train_step(p, 1, 1, 1e-4)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / n
        print("GMVAE K=%d B=%d fused_heads=%s: %.2f ms/step -> %.3f M cells/s, ELBO %.2f, chunk %d" % (
            Kc, B, fused and p.fused_done, ms, B / ms / 1e3, p.bound[0].item(), p.chunk), flush=True)
        del eng, p
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
