"""Time / profile the fused heads kernel alone at the C2 shape (development aid).

    python tools/fused_bench.py [B] [likelihood] [--timeline]

--timeline prints the clock64 timeline of CTA 0 (needs a library built with
-DSCVAE_FUSED_TIMELINE, e.g. NVCC_FLAGS of scvae_b200/_build.py extended by hand).
    ncu --set full --import-source on --clock-control none -k regex:heads_fused_kernel -s 2 -c 1 \
        -o gpurun_out/fused python tools/fused_bench.py
"""
import sys

import torch

sys.path.insert(0, ".")
from scvae_b200 import kernels as K  # noqa: E402
from scvae_b200.engine import VAEEngine  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4096
    lik = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "negative binomial"
    G, L, H = 20000, 50, [100]
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(0)
    rate = torch.rand(G, generator=gen, device=dev) * 0.14
    mask = torch.rand(B, G, generator=gen, device=dev) < rate
    vals = torch.floor(1.0 - torch.log(torch.rand(B, G, generator=gen, device=dev)) * 1.2).clamp_(1, 500)
    x = mask.float() * vals
    eng = VAEEngine(G, L, H, lik, device=dev, tensor_cores=True)
    p = eng._plan(B, 1)
    eng.set_batch_dense(p, x)
    p.row_const.copy_(torch.lgamma(1.0 + x).sum(1))
    p.have_row_const = True
    eng.sample_noise(p, 1, 0)
    for _ in range(3):
        eng.train_step(p, 1, 1, 1e-4)
    torch.cuda.synchronize()
    l = eng.head
    dd = p.d_decH[-1]
    t16 = p.X16 if p.t16_is_x16 else p.T16

    def run():
        K.heads_fused_bwd(eng.kind, p.D16, p.W16, eng.Gh, t16, B, G, p.dA16, dd, l.n_in, p.logp,
                          p.fused_ws, row_const=p.row_const, go=None, go_scalar=-1.0 / B,
                          scale=p.fused_scale)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    s.record()
    for _ in range(n):
        run()
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) / n * 1e3
    if "--timeline" in sys.argv:
        import ctypes
        from scvae_b200 import _lib
        lib = _lib.load()
        buf = torch.zeros(40 * 16, dtype=torch.int64, device=dev)
        lib.scvae_heads_fused_debug(ctypes.c_void_p(buf.data_ptr()))
        run()
        torch.cuda.synchronize()
        lib.scvae_heads_fused_debug(ctypes.c_void_p(0))
        t = buf.cpu().view(40, 16)
        t0 = int(t[0, 9])
        names = ["mma1_ready", "mma2_go", "store_go", "store_read", "epi0_start", "epi0_preAE", "epi0_AE",
                 "epi0_AS", "epi0_end", "W_issue", "epiL_start", "epiL_end"]
        print("tile " + " ".join("%10s" % n for n in names))
        for n in range(2, 14):
            print("%4d " % n + " ".join("%10d" % (int(t[n, i]) - t0) for i in range(12)))
    print("heads_fused_bwd[%s] B=%d: %.1f us (incl. memset + finish), logp mean %.3f, t_is_half=%s" % (
        lik, B, us, p.logp.mean().item(), p.t16_is_x16))


if __name__ == "__main__":
    main()
