"""Ad-hoc per-kernel timing at a BASELINE config (development aid, not the bench contract)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from scvae_b200 import kernels as K  # noqa: E402
from scvae_b200.engine import VAEEngine  # noqa: E402


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3  # us


def main():
    G, L, H, lik, B = 20000, 50, [100], "negative binomial", 4096
    if len(sys.argv) > 1:
        B = int(sys.argv[1])
    dev = torch.device("cuda:0")
    x = (torch.rand(B, G, device=dev) < 0.07).float() * torch.randint(1, 6, (B, G), device=dev).float()
    for tc in (True, False):
        eng = VAEEngine(G, L, H, lik, device=dev, tensor_cores=tc)
        p = eng._plan(B, 1)
        eng.set_batch_dense(p, x)
        eng.sample_noise(p, 1, 0)
        t = timed(lambda: eng.train_step(p, 1, 1, 1e-4), n=5, warm=2)
        print("tensor_cores=%s  train_step %.1f us  -> %.3f M cells/s  bound=%s" % (
            tc, t, B / t, p.bound.cpu().tolist()), flush=True)
        if not tc:
            continue
        eng._plan_backward(p)
        M = B
        head = eng.head
        rows = [
            ("likelihood_bwd", lambda: K.likelihood_bwd(eng.kind, p.X, p.A, eng.Gn, M, G, p.dA, logp=p.logp, go=None, go_scalar=-1.0 / B),
             (eng.P * 2 + 1) * 4 * B * G),
            ("likelihood_fwd", lambda: K.likelihood_fwd(eng.kind, p.X, p.A, eng.Gn, M, G, p.logp), (eng.P + 1) * 4 * B * G),
            ("enc1 fwd NT", lambda: eng._gemm(p, K.GEMM_NT, B, 100, G + 1, p.X, eng.enc[0].w, p.encY[0]), 4 * B * G),
            ("heads fwd NT", lambda: eng._gemm(p, K.GEMM_NT, M, head.n_out, head.n_in + 1, p.decH[-1], head.w, p.A), 4 * B * eng.P * G),
            ("heads wgrad TN", lambda: eng._gemm(p, K.GEMM_TN, head.n_out, head.in_p, M, p.dA, p.decH[-1], head.dw), 4 * B * eng.P * G),
            ("heads dgrad NN", lambda: eng._gemm(p, K.GEMM_NN, M, head.n_in, head.n_out, p.dA, head.w, p.d_decH[-1]), 4 * B * eng.P * G),
            ("enc1 wgrad TN", lambda: eng._gemm(p, K.GEMM_TN, 100, eng.enc[0].in_p, B, p.d_encY[0], p.X, eng.enc[0].dw), 4 * B * G),
            ("adam", lambda: eng.optimiser_step(1e-4), 7 * 4 * eng.store.total),
        ]
        for name, fn, nbytes in rows:
            t = timed(fn)
            print("   %-18s %8.1f us   %7.1f GB/s (algorithmic)" % (name, t, nbytes / t / 1e3), flush=True)


if __name__ == "__main__":
    main()
