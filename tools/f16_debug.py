import math, os, sys, torch
sys.path.insert(0, ".")
from scvae_b200 import kernels as K
dev = "cuda"
for Kd in (64, 192, 320, 384, 512, 1024, 2001):
    for M, N in ((128, 128), (100, 104)):
        g = torch.Generator().manual_seed(1)
        A = torch.randn(M, Kd, generator=g); B = torch.randn(N, Kd, generator=g)
        ld = (Kd + 7) & ~7
        Ad = torch.zeros(M, ld, dtype=torch.float16, device=dev); Ad[:, :Kd] = A.half()
        Bd = torch.zeros(N, ld, dtype=torch.float16, device=dev); Bd[:, :Kd] = B.half()
        ref = Ad[:, :Kd].cpu().double() @ Bd[:, :Kd].cpu().double().t()
        C = torch.full((M, (N + 3) & ~3), 3.0, device=dev)
        wsb = K.gemm_f16_workspace_bytes(0, M, N, Kd)
        ws = torch.empty(max(wsb // 4, 1), device=dev)
        K.gemm_f16(0, M, N, Kd, Ad, Bd, C, workspace=ws)
        torch.cuda.synchronize()
        got = C[:, :N].cpu().double()
        print("K=%5d M=%d N=%d ws=%d err=%.3g  C[0,:3]=%s ref[0,:3]=%s untouched=%d" % (
            Kd, M, N, wsb, (got - ref).abs().max().item(), got[0, :3].tolist(), ref[0, :3].tolist(),
            int((got == 3.0).sum())), flush=True)
