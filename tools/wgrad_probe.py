"""Development aid: the (cells x genes)-operand weight-gradient products alone -- stream-K over 157
gene tiles against one lock-step round of 148 tiles (is the MN-major stream DRAM-page bound?)."""
import sys

import torch

sys.path.insert(0, ".")
from scvae_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
B, G, H = 4096, 20000, 104
Gp = 20008
Xs = [(torch.rand(B, Gp, device=dev) < 0.07).half() for _ in range(2)]
dY = torch.randn(B, H, device=dev).half()
ws = torch.empty(64 << 20, dtype=torch.float32, device=dev)


def timed(fn, n=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(n):
        fn(i)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


W16 = torch.randn(H, Gp, device=dev).half()
W16lo = (torch.randn(H, Gp, device=dev) * 1e-3).half()
for pitch, pitch_y in ((Gp, H), (20032, H), (20032, 128), (20096, 128), (20480, 128)):
    Xp = [torch.zeros(B, pitch, device=dev, dtype=torch.float16) for _ in range(2)]
    for a, b in zip(Xp, Xs):
        a[:, :Gp].copy_(b)
    dYp = torch.zeros(B, pitch_y, device=dev, dtype=torch.float16)
    dYp[:, :H].copy_(dY)
    Wp = torch.zeros(H, pitch, device=dev, dtype=torch.float16); Wp[:, :Gp].copy_(W16)
    Wl = torch.zeros(H, pitch, device=dev, dtype=torch.float16); Wl[:, :Gp].copy_(W16lo)
    for N in (Gp, 148 * 128):
        C = torch.zeros(H, pitch, device=dev)
        t = timed(lambda i: K.gemm_f16(K.GEMM_TN, H, N, B, dYp[:, :H], Xp[i & 1][:, :N], C[:, :N], workspace=ws))
        print("pitch %5d / %3d  dW1 (M=104, N=%5d genes, K=4096): %6.1f us  %5.2f TB/s" % (pitch, pitch_y, N, t, B * N * 2 / t / 1e6), flush=True)
    C = torch.zeros(pitch, H, device=dev)
    t = timed(lambda i: K.gemm_f16(K.GEMM_TN, G, H, B, Xp[i & 1][:, :G], dYp[:, :H], C[:G], workspace=ws))
    print("pitch %5d / %3d  head orientation (M=20000 genes, N=104): %6.1f us  %5.2f TB/s" % (pitch, pitch_y, t, B * G * 2 / t / 1e6), flush=True)
    Y = torch.zeros(B, H, device=dev)
    t = timed(lambda i: K.gemm_f16_split(K.GEMM_NT, B, H, Gp, Xp[i & 1][:, :Gp], Wp[:, :Gp], Wl[:, :Gp], 2, Y, workspace=ws))
    print("pitch %5d        forward (split weights, pair): %6.1f us  %5.2f TB/s" % (pitch, t, B * Gp * 2 / t / 1e6), flush=True)
