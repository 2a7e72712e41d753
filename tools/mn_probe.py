"""Debug aid: try MN-major UMMA descriptor variants of the tf32 GEMM in sub-processes."""
import os
import subprocess
import sys

CHILD = r'''
import math, sys, torch
sys.path.insert(0, ".")
from scvae_b200 import kernels as K
def run(layout, M, N, Kd):
    gen = torch.Generator().manual_seed(1)
    shapes = {0: ((M, Kd), (N, Kd)), 1: ((M, Kd), (Kd, N)), 2: ((Kd, M), (Kd, N))}[layout]
    A = torch.randn(shapes[0], generator=gen); B = torch.randn(shapes[1], generator=gen)
    ref = {0: lambda: A.double() @ B.double().t(), 1: lambda: A.double() @ B.double(), 2: lambda: A.double().t() @ B.double()}[layout]()
    C = torch.zeros(M, N, device="cuda")
    K.gemm(layout, M, N, Kd, A.cuda(), B.cuda(), C, tensor_cores=True, workspace=None)
    torch.cuda.synchronize()
    return (C.cpu().double() - ref).abs().max().item() / math.sqrt(Kd)
for layout in (1, 2):
    for shp in ((128, 128, 32), (128, 128, 8), (256, 384, 96)):
        print("  layout", layout, shp, "err/sqrtK = %.4g" % run(layout, *shp), flush=True)
'''

variants = [(4096, 512), (512, 4096), (4096, 1024), (1024, 4096), (4096, 256), (128, 512)]
for lbo, sbo in variants:
    env = dict(os.environ, SCVAE_TC_MN_LBO=str(lbo), SCVAE_TC_MN_SBO=str(sbo))
    print("LBO", lbo, "SBO", sbo, flush=True)
    try:
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, timeout=120, capture_output=True, text=True)
        print(r.stdout, r.stderr[-500:] if r.returncode else "", flush=True)
    except subprocess.TimeoutExpired:
        print("  TIMEOUT", flush=True)
