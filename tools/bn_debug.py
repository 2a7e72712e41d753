import sys, torch
sys.path.insert(0, ".")
from oracle import scvae_oracle as O
from scvae_b200 import kernels as K
for (M, H, groups) in [(4096, 130, 1), (1024, 130, 1), (4096, 100, 1), (640, 40, 1)]:
    torch.manual_seed(1)
    y = (torch.randn(M, H, dtype=torch.float64) * 3 + 5).requires_grad_(True)
    beta = torch.randn(H, dtype=torch.float64, requires_grad=True)
    params = {"s/BATCH_NORM/beta": beta, "s/BATCH_NORM/moving_mean": torch.zeros(H, dtype=torch.float64),
              "s/BATCH_NORM/moving_variance": torch.ones(H, dtype=torch.float64)}
    out = torch.relu(O.batch_norm(y, "s", params, True, [], groups))
    dout = torch.randn(M, H, dtype=torch.float64)
    (out * dout).sum().backward()
    dev = "cuda"
    ldy, ldo = (H + 3) & ~3, (H + 4) & ~3
    yd = torch.zeros(M, ldy, device=dev); yd[:, :H] = y.detach().float()
    outd = torch.zeros(M, ldo, device=dev)
    mm, mv = torch.zeros(H, device=dev), torch.ones(H, device=dev)
    sm, sr = torch.zeros(groups * H, device=dev), torch.zeros(groups * H, device=dev)
    scratch = torch.zeros(K.bn_scratch_floats(M, H, groups), device=dev)
    K.bn_act_fwd(yd, H, beta.detach().float().to(dev), mm, mv, outd, sm, sr, scratch, training=True)
    doutd = torch.zeros(M, ldo, device=dev); doutd[:, :H] = dout.float()
    dy = torch.zeros(M, ldy, device=dev); dbeta = torch.zeros(H, device=dev)
    K.bn_act_bwd(doutd, yd, outd, H, sm, sr, dy, dbeta, scratch)
    torch.cuda.synchronize()
    err = (dy[:, :H].cpu().double() - y.grad).abs()
    print(M, H, "max err", err.max().item(), "bad cols", (err.max(0).values > 1e-3).nonzero().flatten().tolist()[:20],
          "bad rows", (err.max(1).values > 1e-3).sum().item(), "out err", (outd[:, :H].cpu().double() - out.detach()).abs().max().item(),
          "dbeta err", (dbeta.cpu().double() - beta.grad).abs().max().item())
    mean_ref = y.detach().mean(0); print("  save_mean err", (sm.cpu().double() - mean_ref).abs().max().item())
