import sys, torch
sys.path.insert(0, ".")
import torch.nn.functional as F
from signnet_basisnet_b200.functional import BatchNormActFn
for C, M in ((1, 400), (3, 400), (1, 5000), (16, 400)):
    torch.manual_seed(C + M)
    ld = (C + 3) // 4 * 4
    x = torch.zeros(M, ld); x[:, :C] = torch.randn(M, C) * 0.5 + 0.3
    g, b = torch.rand(C) + 0.5, torch.randn(C) * 0.2
    w = torch.zeros(M, ld); w[:, :C] = torch.randn(M, C)
    xr, gr, br = x[:, :C].double().requires_grad_(True), g.double().requires_grad_(True), b.double().requires_grad_(True)
    (F.relu(F.batch_norm(xr, None, None, gr, br, True, 0.1, 1e-5)) * w[:, :C].double()).sum().backward()
    xg, gg, bg = x.cuda().requires_grad_(True), g.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    out = BatchNormActFn.apply(xg, gg, bg, None, None, None, True, True, C)
    (out * w.cuda()).sum().backward()
    print(C, M, "dx", (xg.grad.cpu()[:, :C].double() - xr.grad).abs().max().item(), "dgamma", (gg.grad.cpu().double() - gr.grad).abs().max().item(),
          "dbeta", (bg.grad.cpu().double() - br.grad).abs().max().item(), br.grad.abs().max().item())
