"""sb_linear_fwd at the rho size of cfg 4 (575 454 rows, K = N = 128): plain and accumulating (y += x W^T) launches."""
import sys, torch
sys.path.insert(0, ".")
from signnet_basisnet_b200.functional import linear_fwd
R = 575454
X = torch.randn(R, 128, device="cuda"); Y = torch.zeros(R, 128, device="cuda")
W = (torch.randn(128, 128) / 11.3).cuda()
for acc in (False, True):
    for _ in range(3):
        linear_fwd(X, 128, W, 128, 1, None, Y, 128, R, 1, 128, 128, accumulate=acc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        linear_fwd(X, 128, W, 128, 1, None, Y, 128, R, 1, 128, 128, accumulate=acc)
    e1.record(); torch.cuda.synchronize()
    print("accumulate" if acc else "plain     ", f"{e0.elapsed_time(e1) * 100:.1f} us per launch")
a, b = torch.randn(R, 128, device="cuda"), torch.randn(R, 128, device="cuda")
torch.cuda.synchronize(); e0.record()
for _ in range(10):
    c = a + b
e1.record(); torch.cuda.synchronize()
print("at::add    ", f"{e0.elapsed_time(e1) * 100:.1f} us per launch")
