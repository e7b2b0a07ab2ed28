"""sb_linear_wgrad at the phi size of cfg 4 (2 x 575 454 rows, N = K = 128): CUDA-event time per launch."""
import sys, torch
sys.path.insert(0, ".")
from signnet_basisnet_b200.functional import linear_wgrad
R, S = 575454, 2
X = torch.randn(S, R, 128, device="cuda"); G = torch.randn(S, R, 128, device="cuda")
dW = torch.empty(128, 128, device="cuda")
for _ in range(4):
    linear_wgrad(G, 128, X, 128, R, S, 128, 128, dW, 128, 1, None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    linear_wgrad(G, 128, X, 128, R, S, 128, 128, dW, 128, 1, None)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) * 100
print(f"wgrad us per launch {t:.1f}  ({2 * 4 * 128 * S * R / t / 1e3:.0f} GB/s)")
