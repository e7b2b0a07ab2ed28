import sys, torch
sys.path.insert(0, ".")
from signnet_basisnet_b200.functional import linear_fwd
R, S = 575454, 2
X = torch.randn(S, R, 128, device="cuda"); H = torch.empty_like(X)
W = (torch.randn(128, 128) / 11.3).cuda()
st = torch.zeros(S, 2, 128, dtype=torch.float64, device="cuda")
for _ in range(4):
    linear_fwd(X, 128, W, 128, 1, None, H, 128, R, S, 128, 128, stats=st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    linear_fwd(X, 128, W, 128, 1, None, H, 128, R, S, 128, 128, stats=st)
e1.record(); torch.cuda.synchronize()
print("us per launch", e0.elapsed_time(e1) * 100)
