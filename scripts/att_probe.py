import sys, torch
sys.path.insert(0, ".")
from signnet_basisnet_b200.layout import GraphIndex
from signnet_basisnet_b200.synth import synth_batch
from signnet_basisnet_b200.transformer import AttentionFn
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
p = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
d = synth_batch(B, "zinc", seed=0).to("cuda")
gi = GraphIndex(d.edge_index, d.batch, d.num_graphs)
sl = gi.slots_all(128)
q, k, v = (torch.randn(sl.R, 128, device="cuda").requires_grad_(True) for _ in range(3))
for it in range(3):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    o = AttentionFn.apply(q, k, v, sl, 4, 32, p, 1)
    e[1].record()
    o.backward(torch.ones_like(o))
    e[2].record()
    torch.cuda.synchronize()
    print(f"fwd {e[0].elapsed_time(e[1]):.3f} ms  bwd {e[1].elapsed_time(e[2]):.3f} ms (includes zero-fills)")
