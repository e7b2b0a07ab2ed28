"""rho attention at the cfg 4 shape (1024 ZINC-shape graphs, 4 heads x 32, k_b = n_b <= 37): CUDA-event time of
sb_attention_fwd / sb_attention_bwd on the tensor-core kernels (attention_mma.cu) and on the FFMA kernels they replace."""
import sys

import torch

sys.path.insert(0, ".")
from signnet_basisnet_b200 import _lib
from signnet_basisnet_b200.layout import GraphIndex
from signnet_basisnet_b200.synth import synth_batch
from signnet_basisnet_b200.transformer import AttentionFn

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
d = synth_batch(B, "zinc", seed=1000)
gi = GraphIndex(d.edge_index.cuda(), d.batch.cuda(), d.num_graphs)
sl = gi.slots_all(128)
q, k, v, w = (torch.randn(sl.R, 128, device="cuda") for _ in range(4))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
print(f"B={B} N={gi.N} R={sl.R} kmax={sl.kmax}; algorithmic bytes fwd {4 * sl.R * 512 / 1e9:.2f} GB, bwd {7 * sl.R * 512 / 1e9:.2f} GB")
import os
for mma in ((1,) if os.environ.get('SB_SKIP_FFMA') else (1, 0)):
    _lib.lib().sb_set_attention_mma(mma)
    for p in ((0.0,) if os.environ.get('SB_SKIP_FFMA') else (0.0, 0.1)):
        tf = tb = 0.0
        for it in range(8):
            qq, kk, vv = (t.clone().requires_grad_(True) for t in (q, k, v))
            flush.zero_()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            e[0].record()
            o = AttentionFn.apply(qq, kk, vv, sl, 4, 32, p, 99)
            e[1].record()
            flush.zero_()
            e[2].record()
            o.backward(w)
            e[3].record()
            torch.cuda.synchronize()
            if it >= 3:
                tf += e[0].elapsed_time(e[1]) / 5
                tb += e[2].elapsed_time(e[3]) / 5
        print(f"{'mma ' if mma else 'ffma'} drop_p={p}: fwd {1e3 * tf:7.1f} us ({4 * sl.R * 512 / tf / 1e6:6.0f} GB/s)   "
              f"bwd {1e3 * tb:7.1f} us ({7 * sl.R * 512 / tb / 1e6:6.0f} GB/s)")
_lib.lib().sb_set_attention_mma(1)
