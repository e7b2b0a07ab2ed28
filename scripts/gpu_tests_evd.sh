#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_evd.py -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_evd.log
python - <<'PY'
import sys, time, torch
sys.path.insert(0, ".")
from signnet_basisnet_b200.ops import laplacian_evd
from signnet_basisnet_b200.synth import synth_batch
d = synth_batch(1024, "zinc", seed=1000).to("cuda")
for _ in range(3): laplacian_evd(d.edge_index, d.batch, d.num_graphs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): laplacian_evd(d.edge_index, d.batch, d.num_graphs)
e1.record(); torch.cuda.synchronize()
print(f"device EVD of 1024 ZINC-shape graphs: {e0.elapsed_time(e1)/10:.3f} ms per batch (incl. CSR build + layout)")
dc = d.to("cpu")
from signnet_basisnet_b200.synth import sym_laplacian
n = dc.num_nodes_per_graph.tolist()
t0 = time.perf_counter()
off = 0
for nb in n[:256]:
    m = (dc.edge_index[0] >= off) & (dc.edge_index[0] < off + nb)
    torch.linalg.eigh(sym_laplacian(dc.edge_index[:, m] - off, nb))
    off += nb
print(f"CPU torch.linalg.eigh loop (reference's EVDTransform arithmetic), 256 graphs: {(time.perf_counter()-t0)*4*1e3:.1f} ms per 1024 graphs (extrapolated)")
PY
