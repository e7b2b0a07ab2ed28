#!/bin/bash
# A/B of the gradient exchange on N GPUs: overlapped buckets on a side stream vs one exchange after the backward.
#   gpurun --gpus 2 -- 'bash scripts/gpu_ddp_ab.sh 2'
N=${1:-2}
mkdir -p gpurun_out
for o in 0 1; do
  SB_DDP_OVERLAP=$o python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/ddp_ab_overlap$o.json 2> gpurun_out/ddp_ab_overlap$o.err
  python - "$o" <<'PY'
import json, sys
o = sys.argv[1]
try:
    j = json.loads(open(f"gpurun_out/ddp_ab_overlap{o}.json").read().strip().splitlines()[-1])
    g = j["grad_exchange"]
    print("overlap", o, "value", j["value"], "ms/step", j["ms_per_step"], "exposed", g["exposed_ms_per_step"],
          "without", g["ms_per_step_without_exchange"], "strong ms/step", j["strong"]["ms_per_step"], j["strong"]["value"])
except Exception as e:
    print("overlap", o, "no line:", e)
PY
done
