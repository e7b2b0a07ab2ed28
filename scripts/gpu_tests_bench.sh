#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json
j=json.load(open('gpurun_out/bench.json'))
print(j['value'], j['ms_per_step'], j['e2e'], j['clocks'])
for k in j['kernels']: print(k['entry'], k['launches_per_step'], round(k['ms_per_step'],3))
"
