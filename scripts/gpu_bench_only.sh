#!/bin/bash
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json
j=json.load(open('gpurun_out/bench.json'))
print(j['value'], j['ms_per_step'], j['e2e'], j['clocks'])
"
