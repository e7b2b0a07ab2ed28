#!/bin/bash
# Does Python's cyclic GC cause the single-step hiccups of bench.py?  A/B: SB_BENCH_GC=1 (collector left on) vs default.
for g in 1 0; do for i in 1 2 3; do
  SB_BENCH_GC=$g SB_BENCH_DEBUG=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/dbg.log > gpurun_out/gc_ab.json
  python - "$g" <<'PY'
import json, sys
j = json.loads(open("gpurun_out/gc_ab.json").read().strip().splitlines()[-1])
rows = [[float(x) for x in l.split(":")[1].split()] for l in open("gpurun_out/dbg.log") if l.startswith("cpu ms")]
print("gc_on" if sys.argv[1] == "1" else "gc_off", "value ms/step", j["ms_per_step"], "e2e ms/step", j["e2e"]["ms_per_step"],
      "| max single step (value, e2e, ...):", [max(r) for r in rows])
PY
done; done
