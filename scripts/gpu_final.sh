#!/bin/bash
# Round-end evidence: smoke, GPU tests, bench (+ reference arm), ncu launch list of the bench command, full captures.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu-launch rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gin_agg_tma -s 6 -c 2 -f -o gpurun_out/prof_gin_agg_fwd \
    python scripts/agg_bench.py 1024 128 4 > gpurun_out/ncu_full_fwd.log 2>&1; echo "ncu-agg rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_tc_kernel -s 12 -c 1 -f -o gpurun_out/prof_linear_tc \
    python scripts/perf_probe.py 1024 128 8 > gpurun_out/ncu_lin.log 2>&1; echo "ncu-lin rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 4 -c 1 -f -o gpurun_out/prof_wgrad_tc \
    python scripts/perf_probe.py 1024 128 8 > gpurun_out/ncu_wg.log 2>&1; echo "ncu-wg rc=$?"
python -c "
import json
j=json.load(open('gpurun_out/bench.json'))
print(j['value'], j['ms_per_step'], j['e2e'], j['roofline']['frac'], j['clocks'])
"
