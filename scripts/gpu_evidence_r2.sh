#!/bin/bash
# Round-2 evidence run (1 GPU):  gpurun --timeout 1500 -- 'bash scripts/gpu_evidence_r2.sh'
# smoke, the GPU test suite, bench (+ reference arm), ncu launch list of one step, --set full captures of the fused
# aggregate->Linear kernel, the TMA-fed weight gradient and the rho attention kernels.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
python -m pytest tests -m gpu -q -rxXs > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; python scripts/show_bench.py gpurun_out/r2_bench.json 14
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_step.csv \
  python scripts/n_steps.py 1024 3 > gpurun_out/r2_n_steps.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches_step.csv 3 > gpurun_out/r2_launches_step_summary.txt; head -30 gpurun_out/r2_launches_step_summary.txt
ncu --set full --clock-control none --import-source on -k regex:gin_lin_fused -s 2 -c 1 -o gpurun_out/r2_fused \
  python scripts/fused_bench.py 1024 > gpurun_out/r2_ncu_fused.log 2>&1; tail -6 gpurun_out/r2_ncu_fused.log
ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc_tma|attention_fast2" -s 6 -c 3 -o gpurun_out/r2_wgrad_att \
  python scripts/n_steps.py 1024 2 > gpurun_out/r2_ncu_wgrad_att.log 2>&1; tail -2 gpurun_out/r2_ncu_wgrad_att.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
tail -c 900 gpurun_out/r2_bench_reference.json
