#!/bin/bash
# Round-2 evidence run (1 GPU):  gpurun --timeout 1500 -- 'bash scripts/gpu_evidence_r2.sh'
# smoke, the GPU test suite, bench (+ reference arm), ncu launch list of one step, --set full captures of the fused
# aggregate->Linear kernel, the TMA-fed weight gradient and the rho attention kernels (mma.sync path), the micro-benchmarks
# quoted in DESIGN.md.  Everything lands in gpurun_out/ with the prefix r2f_ ("final").
mkdir -p gpurun_out
P=gpurun_out/r2f
python __graft_entry__.py smoke > ${P}_smoke.log 2>&1; tail -1 ${P}_smoke.log
python -m pytest tests -m gpu -q -rxXs > ${P}_pytest_gpu.log 2>&1; tail -3 ${P}_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > ${P}_bench.json 2> ${P}_bench.err; python scripts/show_bench.py ${P}_bench.json 14
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${P}_launches_step.csv \
  python scripts/n_steps.py 1024 3 > ${P}_n_steps.log 2>&1
python scripts/launch_summary.py ${P}_launches_step.csv 3 > ${P}_launches_step_summary.txt; head -24 ${P}_launches_step_summary.txt
python scripts/fused_bench.py 1024 > ${P}_fused_bench.log 2>&1; tail -5 ${P}_fused_bench.log
python scripts/attention_bench.py 1024 > ${P}_attention_bench.log 2>&1; tail -4 ${P}_attention_bench.log
python scripts/wgrad_only.py > ${P}_wgrad_lin.log 2>&1; python scripts/lin_only.py >> ${P}_wgrad_lin.log 2>&1; cat ${P}_wgrad_lin.log
ncu --set full --clock-control none --import-source on -k regex:gin_lin_fused -s 2 -c 1 -o ${P}_fused \
  python scripts/fused_bench.py 1024 > ${P}_ncu_fused.log 2>&1; tail -2 ${P}_ncu_fused.log
SB_SKIP_FFMA=1 ncu --set full --clock-control none --import-source on -k regex:"attention_mma" -s 4 -c 2 -o ${P}_attention \
  python scripts/attention_bench.py 1024 > ${P}_ncu_attention.log 2>&1; tail -2 ${P}_ncu_attention.log
ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc_tma" -s 4 -c 1 -o ${P}_wgrad \
  python scripts/wgrad_only.py > ${P}_ncu_wgrad.log 2>&1; tail -2 ${P}_ncu_wgrad.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > ${P}_bench_reference.json 2> ${P}_bench_reference.err
tail -c 600 ${P}_bench_reference.json
