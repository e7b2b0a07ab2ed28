#!/bin/bash
# Round-2 evidence run (1 GPU):  gpurun --timeout 1500 -- 'bash scripts/gpu_evidence_r2.sh'
# smoke, the GPU test suite, bench (+ reference arm), ncu launch list of one step, --set full captures of the fused
# aggregate->Linear kernel, the TMA-fed weight gradient and the rho attention kernels (mma.sync path), the micro-benchmarks
# quoted in DESIGN.md.  Everything lands in gpurun_out/ with the prefix r2f_ ("final").
mkdir -p gpurun_out
P=gpurun_out/r2f
python __graft_entry__.py smoke > ${P}_smoke.log 2>&1; tail -1 ${P}_smoke.log
python -m pytest tests -m gpu -q -rxXs > ${P}_pytest_gpu.log 2>&1; tail -3 ${P}_pytest_gpu.log
# three windows of the same build: the pool's hosts stall the issuing thread for 60-160 ms in some windows
for i in 1 2 3; do
  SB_BENCH_DEBUG=1 python bench.py --steps 20 --warmup 5 > ${P}_bench_run$i.json 2> ${P}_bench_run$i.err
  python scripts/show_bench.py ${P}_bench_run$i.json 0 | head -1
done
cp ${P}_bench_run1.json ${P}_bench.json; python scripts/show_bench.py ${P}_bench.json 14
python scripts/configs_probe.py > ${P}_configs.log 2>&1; tail -6 ${P}_configs.log
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
