#!/bin/bash
# One gpurun call: GPU tests, bench (step kernel on/off), ncu launch list + one full capture of the top kernel.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_step.json 2> gpurun_out/bench_step.err; echo "bench rc=$?"
GA_NO_STEP_KERNEL=1 timeout 600 python bench.py --no-cpu > gpurun_out/bench_nostep.json 2> gpurun_out/bench_nostep.err; echo "bench nostep rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 305 -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gs_step -s 120 -c 2 -o gpurun_out/prof_gs_step $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
cat gpurun_out/bench_step.json gpurun_out/bench_nostep.json
