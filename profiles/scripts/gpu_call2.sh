#!/bin/bash
# GPU tests, bench with the streamed SpMV vs the first SpMV kernel, ncu full capture of k_spmv_stream.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu > gpurun_out/bench_stream.json 2> gpurun_out/bench_stream.err; echo "bench rc=$?"
GA_SPMV_V1=1 timeout 600 python bench.py --no-cpu --no-e2e > gpurun_out/bench_v1.json 2> gpurun_out/bench_v1.err; echo "bench v1 rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 305 -c 200 --csv --log-file gpurun_out/launches2.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmv_stream -s 120 -c 2 -o gpurun_out/prof_spmv $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
cat gpurun_out/bench_stream.json gpurun_out/bench_v1.json
