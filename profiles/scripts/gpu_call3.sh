#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
: > gpurun_out/spmv_bench.jsonl
for mat in lap c4 c5 c1; do
  GA_SPMV_V1=1 timeout 300 python profiles/microbench/spmv_bench.py $mat >> gpurun_out/spmv_bench.jsonl 2>> gpurun_out/spmv_bench.err
  for st in 4 5 6 7 8; do
    GA_SPMV_STAGES=$st timeout 300 python profiles/microbench/spmv_bench.py $mat >> gpurun_out/spmv_bench.jsonl 2>> gpurun_out/spmv_bench.err
  done
done
cat gpurun_out/spmv_bench.jsonl
timeout 600 python bench.py --no-cpu --no-e2e > gpurun_out/bench_stream2.json 2> gpurun_out/bench_stream2.err; echo "bench rc=$?"
cat gpurun_out/bench_stream2.json | cut -c1-400
