#!/bin/bash
# Round 2, call 20: CUDA graph of the generalised problem's CG iterations (tests, A/B).
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_general.py -m gpu -x -q > $OUT/pytest_gpu20.log 2>&1; echo "pytest general rc=$?"; tail -3 $OUT/pytest_gpu20.log
python profiles/microbench/general_bench.py 2>&1 | tail -1 | tee $OUT/general_bench.jsonl
GA_CG_GRAPH=0 python profiles/microbench/general_bench.py 2>&1 | tail -1 | tee -a $OUT/general_bench.jsonl
