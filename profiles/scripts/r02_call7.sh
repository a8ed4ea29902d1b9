#!/bin/bash
# Round 2, call 7: V*Q kernel with 2 rows x 4*NOB outputs per thread (tests + timing).
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
GA_TEST_SKIP_C5_FULL=1 timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu7.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu7.log
for dt in f64 c64 dd f32; do
  mm=4000; [ $dt = c64 ] && mm=2000; [ $dt = dd ] && mm=2000
  python profiles/microbench/vq_bench.py $dt $mm 40 10 2>&1 | tail -1
done | tee $OUT/vq_bench7.jsonl
python profiles/microbench/vq_bench.py f64 4000 40 20 2>&1 | tail -1 | tee -a $OUT/vq_bench7.jsonl
python profiles/microbench/vq_bench.py f64 4000 40 5 2>&1 | tail -1 | tee -a $OUT/vq_bench7.jsonl
python profiles/microbench/vq_bench.py f64 4000 40 14 2>&1 | tail -1 | tee -a $OUT/vq_bench7.jsonl
python profiles/microbench/vq_bench.py f64 1414 40 10 2>&1 | tail -1 | tee -a $OUT/vq_bench7.jsonl
