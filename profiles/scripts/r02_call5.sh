#!/bin/bash
# Round 2, call 5: chunk-staged V*Q kernel (KC sweep), new BASELINE-size tests (C1, C3, C5 full size, dlascl branch, singular svds, F32 lock step).
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
free -g | head -2; nproc
timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/pytest_gpu5.log 2>&1; echo "pytest rc=$?"; tail -22 $OUT/pytest_gpu5.log
for kc in 4 8 10 16; do
  GA_VQ_KC=$kc python profiles/microbench/vq_bench.py f64 4000 40 10 2>&1 | tail -1
done | tee $OUT/vq_bench5.jsonl
for dt in c64 dd f32; do
  mm=4000; [ $dt = c64 ] && mm=2000; [ $dt = dd ] && mm=2000
  python profiles/microbench/vq_bench.py $dt $mm 40 10 2>&1 | tail -1
done | tee -a $OUT/vq_bench5.jsonl
python profiles/microbench/vq_bench.py f64 4000 40 20 2>&1 | tail -1 | tee -a $OUT/vq_bench5.jsonl
python profiles/microbench/vq_bench.py f64 1414 40 10 2>&1 | tail -1 | tee -a $OUT/vq_bench5.jsonl
