#!/bin/bash
# Round 2, call 9: column-blocked normal operator (tests, C5 full size), interleaved warp reductions (tests, trace), V*Q unroll.
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
GA_TEST_SKIP_C5_FULL=1 timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu9.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu9.log
python profiles/microbench/trace_gs.py > $OUT/trace9.log 2>&1; cat $OUT/trace9.log
python profiles/microbench/gs_per_j.py f64 4000 40 > $OUT/perj9_exact.json 2> $OUT/perj9_exact.err
python profiles/microbench/gs_per_j.py f64 1414 40 > $OUT/perj9_2M.json 2> $OUT/perj9_2M.err
python - <<'PY'
import json
for name in ("perj9_exact","perj9_2M"):
    x=json.load(open("gpurun_out/r02/%s.json"%name))["per_j_us_GBps"]
    print(name, [(r[0],r[1],r[2]) for r in x if r[0] in (1,8,11,16,20,24,28,32,33,36,40)], "sum11..40 us", round(sum(r[1] for r in x if r[0]>=11)))
PY
for dt in f64 c64 dd f32; do
  mm=4000; [ $dt = c64 ] && mm=2000; [ $dt = dd ] && mm=2000
  python profiles/microbench/vq_bench.py $dt $mm 40 10 2>&1 | tail -1
done | tee $OUT/vq_bench9.jsonl
timeout 900 python profiles/microbench/c5_normal_bench.py > $OUT/c5_normal.jsonl 2> $OUT/c5_normal.err; echo "c5 rc=$?"; tail -1 $OUT/c5_normal.jsonl; tail -2 $OUT/c5_normal.err
