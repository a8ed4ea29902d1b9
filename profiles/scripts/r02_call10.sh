#!/bin/bash
# Round 2, call 10: transposed warp reduction (tests, timeline, per-j at 2M rows), C5 column-block width sweep.
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
GA_TEST_SKIP_C5_FULL=1 timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu10.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu10.log
python profiles/microbench/trace_gs.py > $OUT/trace10.log 2>&1; cat $OUT/trace10.log
python profiles/microbench/gs_per_j.py f64 4000 40 > $OUT/perj10_exact.json 2> $OUT/perj10_exact.err
python profiles/microbench/gs_per_j.py f64 1414 40 > $OUT/perj10_2M.json 2> $OUT/perj10_2M.err
python - <<'PY'
import json
for name in ("perj10_exact","perj10_2M"):
    x=json.load(open("gpurun_out/r02/%s.json"%name))["per_j_us_GBps"]
    print(name, [(r[0],r[1],r[2]) for r in x if r[0] in (1,8,11,16,20,24,28,32,33,36,40)], "sum11..40 us", round(sum(r[1] for r in x if r[0]>=11)))
PY
timeout 900 python profiles/microbench/c5_normal_bench.py 1.0 sweep > $OUT/c5_normal_sweep.jsonl 2> $OUT/c5_normal_sweep.err; echo "c5 rc=$?"; tail -1 $OUT/c5_normal_sweep.jsonl; tail -2 $OUT/c5_normal_sweep.err
