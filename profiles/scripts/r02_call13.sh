#!/bin/bash
# Round 2, call 13: two consumer sets with the conflict-free row mapping (A/B), tests.
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
for m in 4000 1414; do
  python profiles/microbench/gs_per_j.py f64 $m 40 > $OUT/perj13_${m}_two.json 2> $OUT/perj13_${m}_two.err
  GA_GS_TWO_SETS=0 python profiles/microbench/gs_per_j.py f64 $m 40 > $OUT/perj13_${m}_one.json 2> $OUT/perj13_${m}_one.err
done
python - <<'PY'
import json
for name in ("perj13_4000_two","perj13_4000_one","perj13_1414_two","perj13_1414_one"):
    try:
        x=json.load(open("gpurun_out/r02/%s.json"%name))["per_j_us_GBps"]
        print(name, [(r[0],r[1],r[2]) for r in x if r[0] in (4,8,11,14,16,20,24,25,28,32)], "sum11.. us", round(sum(r[1] for r in x if r[0]>=11)), "sum7..30", round(sum(r[1] for r in x if 7<=r[0]<=30)))
    except Exception as e: print(name, "ERR", e)
PY
GA_TEST_SKIP_C5_FULL=1 timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu13.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu13.log
