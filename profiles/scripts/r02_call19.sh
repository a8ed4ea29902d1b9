#!/bin/bash
# Round 2, call 19: device block cache (tests, e2e context time over three default-flow runs).
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
GA_TEST_SKIP_C5_FULL=1 timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu19.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu19.log
for i in 1 2 3; do
  GA_TIMING=1 timeout 300 python bench.py --no-cpu --steps 10 --warmup 3 > $OUT/bench19_$i.json 2> $OUT/bench19_$i.err
  grep -E "ga_create|ga_set_csr" $OUT/bench19_$i.err | tail -2
  python - $OUT/bench19_$i.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["breakdown_s"], d["time_to_solution"]["seconds"], d["parity_gate"]["passed"])
PY
done
nvidia-smi --query-gpu=memory.used --format=csv
