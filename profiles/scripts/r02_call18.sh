#!/bin/bash
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
for i in 1 2 3; do
  GA_TIMING=1 timeout 300 python bench.py --no-cpu --no-tts --steps 10 --warmup 3 > $OUT/bench18.json 2> $OUT/bench18.err
  grep -E "ga_create|csr_upload|ga_set_csr" $OUT/bench18.err | tail -4
  python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench18.json").read().strip().splitlines()[-1]); print(d["value"], d["e2e"]["value"], d["e2e"]["breakdown_s"])
PY
done
