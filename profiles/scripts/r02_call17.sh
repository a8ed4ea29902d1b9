#!/bin/bash
# Round 2, call 17: e2e context time in the full default flow (gate + tts on), upload threads 1 vs 8, 10 vs 20 cycles.
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
for cfg in "8 10" "1 10" "8 20" "8 10 --no-gate" "8 10 --no-tts"; do
  set -- $cfg
  GA_UPLOAD_THREADS=$1 timeout 300 python bench.py --no-cpu --steps $2 --warmup 3 ${3:-} > $OUT/bench17.json 2> $OUT/bench17.err
  python - "$cfg" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/r02/bench17.json").read().strip().splitlines()[-1]); print(sys.argv[1], "|", d["value"], d["e2e"]["value"], d["e2e"]["breakdown_s"])
PY
done
