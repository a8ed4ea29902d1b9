#!/bin/bash
# Round 2, call 16: where does the e2e leg's context time go (upload threads A/B inside bench.py)?
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
for thr in 1 8; do
  GA_UPLOAD_THREADS=$thr timeout 300 python bench.py --no-cpu --no-tts --no-gate --steps 10 --warmup 3 > $OUT/bench16_thr$thr.json 2> $OUT/bench16_thr$thr.err
  python - $OUT/bench16_thr$thr.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["breakdown_s"])
PY
done
