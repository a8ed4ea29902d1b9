#!/bin/bash
# Round 2: C2 at its stated size (n = 16M) run to convergence on one B200, tol = 1e-10 (the parity bar), wall-clock guard 1300 s.
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
timeout 1500 python bench.py --no-cpu --no-e2e --no-tts --tts-full --tts-budget 1300 --steps 5 --warmup 3 > $OUT/bench_tts_full.json 2> $OUT/bench_tts_full.err; echo "tts-full rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench_tts_full.json").read().strip().splitlines()[-1])
print(d["value"], d["roofline"]["frac"]); print(d.get("time_to_solution"))
PY
tail -3 $OUT/bench_tts_full.err
