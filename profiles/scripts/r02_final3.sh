#!/bin/bash
# Round 2, last check of the final tree on one GPU: complete GPU suite (incl. C5 at 50M x 2M), smoke, default bench line.
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 > $OUT/pytest_gpu_final3.log 2>&1; echo "pytest rc=$?"; tail -10 $OUT/pytest_gpu_final3.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke_final3.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke_final3.log
( time timeout 600 python bench.py > $OUT/bench_final3_n1.json 2> $OUT/bench_final3_n1.err ) 2>&1 | tail -3; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench_final3_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["clocks"])
print("e2e", d["e2e"]["value"]); print("gate", d["parity_gate"]["passed"]); print("tts", {k: d["time_to_solution"][k] for k in ("seconds","restarts","converged","max_rel_err_vs_analytic")})
print("cpu", d["cpu_baseline"]["value"], d["gpu_launches"])
PY
