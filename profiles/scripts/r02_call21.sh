#!/bin/bash
# Round 2, call 21: one rendezvous per phase in the step kernel (tests, timeline, per-j at 2M and 16M rows, C1, bench).
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
GA_TEST_SKIP_C5_FULL=1 timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu21.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu21.log
python profiles/microbench/trace_gs.py > $OUT/trace21.log 2>&1; cat $OUT/trace21.log
python profiles/microbench/gs_per_j.py f64 1414 40 > $OUT/perj21_2M.json 2> $OUT/perj21_2M.err
python profiles/microbench/gs_per_j.py f64 4000 40 > $OUT/perj21_16M.json 2> $OUT/perj21_16M.err
python - <<'PY'
import json
for name in ("perj21_2M","perj21_16M"):
    x=json.load(open("gpurun_out/r02/%s.json"%name))["per_j_us_GBps"]
    print(name, [(r[0],r[1],r[2]) for r in x if r[0] in (4,8,11,16,20,24,28,32,36,40)], "sum11..40 us", round(sum(r[1] for r in x if r[0]>=11)))
PY
timeout 300 python bench.py --no-cpu --steps 10 --warmup 3 > $OUT/bench21.json 2> $OUT/bench21.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench21.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["e2e"]["value"], d["time_to_solution"]["seconds"], d["parity_gate"]["passed"], d["clocks"])
PY
timeout 300 python profiles/run_configs.py c1 > $OUT/configs21.log 2>&1; python - <<'PY'
import json
d=json.load(open("gpurun_out/configs.json"))
for k,v in d.items(): print(k, {kk: v[kk] for kk in ("solve_s","us_per_lanczos_step","restarts","lanczos_steps") if kk in v})
PY
