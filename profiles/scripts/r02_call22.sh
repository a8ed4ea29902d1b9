#!/bin/bash
# Round 2, call 22 (experiment, reverted: see profiles/r02/README.md section 17 and experiments/two_sets_1k_segments.diff; the GA_GS_TWO_SETS_1K switch exists only with that diff applied): two consumer sets for the 1 KiB-segment step kernel (wide bases, j >= 33): A/B per j, then tests + bench with it on.
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
GA_GS_TWO_SETS_1K=0 timeout 120 python profiles/microbench/gs_per_j.py f64 4000 40 > $OUT/perj22_16M_one.json 2> $OUT/perj22_one.err
GA_GS_TWO_SETS_1K=1 timeout 120 python profiles/microbench/gs_per_j.py f64 4000 40 > $OUT/perj22_16M_two.json 2> $OUT/perj22_two.err
GA_GS_TWO_SETS_1K=0 timeout 120 python profiles/microbench/gs_per_j.py f64 2828 56 > $OUT/perj22_8M56_one.json 2>> $OUT/perj22_one.err
GA_GS_TWO_SETS_1K=1 timeout 120 python profiles/microbench/gs_per_j.py f64 2828 56 > $OUT/perj22_8M56_two.json 2>> $OUT/perj22_two.err
python - <<'PY'
import json
for a in ("16M","8M56"):
    try:
        x=json.load(open("gpurun_out/r02/perj22_%s_one.json"%a))["per_j_us_GBps"]; y=json.load(open("gpurun_out/r02/perj22_%s_two.json"%a))["per_j_us_GBps"]
        print(a, "sum us one/two (j>=11):", round(sum(r[1] for r in x[10:])), round(sum(r[1] for r in y[10:])))
        print(" j>=31:", [(r[0], r[1], s[1]) for r,s in zip(x,y) if r[0]>=31])
    except Exception as e: print(a, "ERR", e)
PY
GA_TEST_SKIP_C5_FULL=1 timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu22.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu22.log
timeout 300 python bench.py --no-cpu > $OUT/bench22.json 2> $OUT/bench22.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench22.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["clocks"]); print("e2e", d["e2e"]["value"], "gate", d["parity_gate"]["passed"], "tts", d["time_to_solution"]["seconds"], d["time_to_solution"]["restarts"])
PY
