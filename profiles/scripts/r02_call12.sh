#!/bin/bash
# Round 2, call 12: two consumer sets for short bases (tests, per-j A/B at 16M and 2M rows, c64), bench, other configs, blocked C5 ncu.
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
GA_TEST_SKIP_C5_FULL=1 timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu12.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu12.log
for m in 4000 1414; do
  python profiles/microbench/gs_per_j.py f64 $m 40 > $OUT/perj12_${m}_two.json 2> $OUT/perj12_${m}_two.err
  GA_GS_TWO_SETS=0 python profiles/microbench/gs_per_j.py f64 $m 40 > $OUT/perj12_${m}_one.json 2> $OUT/perj12_${m}_one.err
done
python profiles/microbench/gs_per_j.py c64 2000 30 > $OUT/perj12_c64_two.json 2> $OUT/perj12_c64_two.err
GA_GS_TWO_SETS=0 python profiles/microbench/gs_per_j.py c64 2000 30 > $OUT/perj12_c64_one.json 2> $OUT/perj12_c64_one.err
GA_DOTS=plain python profiles/microbench/gs_per_j.py f64 4000 40 > $OUT/perj12_plain_two.json 2> $OUT/perj12_plain_two.err
python - <<'PY'
import json
for name in ("perj12_4000_two","perj12_4000_one","perj12_1414_two","perj12_1414_one","perj12_c64_two","perj12_c64_one","perj12_plain_two"):
    try:
        x=json.load(open("gpurun_out/r02/%s.json"%name))["per_j_us_GBps"]
        print(name, [(r[0],r[1],r[2]) for r in x if r[0] in (4,8,11,14,16,20,24,25,28,32)], "sum11.. us", round(sum(r[1] for r in x if r[0]>=11)), "sum7..30", round(sum(r[1] for r in x if 7<=r[0]<=30)))
    except Exception as e: print(name, "ERR", e)
PY
timeout 600 python bench.py --no-cpu --no-e2e --no-tts --steps 10 --warmup 3 > $OUT/bench12.json 2> $OUT/bench12.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench12.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["config"]["ms_per_cycle_breakdown"], d["clocks"], d["parity_gate"]["passed"])
PY
timeout 600 python profiles/run_configs.py c1 c3 c4 > $OUT/configs12.log 2>&1; echo "configs rc=$?"; cp gpurun_out/configs.json $OUT/configs12.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02/configs12.json"))
for k,v in d.items(): print(k, {kk: v[kk] for kk in ("solve_s","restarts","lanczos_steps","steps_per_s","gs_GBps") if kk in v})
PY
