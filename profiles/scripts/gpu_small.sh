set -u
for m in 1414 2000; do
python bench.py --m $m --no-cpu --no-e2e > gpurun_out/bench_m$m.json 2>/dev/null
GA_NO_STEP_KERNEL=1 python bench.py --m $m --no-cpu --no-e2e > gpurun_out/bench_m${m}_nostep.json 2>/dev/null
for f in gpurun_out/bench_m$m.json gpurun_out/bench_m${m}_nostep.json; do python - $f <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["gs_share_of_step_time"], d["roofline"]["avg_launch_ms"], d["config"]["ms_per_cycle_breakdown"])
PY
done
done
python profiles/microbench/trace_gs.py
