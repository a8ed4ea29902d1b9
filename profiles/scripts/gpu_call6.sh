set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
: > gpurun_out/spmv_auto.jsonl
for mat in lap c4 c5 c1; do python profiles/microbench/spmv_bench.py $mat 2>/dev/null >> gpurun_out/spmv_auto.jsonl; done
cat gpurun_out/spmv_auto.jsonl
timeout 300 python bench.py > gpurun_out/bench_auto.json 2> gpurun_out/bench_auto.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_auto.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["config"]["ms_per_cycle_breakdown"], d["roofline"]["achieved"], d["e2e"], d["cpu_baseline"])
PY
