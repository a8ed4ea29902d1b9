set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
: > gpurun_out/spmv_groups.jsonl
for mat in lap c4 c5; do for g in 4 6; do for st in 6 8; do
  GA_SPMV_GROUPS=$g GA_SPMV_STAGES=$st python profiles/microbench/spmv_bench.py $mat 2>/dev/null | sed "s/}/, \"groups\": $g}/" >> gpurun_out/spmv_groups.jsonl
done; done; done
cat gpurun_out/spmv_groups.jsonl
timeout 300 python bench.py --no-cpu > gpurun_out/bench_vq2.json 2> gpurun_out/bench_vq2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_vq2.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["config"]["ms_per_cycle_breakdown"], d["roofline"]["achieved"], d["e2e"])
PY
