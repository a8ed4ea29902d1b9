#!/bin/bash
# Round 2, call 14: streaming two-group V*Q kernel (tests, A/B), ncu capture of it and of the blocked C5 SpMV.
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
GA_TEST_SKIP_C5_FULL=1 timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu14.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu14.log
for kev in 5 10 14; do
  python profiles/microbench/vq_bench.py f64 4000 40 $kev 2>&1 | tail -1
  GA_VQ_STREAM=0 python profiles/microbench/vq_bench.py f64 4000 40 $kev 2>&1 | tail -1
done | tee $OUT/vq_bench14.jsonl
python profiles/microbench/vq_bench.py f64 1414 40 10 2>&1 | tail -1 | tee -a $OUT/vq_bench14.jsonl
GA_VQ_STREAM=0 python profiles/microbench/vq_bench.py f64 1414 40 10 2>&1 | tail -1 | tee -a $OUT/vq_bench14.jsonl
timeout 300 python bench.py --no-cpu --no-e2e --no-tts --steps 10 --warmup 3 > $OUT/bench14.json 2> $OUT/bench14.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench14.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["config"]["ms_per_cycle_breakdown"], d["clocks"], d["parity_gate"]["passed"])
PY
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-tts --no-gate"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vq -s 1 -c 1 -o $OUT/prof14_vq $CMD > $OUT/ncu14_vq.log 2>&1; echo "ncu vq rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmv_stream -s 60 -c 10 -o $OUT/prof14_c5 python profiles/microbench/c5_normal_bench.py 0.5 blocked > $OUT/ncu14_c5.log 2>&1; echo "ncu c5 rc=$?"
