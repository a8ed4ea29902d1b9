#!/bin/bash
# Round 2, call 6: register-blocked V*Q kernel (tests + timing), the new bench line (parity gate, time to solution, CPU legs).
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
GA_TEST_SKIP_C5_FULL=1 timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu6.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu6.log
for dt in f64 c64 dd f32; do
  mm=4000; [ $dt = c64 ] && mm=2000; [ $dt = dd ] && mm=2000
  python profiles/microbench/vq_bench.py $dt $mm 40 10 2>&1 | tail -1
done | tee $OUT/vq_bench6.jsonl
python profiles/microbench/vq_bench.py f64 4000 40 20 2>&1 | tail -1 | tee -a $OUT/vq_bench6.jsonl
python profiles/microbench/vq_bench.py f64 4000 40 5 2>&1 | tail -1 | tee -a $OUT/vq_bench6.jsonl
python profiles/microbench/vq_bench.py f64 1414 40 10 2>&1 | tail -1 | tee -a $OUT/vq_bench6.jsonl
( time timeout 900 python bench.py > $OUT/bench6.json 2> $OUT/bench6.err ) 2>&1 | tail -3; echo "bench rc=$?"
tail -3 $OUT/bench6.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench6.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["config"]["ms_per_cycle_breakdown"])
print("e2e", d["e2e"]); print("gate", d["parity_gate"]); print("tts", d.get("time_to_solution")); print("cpu", d["cpu_baseline"])
PY
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench6_ref.json 2> $OUT/bench6_ref.err ) 2>&1 | tail -3
cat $OUT/bench6_ref.json | cut -c1-600
