#!/bin/bash
# Round 2, call 1: Dot2 accumulators -- bit-level parity probe, A/B/C timing of the three builds on one box, GPU tests, C2 time to solution.
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
P=genericarpack.jl_b200
timeout 600 python profiles/scripts/parity_probe.py $OUT/parity_probe.json > $OUT/parity_probe.log 2>&1; echo "probe rc=$?"
tail -6 $OUT/parity_probe.log
for v in dot2 fast plain; do
  L=$P/libgarpack_b200_$v.so; [ $v = dot2 ] && L=$P/libgarpack_b200.so
  GA_B200_LIB=$PWD/$L timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 > $OUT/bench_$v.json 2> $OUT/bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
d=json.loads(open("$OUT/bench_$v.json").read().strip().splitlines()[-1])
print("$v", d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["clocks"])
PY
done
GA_B200_LIB=$PWD/$P/libgarpack_b200_fast.so timeout 600 python profiles/scripts/parity_probe.py $OUT/parity_probe_fast.json > $OUT/parity_probe_fast.log 2>&1; echo "probe fast rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
timeout 600 python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 --tts --tts-maxiter 8000 --tts-budget 400 > $OUT/bench_tts.json 2> $OUT/bench_tts.err; echo "tts rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_tts.json").read().strip().splitlines()[-1]); print(d.get("time_to_solution"))
except Exception as e: print("tts parse", e)
PY
