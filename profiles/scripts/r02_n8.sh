#!/bin/bash
# Round 2, 8 GPUs: parity checks under torchrun, bench line at n = 16M (gate + tts), steps/s at the 64M target size, in-kernel timeline.
set -u
N=${1:-8}
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR tests/dist_gpu_check.py 300 > $OUT/dist_check_n$N.json 2> $OUT/dist_check_n$N.err; echo "dist check rc=$?"
tail -1 $OUT/dist_check_n$N.json
( time timeout 400 $TR bench.py --gpus $N --no-cpu --trace > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err ) 2>&1 | tail -3; echo "bench rc=$?"
( time timeout 300 $TR bench.py --gpus $N --no-cpu --no-e2e --no-tts --no-gate --grid 8000 --steps 5 > $OUT/bench_64M_n$N.json 2> $OUT/bench_64M_n$N.err ) 2>&1 | tail -3; echo "bench 64M rc=$?"
for f in $OUT/bench_n$N.json $OUT/bench_64M_n$N.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["gs_share_of_step_time"], d["roofline"]["whole_step_achieved_GBps"])
    print("e2e", d.get("e2e")); print("gate", d.get("parity_gate")); print("tts", d.get("time_to_solution")); print("timeline", d.get("gs_step_timeline_us"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
tail -2 $OUT/bench_n$N.err
