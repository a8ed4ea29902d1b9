#!/bin/bash
# Round 2: C2 at its stated size (n = 16M) to convergence on N GPUs (tol = 1e-10); at N = 8 also the bench line, the 64M target's steps/s
# and C5 at full size sharded over the 8 GPUs.  usage: r02_tts_multi.sh N
set -u
N=${1:-8}
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
( time timeout 900 $TR bench.py --gpus $N --no-cpu --no-e2e --no-tts --tts-full --tts-budget 600 --steps 5 > $OUT/c2_tts_full_n$N.json 2> $OUT/c2_tts_full_n$N.err ) 2>&1 | tail -3; echo "tts-full rc=$?"
python - $OUT/c2_tts_full_n$N.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["value"], d["roofline"]["frac"], d["parity_gate"]["passed"]); print(d.get("time_to_solution"))
except Exception as e: print("ERR", e)
PY
if [ $N = 8 ]; then
  ( time timeout 300 $TR bench.py --gpus $N --no-cpu > $OUT/bench_final_n$N.json 2> $OUT/bench_final_n$N.err ) 2>&1 | tail -3; echo "bench rc=$?"
  ( time timeout 300 $TR bench.py --gpus $N --no-cpu --no-e2e --no-tts --no-gate --grid 8000 --steps 5 > $OUT/bench_64M_n$N.json 2> $OUT/bench_64M_n$N.err ) 2>&1 | tail -3; echo "bench 64M rc=$?"
  for f in $OUT/bench_final_n$N.json $OUT/bench_64M_n$N.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["whole_step_achieved_GBps"], (d.get("e2e") or {}).get("value"), (d.get("time_to_solution") or {}).get("seconds"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
  done
  ( time timeout 600 $TR profiles/scripts/c5_dist_full.py 1.0 $OUT/c5_dist_full_n$N.json > $OUT/c5_dist_full_n$N.log 2> $OUT/c5_dist_full_n$N.err ) 2>&1 | tail -3; echo "c5 dist rc=$?"
  tail -1 $OUT/c5_dist_full_n$N.log | cut -c1-1200; tail -3 $OUT/c5_dist_full_n$N.err | cut -c1-400
fi
