#!/bin/bash
# Round 2, multi-GPU: parity checks under torchrun + bench line (parity gate, time to solution) on N GPUs.  usage: r02_n2.sh N
set -u
N=${1:-2}
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tests/dist_gpu_check.py 300 > $OUT/dist_check_n$N.json 2> $OUT/dist_check_n$N.err; echo "dist check rc=$?"
tail -1 $OUT/dist_check_n$N.json
timeout 300 $TR tests/dist_gpu_check_normal.py > $OUT/dist_normal_n$N.json 2> $OUT/dist_normal_n$N.err; echo "dist normal rc=$?"
tail -1 $OUT/dist_normal_n$N.json | cut -c1-600
if [ $N = 2 ]; then timeout 900 python -m pytest tests/test_gpu_fullsize.py -k two_gpu -x -q 2>&1 | tail -3; fi
( time timeout 600 $TR bench.py --gpus $N --no-cpu > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err ) 2>&1 | tail -3; echo "bench rc=$?"
tail -2 $OUT/bench_n$N.err
python - $OUT/bench_n$N.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["gs_share_of_step_time"])
    print("e2e", d.get("e2e")); print("gate", d.get("parity_gate")); print("tts", d.get("time_to_solution"))
except Exception as e: print("ERR", e)
PY
