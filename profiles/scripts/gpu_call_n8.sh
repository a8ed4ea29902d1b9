#!/bin/bash
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR tests/dist_gpu_check.py 300 > gpurun_out/dist_check_n$N.json 2> gpurun_out/dist_check_n$N.err; echo "dist check rc=$?"
tail -1 gpurun_out/dist_check_n$N.json
timeout 200 $TR bench.py --gpus $N --no-cpu --trace > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
GA_NO_STEP_KERNEL=1 timeout 200 $TR bench.py --gpus $N --no-cpu --no-e2e > gpurun_out/bench_n${N}_nostep.json 2> gpurun_out/bench_n${N}_nostep.err; echo "bench nostep rc=$?"
for f in gpurun_out/bench_n$N.json gpurun_out/bench_n${N}_nostep.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["gs_share_of_step_time"], d.get("e2e"), d.get("gs_step_timeline_us"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
