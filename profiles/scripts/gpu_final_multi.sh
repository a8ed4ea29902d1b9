#!/bin/bash
# Final round-1 multi-GPU evidence.  usage: gpu_final_multi.sh N [check]
#   check = 1: torchrun parity checks (eigs with halo exchange, svds with the sharded normal operator) first
# then the default bench (C2, n = 16M) and the n = 64M target size on N GPUs.
set -u
N=${1:-2}
CHECK=${2:-0}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$CHECK" = "1" ]; then
  timeout 200 $TR tests/dist_gpu_check.py 300 > gpurun_out/final_dist_check_n$N.json 2> gpurun_out/final_dist_check_n$N.err; echo "dist check rc=$?"
  tail -1 gpurun_out/final_dist_check_n$N.json | cut -c1-600
  timeout 200 $TR tests/dist_gpu_check_normal.py > gpurun_out/final_dist_normal_n$N.json 2> gpurun_out/final_dist_normal_n$N.err; echo "dist normal rc=$?"
  tail -1 gpurun_out/final_dist_normal_n$N.json | cut -c1-600
fi
timeout 200 $TR bench.py --gpus $N --no-cpu > gpurun_out/final_bench_n$N.json 2> gpurun_out/final_bench_n$N.err; echo "bench rc=$?"
timeout 300 $TR bench.py --gpus $N --grid 8000 --steps 5 --no-cpu --no-e2e > gpurun_out/final_bench_64M_n$N.json 2> gpurun_out/final_bench_64M_n$N.err; echo "bench 64M rc=$?"
for f in gpurun_out/final_bench_n$N.json gpurun_out/final_bench_64M_n$N.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["gs_share_of_step_time"], d.get("e2e") and d["e2e"]["value"], d["clocks"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
tail -n 3 gpurun_out/final_bench_n$N.err gpurun_out/final_bench_64M_n$N.err
