#!/bin/bash
# multi-GPU parity (torchrun) + strong-scaling bench, step kernel on/off.  usage: gpu_call_n2.sh N
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tests/dist_gpu_check.py 300 > gpurun_out/dist_check_n$N.json 2> gpurun_out/dist_check_n$N.err; echo "dist check rc=$?"
tail -1 gpurun_out/dist_check_n$N.json
timeout 300 $TR bench.py --gpus $N --no-cpu > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
GA_NO_STEP_KERNEL=1 timeout 300 $TR bench.py --gpus $N --no-cpu --no-e2e > gpurun_out/bench_n${N}_nostep.json 2> gpurun_out/bench_n${N}_nostep.err; echo "bench nostep rc=$?"
timeout 300 $TR bench.py --gpus $N --no-cpu --no-e2e --nccl-only > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err; echo "bench nccl rc=$?"
for f in gpurun_out/bench_n$N.json gpurun_out/bench_n${N}_nostep.json gpurun_out/bench_n${N}_nccl.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["gs_share_of_step_time"], d.get("e2e"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
tail -3 gpurun_out/bench_n$N.err
