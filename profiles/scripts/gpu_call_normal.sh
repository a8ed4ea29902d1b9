#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
timeout 300 $TR tests/dist_gpu_check_normal.py > gpurun_out/dist_normal_n$N.json 2> gpurun_out/dist_normal_n$N.err; echo "normal fused rc=$?"
tail -1 gpurun_out/dist_normal_n$N.json; grep -v "^\*\|OMP_NUM" gpurun_out/dist_normal_n$N.err | tail -5
GA_NCCL_ONLY=1 timeout 300 $TR tests/dist_gpu_check_normal.py > gpurun_out/dist_normal_n${N}_nccl.json 2> gpurun_out/dist_normal_n${N}_nccl.err; echo "normal nccl rc=$?"
tail -1 gpurun_out/dist_normal_n${N}_nccl.json; grep -v "^\*\|OMP_NUM" gpurun_out/dist_normal_n${N}_nccl.err | tail -5
