#!/bin/bash
# Round 2, generalised problem on several GPUs: parity check under torchrun, the single-GPU generalised tests (the kernels
# gained a communicator argument) and the normal-operator check (k_push_halo gained the gate).  usage: r02_n2_general.sh N
set -u
N=${1:-2}
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR tests/dist_gpu_check_general.py > $OUT/dist_general_n$N.json 2> $OUT/dist_general_n$N.err; echo "dist general rc=$?"
tail -1 $OUT/dist_general_n$N.json | cut -c1-2000; grep -v "OMP_NUM\|\*\*\*\*" $OUT/dist_general_n$N.err | tail -25
timeout 300 python -m pytest tests/test_gpu_general.py -x -q 2>&1 | tail -3
timeout 240 $TR tests/dist_gpu_check_normal.py > $OUT/dist_normal_c_n$N.json 2> $OUT/dist_normal_c_n$N.err; echo "dist normal rc=$?"
tail -1 $OUT/dist_normal_c_n$N.json | cut -c1-400
