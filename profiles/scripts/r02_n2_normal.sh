#!/bin/bash
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tests/dist_gpu_check_normal.py > $OUT/dist_normal_n2_push.json 2> $OUT/dist_normal_n2_push.err; echo "dist normal rc=$?"
tail -1 $OUT/dist_normal_n2_push.json | cut -c1-700
timeout 300 $TR tests/dist_gpu_check.py 300 > $OUT/dist_check_n2_push.json 2> $OUT/dist_check_n2_push.err; echo "dist check rc=$?"
( time timeout 600 $TR profiles/scripts/c5_dist_full.py 0.2 $OUT/c5_dist_02_n2.json > $OUT/c5_dist_02_n2.log 2> $OUT/c5_dist_02_n2.err ) 2>&1 | tail -3
tail -1 $OUT/c5_dist_02_n2.log | cut -c1-700; tail -2 $OUT/c5_dist_02_n2.err | cut -c1-300
