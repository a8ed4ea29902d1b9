#!/bin/bash
# Round 2, Float32 vectors on several GPUs: parity checks under torchrun (peer-memory and NCCL paths) + the Float64 check
# again (the halo wire format code changed).  usage: r02_n2_f32.sh N
set -u
N=${1:-2}
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515"
timeout 240 $TR tests/dist_gpu_check_f32.py > $OUT/dist_f32_n$N.json 2> $OUT/dist_f32_n$N.err; echo "dist f32 rc=$?"
tail -1 $OUT/dist_f32_n$N.json | cut -c1-1500; grep -v "OMP_NUM\|\*\*\*\*" $OUT/dist_f32_n$N.err | tail -15
GA_NCCL_ONLY=1 timeout 240 $TR tests/dist_gpu_check_f32.py > $OUT/dist_f32_nccl_n$N.json 2> $OUT/dist_f32_nccl_n$N.err; echo "dist f32 nccl rc=$?"
tail -1 $OUT/dist_f32_nccl_n$N.json | cut -c1-1500; grep -v "OMP_NUM\|\*\*\*\*" $OUT/dist_f32_nccl_n$N.err | tail -15
timeout 240 $TR tests/dist_gpu_check.py 300 > $OUT/dist_check_b_n$N.json 2> $OUT/dist_check_b_n$N.err; echo "dist f64 rc=$?"
tail -1 $OUT/dist_check_b_n$N.json | cut -c1-700
GA_NCCL_ONLY=1 timeout 240 $TR tests/dist_gpu_check_normal.py > $OUT/dist_normal_nccl_b_n$N.json 2> $OUT/dist_normal_nccl_b_n$N.err; echo "dist normal nccl rc=$?"
tail -1 $OUT/dist_normal_nccl_b_n$N.json | cut -c1-500
