#!/bin/bash
# Round 2, call 2: where does the Dot2 build lose its 15 %?  per-j step-kernel times (dot2 vs plain), in-kernel timeline,
# ncu full capture (source-level stalls) of k_gs_step (Dot2) and k_vq.
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
P=genericarpack.jl_b200
python profiles/microbench/gs_per_j.py f64 4000 40 > $OUT/perj_dot2.json 2> $OUT/perj_dot2.err; echo "perj dot2 rc=$?"
GA_B200_LIB=$PWD/$P/libgarpack_b200_plain.so python profiles/microbench/gs_per_j.py f64 4000 40 > $OUT/perj_plain.json 2> $OUT/perj_plain.err; echo "perj plain rc=$?"
python - <<'PY'
import json
a=json.load(open("gpurun_out/r02/perj_dot2.json"))["per_j_us_GBps"]; b=json.load(open("gpurun_out/r02/perj_plain.json"))["per_j_us_GBps"]
print("j, dot2 us, plain us, dot2 GB/s, plain GB/s")
for x,y in zip(a,b):
    if x[0] in (1,4,8,11,16,20,24,28,31,32,33,36,40): print(x[0], x[1], y[1], x[2], y[2])
print("sum11..40", sum(x[1] for x in a if x[0]>=11), sum(y[1] for y in b if y[0]>=11))
PY
python profiles/microbench/trace_gs.py > $OUT/trace_dot2.log 2>&1; echo "trace rc=$?"; cat $OUT/trace_dot2.log
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
timeout 300 $CMD > $OUT/plainrun.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gs_step -s 120 -c 3 -o $OUT/prof_dot2 $CMD > $OUT/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 $OUT/ncu_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vq -s 1 -c 1 -o $OUT/prof_vq $CMD > $OUT/ncu_vq.log 2>&1
echo "ncu vq rc=$?"; tail -3 $OUT/ncu_vq.log
ls -la $OUT
