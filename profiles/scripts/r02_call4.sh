#!/bin/bash
# Round 2, call 4: TMA-staged V*Q kernel (tests + timing, 1 KiB vs 2 KiB segments), 1 KiB-segment step kernel for j > 32 (A/B).
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
P=genericarpack.jl_b200
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu4.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_gpu4.log
for dt in f64 c64 dd f32; do
  mm=4000; [ $dt = c64 ] && mm=2000; [ $dt = dd ] && mm=2000
  python profiles/microbench/vq_bench.py $dt $mm 40 10 2>&1 | tail -1
  GA_VQ_MIN2K=2 python profiles/microbench/vq_bench.py $dt $mm 40 10 2>&1 | tail -1
done | tee $OUT/vq_bench.jsonl
python profiles/microbench/vq_bench.py f64 4000 40 20 2>&1 | tail -1 | tee -a $OUT/vq_bench.jsonl
python profiles/microbench/vq_bench.py f64 1414 40 10 2>&1 | tail -1 | tee -a $OUT/vq_bench.jsonl
GA_B200_LIB=$PWD/$P/libgarpack_b200_seg1k.so GA_GS_SEG1024_FROM=33 python profiles/microbench/gs_per_j.py f64 4000 40 > $OUT/perj4_seg1k.json 2> $OUT/perj4_seg1k.err; echo "perj seg1k rc=$?"
GA_B200_LIB=$PWD/$P/libgarpack_b200_seg1k.so GA_GS_SEG1024_FROM=33 GA_DOTS=plain python profiles/microbench/gs_per_j.py f64 4000 40 > $OUT/perj4_seg1k_plain.json 2> $OUT/perj4_seg1k_plain.err; echo "perj seg1k plain rc=$?"
python profiles/microbench/gs_per_j.py f64 4000 40 > $OUT/perj4_exact.json 2> $OUT/perj4_exact.err
python - <<'PY'
import json
def L(f):
    try: return json.load(open("gpurun_out/r02/%s.json"%f))["per_j_us_GBps"]
    except Exception as e: print(f, "ERR", e); return None
for name in ("perj4_exact","perj4_seg1k","perj4_seg1k_plain"):
    x=L(name)
    if x: print(name, [(r[0],r[1],r[2]) for r in x if r[0] in (28,31,32,33,34,36,38,40)], "sum11..40 us", round(sum(r[1] for r in x if r[0]>=11)))
PY
timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 > $OUT/bench4_exact.json 2> $OUT/bench4_exact.err; echo "bench exact rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench4_exact.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["config"]["ms_per_cycle_breakdown"], d["clocks"])
PY
