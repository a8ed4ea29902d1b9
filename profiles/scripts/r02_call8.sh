#!/bin/bash
# Round 2, call 8: step kernel with on-demand second-refinement dots (tests, per-j, bench) + ncu full captures of k_gs_step / k_vq / k_spmv_stream.
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
GA_TEST_SKIP_C5_FULL=1 timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu8.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu8.log
python profiles/microbench/gs_per_j.py f64 4000 40 > $OUT/perj8_exact.json 2> $OUT/perj8_exact.err; echo "perj exact rc=$?"
GA_DOTS=plain python profiles/microbench/gs_per_j.py f64 4000 40 > $OUT/perj8_plain.json 2> $OUT/perj8_plain.err; echo "perj plain rc=$?"
python profiles/microbench/gs_per_j.py f64 1414 40 > $OUT/perj8_2M.json 2> $OUT/perj8_2M.err
python - <<'PY'
import json
def L(f):
    try: return json.load(open("gpurun_out/r02/%s.json"%f))["per_j_us_GBps"]
    except Exception as e: print(f, "ERR", e); return None
for name in ("perj8_exact","perj8_plain","perj8_2M"):
    x=L(name)
    if x: print(name, [(r[0],r[1],r[2]) for r in x if r[0] in (1,8,11,16,20,24,28,32,33,36,40)], "sum11..40 us", round(sum(r[1] for r in x if r[0]>=11)))
PY
python profiles/microbench/trace_gs.py > $OUT/trace8.log 2>&1; cat $OUT/trace8.log
timeout 300 python bench.py --no-cpu --no-e2e --no-tts --no-gate --steps 10 --warmup 3 > $OUT/bench8_exact.json 2> $OUT/bench8_exact.err; echo "bench exact rc=$?"
timeout 300 python bench.py --no-cpu --no-e2e --no-tts --no-gate --dots plain --steps 10 --warmup 3 > $OUT/bench8_plain.json 2> $OUT/bench8_plain.err; echo "bench plain rc=$?"
python - <<'PY'
import json
for v in ("exact","plain"):
    try:
        d=json.loads(open("gpurun_out/r02/bench8_%s.json"%v).read().strip().splitlines()[-1])
        print(v, d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["config"]["ms_per_cycle_breakdown"], d["clocks"])
    except Exception as e: print(v, "ERR", e)
PY
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-tts --no-gate"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gs_step -s 120 -c 3 -o $OUT/prof8_gs $CMD > $OUT/ncu8_gs.log 2>&1; echo "ncu gs rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vq -s 1 -c 1 -o $OUT/prof8_vq $CMD > $OUT/ncu8_vq.log 2>&1; echo "ncu vq rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_stream -s 100 -c 1 -o $OUT/prof8_spmv $CMD > $OUT/ncu8_spmv.log 2>&1; echo "ncu spmv rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 305 -c 200 --csv --log-file $OUT/launches8.csv $CMD > $OUT/ncu8_list.log 2>&1; echo "ncu list rc=$?"
ls -la $OUT | tail -8
