#!/bin/bash
# Round 2, call 3: branch-free step kernel (per-NCT instantiations) -- GPU tests, per-j times exact/plain, other dtypes, bench.
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu3.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_gpu3.log
python profiles/microbench/gs_per_j.py f64 4000 40 > $OUT/perj3_exact.json 2> $OUT/perj3_exact.err; echo "perj exact rc=$?"
GA_DOTS=plain python profiles/microbench/gs_per_j.py f64 4000 40 > $OUT/perj3_plain.json 2> $OUT/perj3_plain.err; echo "perj plain rc=$?"
python profiles/microbench/gs_per_j.py c64 2000 40 > $OUT/perj3_c64.json 2> $OUT/perj3_c64.err; echo "perj c64 rc=$?"
python profiles/microbench/gs_per_j.py f64 1414 40 > $OUT/perj3_2M.json 2> $OUT/perj3_2M.err; echo "perj 2M rc=$?"
python - <<'PY'
import json
def L(f):
    try: return json.load(open("gpurun_out/r02/%s.json"%f))["per_j_us_GBps"]
    except Exception as e: print(f, "ERR", e); return None
a,b,c,d=L("perj3_exact"),L("perj3_plain"),L("perj3_c64"),L("perj3_2M")
for name,x in (("exact",a),("plain",b),("c64",c),("2M",d)):
    if x: print(name, [(r[0],r[1],r[2]) for r in x if r[0] in (1,8,11,16,20,24,28,31,32,33,36,40)], "sum11..40 us", round(sum(r[1] for r in x if r[0]>=11)))
PY
python profiles/microbench/trace_gs.py > $OUT/trace3.log 2>&1; echo "trace rc=$?"; cat $OUT/trace3.log
timeout 300 python bench.py --no-cpu --steps 10 --warmup 3 > $OUT/bench3_exact.json 2> $OUT/bench3_exact.err; echo "bench exact rc=$?"
timeout 300 python bench.py --no-cpu --no-e2e --dots plain --steps 10 --warmup 3 > $OUT/bench3_plain.json 2> $OUT/bench3_plain.err; echo "bench plain rc=$?"
python - <<'PY'
import json
for v in ("exact","plain"):
    try:
        d=json.loads(open("gpurun_out/r02/bench3_%s.json"%v).read().strip().splitlines()[-1])
        print(v, d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d.get("e2e"), d["clocks"])
    except Exception as e: print(v, "ERR", e)
PY
