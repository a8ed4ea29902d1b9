#!/bin/bash
# Round 2, call 11: tensor-map tiles for wide bases (tests, per-j A/B against the 2-stage 2 KiB rule), upload overlap (e2e), ncu captures
# of the C5 normal-operator SpMV and of the ComplexF64 step kernel.
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
GA_TEST_SKIP_C5_FULL=1 timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu11.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu11.log
python profiles/microbench/gs_per_j.py f64 4000 40 > $OUT/perj11_tmap.json 2> $OUT/perj11_tmap.err
GA_GS_MIN2K=2 python profiles/microbench/gs_per_j.py f64 4000 40 > $OUT/perj11_2stage.json 2> $OUT/perj11_2stage.err
GA_DOTS=plain python profiles/microbench/gs_per_j.py f64 4000 40 > $OUT/perj11_tmap_plain.json 2> $OUT/perj11_tmap_plain.err
python profiles/microbench/gs_per_j.py c64 2000 40 > $OUT/perj11_c64.json 2> $OUT/perj11_c64.err
python profiles/microbench/gs_per_j.py dd 2000 40 > $OUT/perj11_dd.json 2> $OUT/perj11_dd.err
python - <<'PY'
import json
for name in ("perj11_tmap","perj11_2stage","perj11_tmap_plain","perj11_c64","perj11_dd"):
    try:
        x=json.load(open("gpurun_out/r02/%s.json"%name))["per_j_us_GBps"]
        print(name, [(r[0],r[1],r[2]) for r in x if r[0] in (11,16,24,28,32,33,34,36,38,40)], "sum11..40 us", round(sum(r[1] for r in x if r[0]>=11)))
    except Exception as e: print(name, "ERR", e)
PY
timeout 600 python bench.py --no-cpu --steps 10 --warmup 3 > $OUT/bench11.json 2> $OUT/bench11.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench11.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["config"]["ms_per_cycle_breakdown"], d["clocks"])
print("e2e", d["e2e"]); print("tts", {k: d["time_to_solution"][k] for k in ("seconds","restarts","converged")})
PY
CMD="python profiles/run_configs.py c4"
timeout 300 $CMD > $OUT/c4_plain.log 2>&1; tail -2 $OUT/c4_plain.log | cut -c1-600
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gs_step -s 200 -c 2 -o $OUT/prof11_c64 $CMD > $OUT/ncu11_c64.log 2>&1; echo "ncu c64 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmv_stream -s 40 -c 16 -o $OUT/prof11_c5 python profiles/microbench/c5_normal_bench.py 0.5 > $OUT/ncu11_c5.log 2>&1; echo "ncu c5 rc=$?"
ls -la $OUT/*.ncu-rep | tail -4
