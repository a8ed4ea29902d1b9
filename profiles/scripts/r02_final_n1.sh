#!/bin/bash
# Round 2, final single-GPU evidence: full GPU test suite (incl. C5 at 50M x 2M), smoke, default bench + reference arm,
# ncu launch list + full captures of the shipped kernels, the other BASELINE configs, Float32.
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
GA_TEST_SKIP_C5_FULL=1 timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > $OUT/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -12 $OUT/pytest_gpu_final.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_final.log
( time timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench_final_n1.json 2> $OUT/bench_final_n1.err ) 2>&1 | tail -3; echo "bench rc=$?"
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_final_ref.json 2> $OUT/bench_final_ref.err ) 2>&1 | tail -3; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench_final_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["config"]["ms_per_cycle_breakdown"], d["clocks"])
print("e2e", d["e2e"]); print("gate", d["parity_gate"]["passed"]); print("tts", {k: d["time_to_solution"][k] for k in ("seconds","restarts","converged","max_rel_err_vs_analytic")})
print("cpu", d["cpu_baseline"])
r=json.loads(open("gpurun_out/r02/bench_final_ref.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["cpu_baseline"]["cores"])
PY
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-tts --no-gate"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 305 -c 200 --csv --log-file $OUT/launches_final.csv $CMD > $OUT/ncu_list_final.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:k_gs_step -s 120 -c 3 -o $OUT/prof_final_gs $CMD > $OUT/ncu_final_gs.log 2>&1; echo "ncu gs rc=$?"
ncu -i $OUT/prof_final_gs.ncu-rep --page raw --csv > $OUT/prof_final_gs_raw.csv 2>/dev/null; rm -f $OUT/prof_final_gs.ncu-rep
python profiles/microbench/upload_bench.py > $OUT/upload_bench.json 2>&1; tail -1 $OUT/upload_bench.json
timeout 600 python profiles/run_configs.py c1 c3 c4 > $OUT/configs_final.log 2>&1; echo "configs rc=$?"; cp gpurun_out/configs.json $OUT/configs_final.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02/configs_final.json"))
for k,v in d.items(): print(k, {kk: v[kk] for kk in ("solve_s","cold_first_solve_s","us_per_lanczos_step","restarts","lanczos_steps","steps_per_s","gs_GBps") if kk in v})
PY
timeout 400 python bench.py --grid 8000 --no-cpu --no-e2e --no-tts --no-gate --steps 5 > $OUT/bench_64M_n1.json 2> $OUT/bench_64M_n1.err; echo "64M rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench_64M_n1.json").read().strip().splitlines()[-1])
print("64M N=1", d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["whole_step_achieved_GBps"])
PY
timeout 300 python profiles/scripts/f32_bench.py > $OUT/f32_final.log 2>&1; echo "f32 rc=$?"; tail -3 $OUT/f32_final.log | cut -c1-600
