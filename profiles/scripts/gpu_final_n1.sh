#!/bin/bash
# Final round-1 single-GPU evidence: tests, bench (+CPU baseline), reference arm, ncu launch list + full captures,
# the other BASELINE configs, and the n = 64M steps/s on one GPU.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 305 -c 200 --csv --log-file gpurun_out/final_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gs_step|k_spmv_stream|k_vq" -s 270 -c 6 -o gpurun_out/final_prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
timeout 400 python profiles/run_configs.py c1 c3 c4 c5 > gpurun_out/final_configs.log 2>&1; echo "configs rc=$?"; cp gpurun_out/configs.json gpurun_out/final_configs.json
timeout 400 python bench.py --grid 8000 --no-cpu --no-e2e --steps 5 > gpurun_out/final_bench_64M_n1.json 2> gpurun_out/final_bench_64M_n1.err; echo "64M rc=$?"
python - <<'PY'
import json
for f in ["final_bench_n1","final_bench_ref","final_bench_64M_n1"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("roofline") or {}).get("achieved"), d.get("e2e",{}) and d["e2e"].get("value"), d.get("cpu_baseline"))
    except Exception as e: print(f, "ERR", e)
PY
