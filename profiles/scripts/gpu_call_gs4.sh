set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python profiles/microbench/gs_per_j.py f64 4000 40 > gpurun_out/perj_f64_16M.json
GA_GS_PF=0 python profiles/microbench/gs_per_j.py f64 4000 40 > gpurun_out/perj_f64_16M_pf0.json
GA_GS_PF=8 python profiles/microbench/gs_per_j.py f64 4000 40 > gpurun_out/perj_f64_16M_pf8.json
python profiles/microbench/gs_per_j.py c64 2000 40 > gpurun_out/perj_c64_4M.json
GA_GS_PF=0 python profiles/microbench/gs_per_j.py c64 2000 40 > gpurun_out/perj_c64_4M_pf0.json
python profiles/microbench/gs_per_j.py f64 1414 40 > gpurun_out/perj_f64_2M.json
python - <<'PY'
import json
for f in ["perj_f64_16M","perj_f64_16M_pf0","perj_f64_16M_pf8","perj_c64_4M","perj_c64_4M_pf0","perj_f64_2M"]:
    d=json.load(open("gpurun_out/%s.json"%f))
    print(f, [(j,t) for j,t,g in d["per_j_us_GBps"] if j in (1,8,11,16,20,24,28,31,32,33,36,40)], "sum11..40=%.0f"%sum(t for j,t,g in d["per_j_us_GBps"] if j>=11), "sum7..30=%.0f"%sum(t for j,t,g in d["per_j_us_GBps"] if 7<=j<=30))
PY
