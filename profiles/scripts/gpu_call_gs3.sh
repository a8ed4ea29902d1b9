set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "lanczos or eigs_kat or laplacian" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python profiles/microbench/gs_per_j.py f64 4000 40 > gpurun_out/perj_f64_16M.json
python profiles/microbench/gs_per_j.py c64 2000 40 > gpurun_out/perj_c64_4M.json
python profiles/microbench/gs_per_j.py f64 1414 40 > gpurun_out/perj_f64_2M.json
python - <<'PY'
import json
for f in ["perj_f64_16M","perj_c64_4M","perj_f64_2M"]:
    d=json.load(open("gpurun_out/%s.json"%f))
    print(f, [(j,t) for j,t,g in d["per_j_us_GBps"] if j in (1,8,9,16,17,24,25,31,32,33,40)])
PY
