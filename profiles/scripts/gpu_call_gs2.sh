set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python profiles/microbench/gs_per_j.py f64 4000 40 > gpurun_out/perj_f64_16M.json
python profiles/microbench/gs_per_j.py c64 2000 40 > gpurun_out/perj_c64_4M.json
python profiles/microbench/gs_per_j.py f64 1414 40 > gpurun_out/perj_f64_2M.json
python profiles/microbench/gs_per_j.py dd 1000 40 > gpurun_out/perj_dd_1M.json
cat gpurun_out/perj_*.json | cut -c1-1700
timeout 300 python bench.py --no-cpu > gpurun_out/bench_gs2.json 2> gpurun_out/bench_gs2.err; echo "bench rc=$?"
cut -c1-2000 gpurun_out/bench_gs2.json
