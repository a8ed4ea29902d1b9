#!/bin/bash
# Round-1 Float32 evidence: whole GPU suite, f32/f64 bench of the C2 operator, ncu launch list + one full capture of
# the Float32 step kernel and the Float32 streamed SpMV (bench first without ncu, then under ncu).
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_f32round.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_f32round.log
timeout 80 python profiles/scripts/f32_bench.py 4000 8 > gpurun_out/f32_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/f32_bench.log | cut -c1-900
CMD="python profiles/scripts/f32_bench.py 4000 2 f32"
timeout 60 $CMD > gpurun_out/f32_plain.log 2>&1 &&
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"k_gs_step|k_spmv_stream" -s 260 -c 3 -o gpurun_out/f32_prof $CMD > gpurun_out/f32_ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out/f32_prof.ncu-rep 2>/dev/null
