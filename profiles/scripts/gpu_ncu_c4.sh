set -u
CMD="python profiles/run_configs.py c4"
timeout 300 $CMD > gpurun_out/c4_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gs_step -s 300 -c 1 -o gpurun_out/prof_c4_gs $CMD > gpurun_out/ncu_c4.log 2>&1
echo "ncu rc=$?"; grep -E "steps_per_s|gs_GBps" gpurun_out/c4_plain.log
