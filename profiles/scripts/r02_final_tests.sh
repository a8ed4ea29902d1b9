#!/bin/bash
# Round 2: the complete GPU suite (incl. C5 at 50M x 2M) + smoke on the final tree.
set -u
mkdir -p gpurun_out/r02
OUT=gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > $OUT/pytest_gpu_final2.log 2>&1; echo "pytest rc=$?"; tail -10 $OUT/pytest_gpu_final2.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke_final2.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke_final2.log
