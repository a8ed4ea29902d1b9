#!/bin/bash
# Next step for profiles/r02/README.md section 17: build the two-set 1 KiB-segment step kernel as a VARIANT library
# (the shipped libgarpack_b200.so is not touched) and run one wide-basis Lanczos pass with it.  compute-sanitizer is closed on this
# pool (tried: the wrapper refuses it), so the run prints the per-j table or the CUDA error; narrow down with GA_GS_SEG1024_FROM /
# small grids / in-kernel asserts.
#   here (no GPU):   bash profiles/scripts/debug_two_sets_1k.sh build
#   on the GPU box:  gpurun --timeout 600 -- 'bash profiles/scripts/debug_two_sets_1k.sh run'
set -eu
cd "$(dirname "$0")/../.."
case "${1:-build}" in
  build)
    git apply --check profiles/r02/experiments/two_sets_1k_segments.diff
    git apply profiles/r02/experiments/two_sets_1k_segments.diff
    # the diff only holds the launcher; the switch lives in the context (see the experiment's description)
    sed -i 's|  bool gs_two_sets = true; |  bool gs_two_sets_1k = true;\n  bool gs_two_sets = true; |' genericarpack.jl_b200/csrc/ga_ctx.cuh
    GA_B200_VARIANT=two1k python genericarpack.jl_b200/build.py
    git checkout -- genericarpack.jl_b200/csrc/k_gs_step.cuh genericarpack.jl_b200/csrc/ga_ctx.cuh
    ls -la genericarpack.jl_b200/libgarpack_b200_two1k.so
    ;;
  run)
    mkdir -p gpurun_out/r02
    export GA_B200_LIB=$PWD/genericarpack.jl_b200/libgarpack_b200_two1k.so
    timeout 200 python profiles/microbench/gs_per_j.py f64 1000 40 > gpurun_out/r02/two1k_perj.log 2>&1 || true
    tail -5 gpurun_out/r02/two1k_perj.log
    ;;
esac
