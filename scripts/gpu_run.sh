#!/bin/bash
# One launcher for everything that runs on the GPU box:  gpurun [--gpus N] -- 'bash scripts/gpu_run.sh <what> [args]'
#   tests                 GPU suite (product library) + the variant tests on the experiments library + smoke
#   kbench [pw|prod]      per-kernel timings of every variant (experiments library) -> gpurun_out/kbench.log
#   prof <regex> [n]      ncu --set full of the kernels matching <regex> while kbench runs one frame -> gpurun_out/prof_<regex>.ncu-rep
#   bench [N]             the driver's bench line at N GPUs (torchrun for N > 1) -> gpurun_out/bench_n<N>.json
#   micro                 the gather micro-benchmarks (binaries built here by nvcc travel in build/)
#   e2e [N]               host-to-host probe with every rank copying at once + the bench line at N
#   sanitize              compute-sanitizer memcheck / racecheck / initcheck / synccheck over the r2 kernels on small inputs
#   final                 the evidence pass (tests, per-config table, bench + reference arm, ncu captures, launch list)
set -u
mkdir -p gpurun_out
what=${1:-tests}
shift || true
case "$what" in
  tests)
    python -m pytest tests -m gpu -x -q 2>&1 | tail -4
    LERF_B200_EXPERIMENTS=1 python -m pytest tests -m gpu -x -q -k "variants or pipeline_kernel or swizzle" 2>&1 | tail -2
    python scripts/pw_check.py 2>&1 | tail -2
    python __graft_entry__.py smoke 2>&1 | tail -1 ;;
  kbench)
    KB_FRAMES=8 KB_ONLY=${1:-} python scripts/kbench.py > gpurun_out/kbench.log 2>&1; cat gpurun_out/kbench.log ;;
  prof)
    KREGEX=${1:?kernel regex} NCAP=${2:-8} OUT=prof_${1} KB_ONLY=pw bash scripts/gpu_prof_one.sh ;;
  bench)
    N=${1:-1} bash scripts/gpu_bench.sh ;;
  micro)
    bash scripts/gpu_r2_micro.sh ;;
  e2e)
    N=${1:-8} bash scripts/gpu_e2e_probe.sh ;;
  sanitize)
    for tool in memcheck racecheck initcheck synccheck; do
      compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_probe.py > gpurun_out/sanitizer_$tool.log 2>&1
      echo "$tool rc=$?"; tail -3 gpurun_out/sanitizer_$tool.log
    done ;;
  final)
    bash scripts/gpu_r2_final.sh ;;
  *) echo "unknown: $what"; exit 2 ;;
esac
