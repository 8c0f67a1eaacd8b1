#!/bin/bash
# gpurun (r1e, 3rd): GPU tests, kernel micro-bench incl. the single-word stage-2 variants, per-config device times.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scripts/kbench.py > gpurun_out/kbench.log 2>&1; cat gpurun_out/kbench.log
python scripts/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; cat gpurun_out/configs.jsonl; tail -5 gpurun_out/configs.err
