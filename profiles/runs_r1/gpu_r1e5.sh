#!/bin/bash
# gpurun (r1e, 5th): GPU tests (incl. the tile / fast-warp kernels), per-config device times.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^Scale\|^Set5" | tail -45
python scripts/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; cat gpurun_out/configs.jsonl; tail -5 gpurun_out/configs.err
