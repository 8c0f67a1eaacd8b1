#!/bin/bash
# gpurun (r1e, 2nd): GPU tests, per-config device times, default bench.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scripts/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; cat gpurun_out/configs.jsonl; tail -5 gpurun_out/configs.err
python bench.py > gpurun_out/bench_natural.json 2> gpurun_out/bench_natural.err; tail -c 3000 gpurun_out/bench_natural.json; tail -5 gpurun_out/bench_natural.err
