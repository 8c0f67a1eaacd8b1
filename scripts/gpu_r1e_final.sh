#!/bin/bash
# gpurun (r1e, final pass): all GPU tests, smoke, per-config table (incl. fixed-kernel warps), default bench.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -2
python scripts/bench_configs.py cfg1 cfg2 cfg3 cfg4 fixed cfg5 > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; cat gpurun_out/configs.jsonl | cut -c1-330; tail -3 gpurun_out/configs.err
python bench.py > gpurun_out/bench_natural.json 2> gpurun_out/bench_natural.err; cut -c1-600 gpurun_out/bench_natural.json; tail -3 gpurun_out/bench_natural.err
