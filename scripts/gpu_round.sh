#!/bin/bash
# gpurun: GPU parity tests, smoke, kernel micro-bench (all variants), default bench.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python __graft_entry__.py smoke 2>&1 | tail -3
python scripts/kbench.py > gpurun_out/kbench.log 2>&1; cat gpurun_out/kbench.log
python bench.py > gpurun_out/bench_natural.json 2> gpurun_out/bench_natural.err; tail -c 3000 gpurun_out/bench_natural.json; tail -5 gpurun_out/bench_natural.err
