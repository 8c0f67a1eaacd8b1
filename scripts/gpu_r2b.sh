#!/bin/bash
# r2b: paired-window stage 2 as production: parity script, GPU test suite, smoke, default bench
mkdir -p gpurun_out
python scripts/pw_check.py 2>&1 | tail -4
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 2500 gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err
