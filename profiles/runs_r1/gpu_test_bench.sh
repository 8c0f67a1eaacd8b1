#!/bin/bash
# gpurun: GPU parity tests, smoke, then the default bench (natural) and the uniform-input bench.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -45
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py > gpurun_out/bench_natural.json 2> gpurun_out/bench_natural.err; tail -c 2500 gpurun_out/bench_natural.json; tail -5 gpurun_out/bench_natural.err
python bench.py --input uniform --no-cpu-baseline > gpurun_out/bench_uniform.json 2> gpurun_out/bench_uniform.err; tail -c 2500 gpurun_out/bench_uniform.json
