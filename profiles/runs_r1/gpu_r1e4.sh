#!/bin/bash
# gpurun (r1e, 4th): tests, bench (natural + uniform), reference arm (short), ncu launch list + --set full of the production kernels.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
TAG=r1e bash scripts/gpu_bench_profile.sh
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1e_reference.json 2>gpurun_out/bench_r1e_reference.err; cat gpurun_out/bench_r1e_reference.json
