#!/bin/bash
# gpurun (r1e, 6th): GPU tests, per-config times, ncu --set full of the tile / fast-warp kernels on cfg-2 and cfg-4.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scripts/bench_configs.py cfg2 cfg4 > gpurun_out/configs24.jsonl 2> gpurun_out/configs.err; cat gpurun_out/configs24.jsonl; tail -5 gpurun_out/configs.err
ncu --set full --clock-control none --import-source on -k regex:"tile_kernel|warp_fast" -s 4 -c 3 -f -o gpurun_out/prof_r1e_tile python scripts/bench_configs.py cfg2 cfg4 > gpurun_out/ncu_tile.log 2>&1; tail -2 gpurun_out/ncu_tile.log
