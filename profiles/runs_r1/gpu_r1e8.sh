#!/bin/bash
# gpurun (r1e, 8th): ncu --set full of the tile kernel (cfg-2, linear; x3.5 Gaussian) and the record-based warp kernels (cfg-4).
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"tile_kernel|warp_fast|warp_records" -s 30 -c 8 -f -o gpurun_out/prof_r1e_tile_warp python scripts/bench_configs.py cfg2 tile cfg4 > gpurun_out/ncu_tile_warp.log 2>&1; tail -2 gpurun_out/ncu_tile_warp.log
