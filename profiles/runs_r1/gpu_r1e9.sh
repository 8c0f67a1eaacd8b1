#!/bin/bash
# gpurun (r1e, 9th): ncu --set full of the record-based warp kernels (cfg-4), the LeRF-L cell-owner kernel and both tile kernels.
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"warp_fast|warp_records" -s 40 -c 4 -f -o gpurun_out/prof_r1e_warp2 python scripts/bench_configs.py cfg4 > gpurun_out/ncu_a.log 2>&1; tail -1 gpurun_out/ncu_a.log
ncu --set full --clock-control none --import-source on -k regex:"int_linear|tile_kernel" -s 6 -c 6 -f -o gpurun_out/prof_r1e_lin python scripts/bench_configs.py lin cfg2 > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log
ncu --set full --clock-control none --import-source on -k regex:"tile_kernel" -s 14 -c 2 -f -o gpurun_out/prof_r1e_tileg python scripts/bench_configs.py tile > gpurun_out/ncu_c.log 2>&1; tail -1 gpurun_out/ncu_c.log
