#!/bin/bash
# gpurun (r1e, 7th): ncu --set full of the fast warp kernel (cfg-4) and of the Gaussian tile kernel (x3.5 on a 2K frame).
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"warp_fast" -s 20 -c 2 -f -o gpurun_out/prof_r1e_warp python scripts/bench_configs.py cfg4 > gpurun_out/ncu_warp.log 2>&1; tail -2 gpurun_out/ncu_warp.log
