#!/bin/bash
# gpurun: ncu --set full of the LUT-stage kernels (all stage-1 variants + stage 2) on one natural 2K frame.
mkdir -p gpurun_out
KB_FRAMES=1 KB_REP=1 ncu --set full --clock-control none --import-source on -k regex:"lut_stage" -c 14 -o gpurun_out/prof_stages python scripts/kbench.py > gpurun_out/prof_stages.log 2>&1
tail -3 gpurun_out/prof_stages.log
ls -la gpurun_out/*.ncu-rep
