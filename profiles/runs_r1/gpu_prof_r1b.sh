#!/bin/bash
# gpurun: ncu --set full of the three production kernels (natural 2K frame, 1 frame) + launch list of the default bench.
mkdir -p gpurun_out
KB_FRAMES=1 KB_REP=1 KB_ONLY=prod KB_NOTIME=1 KB_KINDS=natural ncu --set full --clock-control none --import-source on -k regex:"lut_stage_kernel|resize_sr_int" -c 10 -o gpurun_out/prof_r1b python scripts/kbench.py > gpurun_out/prof_r1b.log 2>&1
tail -3 gpurun_out/prof_r1b.log
ls -la gpurun_out/*.ncu-rep
