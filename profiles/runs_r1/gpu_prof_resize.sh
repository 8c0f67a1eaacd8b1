#!/bin/bash
# gpurun: time the resize tuning variants, then ncu --set full of them on one natural 2K frame.
mkdir -p gpurun_out
KB_KINDS=natural python scripts/kbench.py 2>&1 | grep resize
KB_FRAMES=1 KB_REP=1 KB_NOTIME=1 KB_KINDS=natural ncu --set full --clock-control none --import-source on -k regex:"resize_sr_int" -c 24 -f -o gpurun_out/prof_resize python scripts/kbench.py > gpurun_out/prof_resize.log 2>&1
tail -2 gpurun_out/prof_resize.log
