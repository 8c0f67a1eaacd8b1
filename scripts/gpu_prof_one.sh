#!/bin/bash
# gpurun: ncu --set full of kernels matching KREGEX while kbench runs its variants on one natural 2K frame.
mkdir -p gpurun_out
KB_FRAMES=${KB_FRAMES:-1} KB_ONLY=${KB_ONLY:-} KB_REP=1 KB_NOTIME=1 KB_KINDS=${KB_KINDS:-natural} ncu --set full --clock-control none --import-source on -k regex:"${KREGEX}" -c ${NCAP:-20} -f -o gpurun_out/${OUT:-prof_one} python scripts/kbench.py > gpurun_out/prof_one.log 2>&1
tail -2 gpurun_out/prof_one.log
