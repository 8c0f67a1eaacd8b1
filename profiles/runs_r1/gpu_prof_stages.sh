#!/bin/bash
# gpurun: ncu --set full of the LUT-stage kernels (variants kbench runs) on one natural 2K frame.
# KREGEX selects kernels (default: every lut_stage kernel); -c bounds the captures.
mkdir -p gpurun_out
KREGEX=${KREGEX:-lut_stage}
KB_FRAMES=1 KB_REP=1 KB_NOTIME=1 KB_KINDS=${KB_KINDS:-natural} ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s ${NSKIP:-0} -c ${NCAP:-16} -f -o gpurun_out/${OUT:-prof_stages} python scripts/kbench.py > gpurun_out/prof_stages.log 2>&1
tail -3 gpurun_out/prof_stages.log
ls -la gpurun_out/*.ncu-rep
