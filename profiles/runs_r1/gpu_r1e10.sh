#!/bin/bash
# gpurun (r1e, 10th): ncu --set full of the two stage kernels on UNIFORM-random frames (the adversarial input of SURVEY 8d).
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"lut_stage" -s 6 -c 2 -f -o gpurun_out/prof_r1e_uniform python bench.py --input uniform --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_u.log 2>&1; tail -1 gpurun_out/ncu_u.log
