#!/bin/bash
# Run on the GPU box via gpurun: bench (natural + uniform inputs), ncu launch list, ncu --set full of the 3 kernels.
mkdir -p gpurun_out
set -x
python bench.py > gpurun_out/bench_r1_natural.json 2> gpurun_out/bench_r1_natural.err
tail -c 3000 gpurun_out/bench_r1_natural.json
python bench.py --input uniform --no-cpu-baseline > gpurun_out/bench_r1_uniform.json 2> gpurun_out/bench_r1_uniform.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --frames 2 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lut_stage_kernel|resize_sr" -s 9 -c 3 -o gpurun_out/prof_r1 python bench.py --steps 1 --warmup 3 --frames 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
