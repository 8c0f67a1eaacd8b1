#!/bin/bash
# gpurun: default bench (natural + uniform inputs), ncu launch list of the same command, ncu --set full of the three
# production kernels at the bench's launch size (8 frames per launch).  TAG names the output files.
TAG=${TAG:-r1c}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${TAG}_natural.json 2> gpurun_out/bench_${TAG}_natural.err
tail -c 2500 gpurun_out/bench_${TAG}_natural.json; tail -3 gpurun_out/bench_${TAG}_natural.err
python bench.py --input uniform --no-cpu-baseline > gpurun_out/bench_${TAG}_uniform.json 2> gpurun_out/bench_${TAG}_uniform.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lut_|resize_sr|warp_kernel|sr_pipeline" -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lut_stage|resize_sr_int" -s 9 -c 3 -f -o gpurun_out/prof_${TAG} python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -8
