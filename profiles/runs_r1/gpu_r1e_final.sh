#!/bin/bash
# gpurun (r1e, final pass): all GPU tests, smoke, per-config table, default bench (natural + uniform), reference arm,
# ncu launch list of the bench command.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -2
python scripts/bench_configs.py cfg1 cfg2 cfg3 cfg4 fixed cfg5 tile lin > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; cut -c1-200 gpurun_out/configs.jsonl; tail -3 gpurun_out/configs.err
python bench.py > gpurun_out/bench_r1e_natural.json 2> gpurun_out/bench_natural.err; cut -c1-300 gpurun_out/bench_r1e_natural.json; tail -3 gpurun_out/bench_natural.err
python bench.py --input uniform --no-cpu-baseline > gpurun_out/bench_r1e_uniform.json 2> gpurun_out/bench_uniform.err; cut -c1-200 gpurun_out/bench_r1e_uniform.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1e_reference.json 2> gpurun_out/bench_reference.err; cut -c1-300 gpurun_out/bench_r1e_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lut_|resize_sr|warp_|sr_pipeline" -c 400 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log | cut -c1-200
