#!/bin/bash
# r2 evidence pass (1 GPU): GPU suite, smoke, per-config table, bench (ours + reference arm), ncu --set full of the three
# production kernels at the bench's launch size (natural and uniform input), ncu launch list of the bench command.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1
python scripts/bench_configs.py cfg1 cfg2 cfg3 cfg4 fixed cfg5 tile lin > gpurun_out/r2_configs.jsonl 2> gpurun_out/r2_configs.err; cut -c1-220 gpurun_out/r2_configs.jsonl; tail -2 gpurun_out/r2_configs.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; cut -c1-400 gpurun_out/bench_r2_n1.json; tail -2 gpurun_out/bench_r2_n1.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err; cut -c1-600 gpurun_out/bench_r2_reference.json
# ncu: the timed step's three kernels (8 frames), last warm instance of each (gpurun merges at most 64 MiB back: keep the reports small)
ncu --set full --clock-control none --import-source on -k regex:"lut_stage_cell_kernel|lut_stage_pw_kernel|resize_sr_int_gauss_kernel" -s 9 -c 3 -f -o gpurun_out/r2_prof \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-extra-arms --no-parity-check > gpurun_out/r2_prof.log 2>&1; tail -1 gpurun_out/r2_prof.log | cut -c1-200
ncu --set full --clock-control none -k regex:"lut_stage_cell_kernel|lut_stage_pw_kernel|resize_sr_int_gauss_kernel" -s 9 -c 3 -f -o gpurun_out/r2_prof_uniform \
  python bench.py --input uniform --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-extra-arms --no-parity-check > gpurun_out/r2_prof_uniform.log 2>&1; tail -1 gpurun_out/r2_prof_uniform.log | cut -c1-200
ncu --set full --clock-control none -k regex:"resize_sr_int_gauss_u8" -c 2 -f -o gpurun_out/r2_prof_u8 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-parity-check > gpurun_out/r2_prof_u8.log 2>&1; tail -1 gpurun_out/r2_prof_u8.log | cut -c1-100
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lut_|resize_sr|warp_" -c 400 --csv --log-file gpurun_out/r2_ncu_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-extra-arms --no-parity-check > gpurun_out/r2_launch.log 2>&1; tail -1 gpurun_out/r2_launch.log | cut -c1-200
# gpurun merges at most 64 MiB back: summarise the captures here and keep only the main report
python scripts/ncu_summary.py gpurun_out/r2_prof.ncu-rep gpurun_out/r2_ncu_summary.md "r2: the three production kernels at the bench's launch size (8 frames 2040x1356, natural-like input)" > /dev/null
python scripts/ncu_summary.py gpurun_out/r2_prof_uniform.ncu-rep gpurun_out/r2_ncu_uniform_summary.md "r2: the same kernels on uniform-random input" > /dev/null
python scripts/ncu_summary.py gpurun_out/r2_prof_u8.ncu-rep gpurun_out/r2_ncu_u8_summary.md "r2: the uint8 epilogue kernels (planar through a lane shuffle, HWC through a staged tile)" > /dev/null
python scripts/ncu_traffic.py gpurun_out/r2_prof.ncu-rep gpurun_out/traffic.json > /dev/null
rm -f gpurun_out/r2_prof_uniform.ncu-rep gpurun_out/r2_prof_u8.ncu-rep
