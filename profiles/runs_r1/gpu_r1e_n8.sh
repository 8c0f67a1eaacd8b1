#!/bin/bash
# gpurun --gpus 8: the driver's launch line at N = 8 (per-image sharding), the reference arm, and the cfg-5 row-band run.
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_r1e_n8.json 2> gpurun_out/bench_r1e_n8.err; cut -c1-200 gpurun_out/bench_r1e_n8.json; tail -2 gpurun_out/bench_r1e_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 scripts/bench_rowband.py --steps 10 --warmup 3 > gpurun_out/rowband_n8.json 2> gpurun_out/rowband_n8.err; cut -c1-200 gpurun_out/rowband_n8.json; tail -2 gpurun_out/rowband_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29553 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 --sample-rows 256 > gpurun_out/bench_r1e_n8_reference.json 2> gpurun_out/bench_r1e_n8_reference.err; cut -c1-300 gpurun_out/bench_r1e_n8_reference.json
