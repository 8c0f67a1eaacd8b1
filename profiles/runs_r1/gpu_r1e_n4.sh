#!/bin/bash
# gpurun --gpus 4: the driver's multi-GPU launch line at N = 4 (per-image sharding) + the cfg-5 row-band run.
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_r1e_n4.json 2> gpurun_out/bench_r1e_n4.err; cat gpurun_out/bench_r1e_n4.json; tail -3 gpurun_out/bench_r1e_n4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 scripts/bench_rowband.py --steps 10 --warmup 3 > gpurun_out/rowband_n4.json 2> gpurun_out/rowband_n4.err; cat gpurun_out/rowband_n4.json; tail -3 gpurun_out/rowband_n4.err
