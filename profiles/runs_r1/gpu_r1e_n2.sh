#!/bin/bash
# gpurun --gpus 2: the driver's multi-GPU launch line for both arms.
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_r1e_n2.json 2> gpurun_out/bench_r1e_n2.err; cat gpurun_out/bench_r1e_n2.json; tail -5 gpurun_out/bench_r1e_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_r1e_n2_reference.json 2> gpurun_out/bench_r1e_n2_reference.err; cat gpurun_out/bench_r1e_n2_reference.json; tail -3 gpurun_out/bench_r1e_n2_reference.err
