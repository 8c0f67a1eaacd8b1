#!/bin/bash
# gpurun --gpus 2: cfg-5 row-band sharding on two GPUs, then on one GPU of the same box.
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/bench_rowband.py --steps 10 --warmup 3 > gpurun_out/rowband_n2.json 2> gpurun_out/rowband_n2.err; cat gpurun_out/rowband_n2.json; tail -3 gpurun_out/rowband_n2.err
python scripts/bench_rowband.py --steps 10 --warmup 3 > gpurun_out/rowband_n1.json 2> gpurun_out/rowband_n1.err; cat gpurun_out/rowband_n1.json; tail -3 gpurun_out/rowband_n1.err
