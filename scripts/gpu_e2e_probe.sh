#!/bin/bash
# gpurun --gpus N: the host-to-host path with every rank copying at once (scripts/e2e_probe.py), then the bench line at N
mkdir -p gpurun_out
N=${N:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 scripts/e2e_probe.py > gpurun_out/r2_e2e_probe_n$N.txt 2> gpurun_out/e2e_probe.err; cat gpurun_out/r2_e2e_probe_n$N.txt; tail -3 gpurun_out/e2e_probe.err
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt 2>&1
N=$N OUT=r2_bench_n$N STEPS=20 WARMUP=5 bash scripts/gpu_bench.sh | cut -c1-1500
