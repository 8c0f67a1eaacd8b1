#!/bin/bash
# gpurun [--gpus N]: the default bench line (N = 1) or the torchrun line the driver uses (N > 1), into gpurun_out/$OUT
mkdir -p gpurun_out
N=${N:-1}; OUT=${OUT:-bench_n$N}
if [ "$N" = "1" ]; then
  python bench.py --steps ${STEPS:-20} --warmup ${WARMUP:-5} > gpurun_out/$OUT.json 2> gpurun_out/$OUT.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps ${STEPS:-20} --warmup ${WARMUP:-5} > gpurun_out/$OUT.json 2> gpurun_out/$OUT.err
fi
tail -c 6000 gpurun_out/$OUT.json; tail -5 gpurun_out/$OUT.err
