#!/bin/bash
# gpurun --gpus 4: host topology, then the N = 4 bench with ranks bound to their GPU's CPUs.
mkdir -p gpurun_out
(nvidia-smi topo -m; nproc; numactl -H 2>/dev/null | head -12; lscpu | grep -i "numa\|socket\|model name") > gpurun_out/topo_n4.txt 2>&1; cat gpurun_out/topo_n4.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_r1e_n4.json 2> gpurun_out/bench_r1e_n4.err; cat gpurun_out/bench_r1e_n4.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['e2e'])"; tail -3 gpurun_out/bench_r1e_n4.err
