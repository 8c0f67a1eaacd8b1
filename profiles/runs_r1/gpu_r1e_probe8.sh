#!/bin/bash
mkdir -p gpurun_out
(nvidia-smi topo -m | head -14; nproc; lscpu | grep -i "numa node(s)\|socket") > gpurun_out/topo_n8.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 scripts/e2e_probe.py > gpurun_out/e2e_probe_n8.txt 2> gpurun_out/e2e_probe_n8.err; cat gpurun_out/e2e_probe_n8.txt; tail -3 gpurun_out/e2e_probe_n8.err; head -16 gpurun_out/topo_n8.txt
