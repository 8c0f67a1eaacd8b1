#!/bin/bash
# gpurun (r1e): GPU parity tests, smoke, L1 gather micro-benchmark, kernel micro-bench (all variants), accuracy probe of
# the resize variants, default bench.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40
python __graft_entry__.py smoke 2>&1 | tail -3
./build/l1gather > gpurun_out/l1gather.csv 2>&1; cat gpurun_out/l1gather.csv
python scripts/kbench.py > gpurun_out/kbench.log 2>&1; cat gpurun_out/kbench.log
for v in 0 3 4; do echo "variant $v"; ACC_VARIANT=$v python scripts/accuracy_probe.py 600 2>&1 | grep -v generic | tee -a gpurun_out/accuracy.log; done
python bench.py > gpurun_out/bench_natural.json 2> gpurun_out/bench_natural.err; tail -c 3000 gpurun_out/bench_natural.json; tail -5 gpurun_out/bench_natural.err
