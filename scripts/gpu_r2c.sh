#!/bin/bash
# r2c: staged uint8 epilogue -- GPU test suite, kernel timings, the fixed L1 gather micro-benchmark
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
KB_FRAMES=8 KB_ONLY=pw KB_KINDS=natural python scripts/kbench.py 2>&1 | grep "resize\|Error\|error\|assert" | tee gpurun_out/r2c_kbench.log
timeout 120 ./build/l1gather2 > gpurun_out/r2_l1gather2.csv 2>&1; cat gpurun_out/r2_l1gather2.csv
