#!/bin/bash
# r2: the gather micro-benchmarks asked for in VERDICT r1 item 1 (binaries are built here by nvcc and travel in build/)
mkdir -p gpurun_out
timeout 120 ./build/l1gather2 > gpurun_out/r2_l1gather2.csv 2>&1; cat gpurun_out/r2_l1gather2.csv
for b in 1 4; do timeout 120 ./build/tmagather4 $b >> gpurun_out/r2_tmagather4.csv 2>&1; done; cat gpurun_out/r2_tmagather4.csv
timeout 120 ./build/dsmemgather > gpurun_out/r2_dsmemgather.csv 2>&1; cat gpurun_out/r2_dsmemgather.csv
