#!/bin/bash
# r2a: paired-window stage kernels -- parity against the production kernels, then per-kernel timing
mkdir -p gpurun_out
python scripts/pw_check.py 2>&1 | tail -4
KB_FRAMES=8 KB_ONLY=pw python scripts/kbench.py > gpurun_out/r2a_kbench.log 2>&1; cat gpurun_out/r2a_kbench.log
