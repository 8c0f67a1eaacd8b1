#!/bin/bash
# r2d: product library through the GPU suite; the experiments library through the variant tests and kbench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
LERF_B200_EXPERIMENTS=1 python -m pytest tests -m gpu -x -q -k "variants or pipeline_kernel or swizzle" 2>&1 | tail -3
python scripts/pw_check.py 2>&1 | tail -3
KB_FRAMES=8 KB_KINDS=natural python scripts/kbench.py > gpurun_out/r2d_kbench.log 2>&1; grep -v "max-tap v\|row-major\|mix" gpurun_out/r2d_kbench.log | tail -40
