#!/bin/bash
# runs every check_gemm group in its own process with a hard timeout
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
for g in basic epilogue dgrad wgrad patch perf; do
  timeout 240 python scripts/check_gemm.py $g 2>&1 | tee gpurun_out/check_gemm_$g.log | tail -40
done
