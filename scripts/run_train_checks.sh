#!/bin/bash
mkdir -p gpurun_out
for g in attnbwd adapters loss pretrain_tiny pretrain_base; do
  timeout 300 python scripts/check_train.py $g 2>&1 | tee gpurun_out/check_$g.log | tail -70
done
