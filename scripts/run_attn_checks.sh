#!/bin/bash
mkdir -p gpurun_out
for g in rowops attn64 attn32 attnperf; do
  timeout 240 python scripts/check_attn_rows.py $g 2>&1 | tee gpurun_out/check_$g.log | tail -60
done
