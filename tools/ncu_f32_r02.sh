#!/bin/bash
# ncu --set full of the float32 CVIIRS and simple kernels (4 granules), summaries into gpurun_out/
set -e
for cfg in cviirs_1km_250m simple_1km_250m; do
  CMD="python bench.py --config $cfg --dtype f32 --steps 2 --warmup 1 --granules-per-gpu 4 --no-cpu-baseline --no-e2e --sustained-seconds 0"
  $CMD > gpurun_out/plain_$cfg.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fast_1km_f32 -s 1 -c 1 -o gpurun_out/r02_${cfg}_f32 $CMD > gpurun_out/ncu_$cfg.log 2>&1
done
