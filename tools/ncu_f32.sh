#!/bin/bash
# pipe utilisation of the float32 fast kernels: tools/ncu_f32.sh > gpurun_out/ncu_f32.txt
M=smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__warps_active.avg.per_cycle_active,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
for cfg in simple_1km_250m cviirs_1km_250m; do
  echo "== $cfg f32"
  ncu --metrics $M --clock-control none -k regex:fast_1km_f32 -c 1 python bench.py --config $cfg --dtype f32 --granules-per-gpu 4 --no-e2e --no-cpu-baseline --steps 1 --warmup 3 --sustained-seconds 0 2>&1 | grep -E "^\s+(smsp|sm__|gpu__)"
done
