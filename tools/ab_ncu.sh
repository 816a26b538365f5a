#!/bin/bash
# per-build instruction mix of the headline kernel (2 granules): tools/ab_ncu.sh <lib.so>...
M=smsp__inst_executed.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,sm__warps_active.avg.per_cycle_active
for lib in "$@"; do
  echo "== $lib"
  GTP_B200_LIB=$lib ncu --metrics $M --clock-control none -k regex:modis_fast -c 1 python bench.py $BENCH_ARGS --granules-per-gpu 4 --no-e2e --no-cpu-baseline --steps 1 --warmup 3 --sustained-seconds 0 2>&1 | grep -E "^\s+(smsp|sm__|gpu__)" 
done
