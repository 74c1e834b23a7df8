#!/bin/bash
# Issue-slot / pipe metrics of the bounce kernels of one cornell wavefront for several library variants:
#   tools/gpu_ab_metrics.sh TAG variants/a.so variants/b.so ...   -> gpurun_out/TAG_<name>_metrics.csv
set -u
mkdir -p gpurun_out
TAG=$1; shift
ncu --query-metrics 2>/dev/null | grep -iE "inst_executed_pipe|issue_active|inst_issued" | awk '{print $1}' | sort -u > gpurun_out/${TAG}_metric_names.txt
M=smsp__inst_executed.sum,smsp__inst_issued.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,gpu__time_duration.sum
M=$M,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_uniform.sum,sm__inst_executed_pipe_cbu.sum,sm__inst_executed_pipe_adu.sum
M=$M,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active
M=$M,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__warps_active.avg.per_cycle_active
for so in "$@"; do
  name=$(basename $so .so)
  PTB_SO=$so timeout 600 ncu --metrics $M --clock-control none -k regex:bounce_kernel -s 8 -c 3 --csv --log-file gpurun_out/${TAG}_${name}_metrics.csv \
      python tools/gpu_one_batch.py 2 cornell.txt 16 > gpurun_out/${TAG}_${name}_ncu.log 2>&1
  tail -1 gpurun_out/${TAG}_${name}_ncu.log
done
