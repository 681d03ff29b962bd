#!/bin/bash
# quick ncu pass: instruction count, duration, issue utilisation, lanes per instruction of one kernel launch
# usage: scripts/ncu_quick.sh <kernel regex> [workload] [skip]
K=${1:-k_windows_half}; WL=${2:-cfg5_torus_1Mfaces_N100k}; SKIP=${3:-4}
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warp_latency_per_inst_issued.ratio \
  --clock-control none -k regex:$K -s $SKIP -c 1 python scripts/perf_probe.py $WL 2>&1 | grep -E "inst_executed|time_duration|issue_active|warps_active|latency_per" 
