#!/bin/bash
# quick ncu pass with the per-issue stall breakdown of one kernel launch
# usage: [CSS_LIB_PATH=...] scripts/ncu_stalls.sh <kernel regex> [workload] [skip]
K=${1:-k_windows_half}; WL=${2:-cfg5_torus_1Mfaces_N100k}; SKIP=${3:-4}
M=smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active
for r in long_scoreboard short_scoreboard wait no_instruction branch_resolving barrier math_pipe_throttle lg_throttle mio_throttle dispatch_stall not_selected selected imc_miss membar sleeping drain tex_throttle; do
  M=$M,smsp__average_warps_issue_stalled_${r}_per_issue_active.ratio
done
ncu --metrics $M --clock-control none -k regex:$K -s $SKIP -c 1 python scripts/perf_probe.py $WL 2>&1 | grep -E "inst_executed|time_duration|issue_active|warps_active|stalled" | sed -E 's/smsp__average_warps_issue_stalled_//; s/_per_issue_active.ratio//' | awk '{print $1, $NF}'
