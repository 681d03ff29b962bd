#!/bin/bash
# one full ncu capture of a kernel (first matching launch after warm-up) on a perf_probe workload
# usage: scripts/ncu_kernel.sh <kernel regex> <workload> <out prefix> [skip]
K=${1:-k_windows}
WL=${2:-cfg5_torus_1Mfaces_N100k}
OUT=${3:-gpurun_out/prof_$K}
SKIP=${4:-4}
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o $OUT python scripts/perf_probe.py $WL > $OUT.log 2>&1
tail -2 $OUT.log
