#!/bin/bash
# one full ncu capture of the tier-0 geodesic kernel (first launch after warm-up) on a workload
WL=${1:-small_torus_24kfaces_N5k}
OUT=${2:-gpurun_out/geo_prof}
ncu --set full --clock-control none --import-source on -k regex:k_geodesic -s 9 -c 1 -f -o $OUT python scripts/perf_probe.py $WL > gpurun_out/ncu_geo.log 2>&1
tail -3 gpurun_out/ncu_geo.log
