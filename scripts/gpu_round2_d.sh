#!/bin/bash
# round 2, call D: full ncu captures of k_patch (v2) and k_walk on config 5
mkdir -p gpurun_out
bash scripts/ncu_kernel.sh k_patch cfg5_torus_1Mfaces_N100k gpurun_out/r2d_patch 4
bash scripts/ncu_kernel.sh k_walk cfg5_torus_1Mfaces_N100k gpurun_out/r2d_walk 4
ls -la gpurun_out/r2d_*
