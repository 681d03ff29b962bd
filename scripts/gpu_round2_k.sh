#!/bin/bash
# round 2, call K: captures of the two stage kernels (current build), bench lines
mkdir -p gpurun_out
bash scripts/ncu_kernel.sh k_windows_half cfg5_torus_1Mfaces_N100k gpurun_out/r2k_windows_half 4
bash scripts/ncu_kernel.sh k_patch_stencil cfg5_torus_1Mfaces_N100k gpurun_out/r2k_patch_stencil 4
python bench.py --steps 20 --warmup 3 > gpurun_out/r2k_bench_cfg5.json 2> gpurun_out/r2k_bench_cfg5.err
tail -c 400 gpurun_out/r2k_bench_cfg5.json
