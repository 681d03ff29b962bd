#!/bin/bash
mkdir -p gpurun_out
for w in cfg5_torus_1Mfaces_N100k cfg4_icosphere_250kfaces_N25k; do
  python scripts/ab_patch.py $w default curvedspacesim_b200/libvariant_minb7.so curvedspacesim_b200/libvariant_minb6.so
done 2>&1 | grep -v "^\[css\]" | tee gpurun_out/r2q_ab.log
bash scripts/ncu_quick.sh k_patch_stencil cfg5_torus_1Mfaces_N100k 4 2>&1 | tee gpurun_out/r2q_ncu_quick.txt
