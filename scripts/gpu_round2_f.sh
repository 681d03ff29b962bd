#!/bin/bash
# round 2, call F: stage 1 through static face stencils against the flood fill
mkdir -p gpurun_out
for w in cfg5_torus_1Mfaces_N100k cfg4_icosphere_250kfaces_N25k cfg3_elephant_N5000_nvt cfg2_torusrb20_N2000_gaussian cfg1_sphere_radius1_N100; do
  python scripts/ab_patch.py $w default:CSS_STENCIL=0 default
done 2>&1 | tee gpurun_out/r2f_ab.log
python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -40 > gpurun_out/r2f_pytest.log
tail -5 gpurun_out/r2f_pytest.log
bash scripts/ncu_quick.sh k_patch_stencil cfg5_torus_1Mfaces_N100k 4 2>&1 | tee gpurun_out/r2f_ncu_quick_stencil.txt
