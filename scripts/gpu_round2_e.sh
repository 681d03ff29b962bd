#!/bin/bash
# round 2, call E: stage 1 with two sources per warp against one warp per source (v2) and round 1's kernel (v1)
mkdir -p gpurun_out
for w in cfg5_torus_1Mfaces_N100k cfg4_icosphere_250kfaces_N25k cfg3_elephant_N5000_nvt cfg2_torusrb20_N2000_gaussian; do
  python scripts/ab_patch.py $w curvedspacesim_b200/libvariant_patchv1.so:CSS_PATCH_PAIR=0 default:CSS_PATCH_PAIR=0 default
done 2>&1 | tee gpurun_out/r2e_ab.log
python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -40 > gpurun_out/r2e_pytest.log
tail -5 gpurun_out/r2e_pytest.log
bash scripts/ncu_quick.sh k_patch_pair cfg5_torus_1Mfaces_N100k 4 2>&1 | tee gpurun_out/r2e_ncu_quick_pair.txt
CSS_PATCH_PAIR=0 bash scripts/ncu_quick.sh k_patch cfg5_torus_1Mfaces_N100k 4 2>&1 | tee gpurun_out/r2e_ncu_quick_v2.txt
