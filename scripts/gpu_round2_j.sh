#!/bin/bash
# round 2, call J: stencil stage 1 with lane-per-item candidates, round-robin sources, next-source prefetch
mkdir -p gpurun_out
for w in cfg5_torus_1Mfaces_N100k cfg4_icosphere_250kfaces_N25k cfg3_elephant_N5000_nvt; do
  python scripts/ab_patch.py $w default:CSS_STENCIL=0 default
done 2>&1 | tee gpurun_out/r2j_ab.log
python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -5 > gpurun_out/r2j_pytest.log
tail -3 gpurun_out/r2j_pytest.log
bash scripts/ncu_quick.sh k_patch_stencil cfg5_torus_1Mfaces_N100k 4 2>&1 | tee gpurun_out/r2j_ncu_quick_stencil.txt
