#!/bin/bash
# round 2, call G: stencil stage 1 with flood-fill fallback; full capture of k_patch_stencil
mkdir -p gpurun_out
for w in cfg5_torus_1Mfaces_N100k cfg4_icosphere_250kfaces_N25k; do
  python scripts/ab_patch.py $w default:CSS_STENCIL=0 default
done 2>&1 | tee gpurun_out/r2g_ab.log
python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -40 > gpurun_out/r2g_pytest.log
tail -3 gpurun_out/r2g_pytest.log
bash scripts/ncu_kernel.sh k_patch_stencil cfg5_torus_1Mfaces_N100k gpurun_out/r2g_stencil 4
