#!/bin/bash
# round 2, call H: stencil stage 1 with byte flags / guarded rounds / 160-face stencils; register-cap variants
mkdir -p gpurun_out
for w in cfg5_torus_1Mfaces_N100k cfg4_icosphere_250kfaces_N25k; do
  python scripts/ab_patch.py $w default:CSS_STENCIL=0 default curvedspacesim_b200/libvariant_minb10.so curvedspacesim_b200/libvariant_minb6.so
done 2>&1 | tee gpurun_out/r2h_ab.log
python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -40 > gpurun_out/r2h_pytest.log
tail -3 gpurun_out/r2h_pytest.log
bash scripts/ncu_quick.sh k_patch_stencil cfg5_torus_1Mfaces_N100k 4 2>&1 | tee gpurun_out/r2h_ncu_quick_stencil.txt
