#!/bin/bash
# round 2, call I: packed stencil records + prefetch; full capture
mkdir -p gpurun_out
for w in cfg5_torus_1Mfaces_N100k cfg4_icosphere_250kfaces_N25k cfg3_elephant_N5000_nvt; do
  python scripts/ab_patch.py $w default:CSS_STENCIL=0 default
done 2>&1 | tee gpurun_out/r2i_ab.log
python -m pytest tests/test_gpu_parity.py tests/test_real_meshes.py -m gpu -q --maxfail=10 -k "neighbours or full_size or golden or dense or edge_cases or stride" 2>&1 | tail -5 > gpurun_out/r2i_pytest.log
tail -3 gpurun_out/r2i_pytest.log
bash scripts/ncu_kernel.sh k_patch_stencil cfg5_torus_1Mfaces_N100k gpurun_out/r2i_stencil 4
