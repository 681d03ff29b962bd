#!/bin/bash
# round 2, call C: k_patch v2 against v1 (timing, equality), parity suite on v2
mkdir -p gpurun_out
for w in cfg5_torus_1Mfaces_N100k cfg4_icosphere_250kfaces_N25k cfg3_elephant_N5000_nvt cfg1_sphere_radius1_N100; do
  python scripts/ab_patch.py $w curvedspacesim_b200/libvariant_patchv1.so default
done 2>&1 | tee gpurun_out/r2c_ab.log
python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -60 > gpurun_out/r2c_pytest.log
tail -5 gpurun_out/r2c_pytest.log
