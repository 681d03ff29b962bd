#!/bin/bash
# block-cooperative long-range tiers: targeted tests, A/B bench of the default-executable shape, full GPU suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_real_meshes.py -m gpu -x -q -s -k "default_executable or self_consistency_on_the_gpu" 2>&1 | tail -15 | tee gpurun_out/r2r_targeted.log
for cta in 1 0; do
  CSS_CTA=$cta timeout 600 python bench.py --workload default_exe_torus_isotropic_N20 --steps 20 --warmup 3 > gpurun_out/r2r_bench_default_exe_cta$cta.json 2> gpurun_out/r2r_bench_default_exe_cta$cta.err
  tail -c 600 gpurun_out/r2r_bench_default_exe_cta$cta.json | head -c 400; echo
done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2r_pytest.log
