#!/bin/bash
# round 2, final evidence run (1 GPU): parity suite, bench lines of every workload and of the reference arm, launch list,
# full ncu captures of the three kernels that make up the step
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s --maxfail=10 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r2_pytest_gpu.log
tail -3 gpurun_out/r2_pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_cfg5_1gpu.json 2> gpurun_out/r2_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_ref_cfg5.json 2>> gpurun_out/r2_bench.err
for w in cfg1_sphere_radius1_N100 cfg2_torusrb20_N2000_gaussian cfg3_elephant_N5000_nvt cfg4_icosphere_250kfaces_N25k default_exe_torus_isotropic_N20; do
  python bench.py --workload $w --steps 20 --warmup 3 --cpu-seconds 4 > gpurun_out/r2_bench_$w.json 2>> gpurun_out/r2_bench.err
  python bench.py --impl reference --workload $w --steps 10 --warmup 1 > gpurun_out/r2_ref_$w.json 2>> gpurun_out/r2_bench.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_cfg5.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
bash scripts/ncu_kernel.sh k_windows_half cfg5_torus_1Mfaces_N100k gpurun_out/r2_k_windows_half 4
bash scripts/ncu_kernel.sh k_patch_stencil cfg5_torus_1Mfaces_N100k gpurun_out/r2_k_patch_stencil 4
bash scripts/ncu_kernel.sh k_walk cfg5_torus_1Mfaces_N100k gpurun_out/r2_k_walk 4
tail -3 gpurun_out/r2_bench.err
nproc; lscpu | grep "Model name"
