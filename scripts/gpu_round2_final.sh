#!/bin/bash
# round 2, final evidence run (1 GPU): parity suite, bench lines of every workload and of the reference arm, launch list,
# full ncu captures of the kernels that make up the step (+ the long-range CTA kernel on the default-executable shape)
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
bash scripts/ncu_kernel.sh k_geodesic_cta default_exe_torus_isotropic_N20 gpurun_out/r2_k_geodesic_cta 4
bash scripts/ncu_stalls.sh k_windows_half > gpurun_out/r2_k_windows_half_stalls.txt 2>&1
SEL='(test_neighbours_distances_tangents_forces and (icosphere16 or torus60x24)) or test_open_mesh_boundary_rules or test_edge_cases or test_stride_guard or test_gpu_vertex_crossings or test_gpu_boundary_vertex or default_executable'
timeout 1500 compute-sanitizer --tool initcheck python -m pytest tests/test_gpu_parity.py tests/test_vertex_crossings.py tests/test_real_meshes.py -m gpu -q -x -k "$SEL" > gpurun_out/r2_sanitizer_initcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_sanitizer_initcheck.log | tail -2
tail -3 gpurun_out/r2_bench.err
nproc; lscpu | grep "Model name"
