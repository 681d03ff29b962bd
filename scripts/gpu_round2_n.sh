#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_real_meshes.py -m gpu -q -s -k "default_executable" 2>&1 | tail -8 > gpurun_out/r2n_pytest.log
tail -4 gpurun_out/r2n_pytest.log
python bench.py --workload default_exe_torus_isotropic_N20 --steps 20 --warmup 3 --cpu-seconds 3 > gpurun_out/r2n_bench_default_exe.json 2> gpurun_out/r2n_bench.err
tail -c 1500 gpurun_out/r2n_bench_default_exe.json; tail -3 gpurun_out/r2n_bench.err
