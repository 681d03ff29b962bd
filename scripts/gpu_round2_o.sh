#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -6 > gpurun_out/r2o_pytest.log
tail -3 gpurun_out/r2o_pytest.log
python bench.py --workload default_exe_torus_isotropic_N20 --steps 20 --warmup 3 --cpu-seconds 3 > gpurun_out/r2o_bench_default_exe.json 2> gpurun_out/r2o_bench.err
python bench.py --workload cfg1_sphere_radius1_N100 --steps 20 --warmup 3 --cpu-seconds 3 > gpurun_out/r2o_bench_cfg1.json 2>> gpurun_out/r2o_bench.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench_cfg5.json 2>> gpurun_out/r2o_bench.err
python -c "
import json
for f in ('r2o_bench_default_exe','r2o_bench_cfg1','r2o_bench_cfg5'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, '%.4e'%d['value'], d['ms_per_step'], '%.3e'%d['e2e']['value'], d['gpu_launches'], {k:round(v,4) for k,v in d['phase_ms_per_step'].items()}, d['cpu_baseline'] and d['cpu_baseline']['value'])
"
tail -3 gpurun_out/r2o_bench.err
