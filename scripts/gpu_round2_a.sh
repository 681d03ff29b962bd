#!/bin/bash
# round 2, call A: full GPU parity suite, bench lines of every workload, launch list of the default bench
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench_cfg5.json 2> gpurun_out/r2a_bench_cfg5.err
tail -c 600 gpurun_out/r2a_bench_cfg5.json
for w in cfg1_sphere_radius1_N100 cfg2_torusrb20_N2000_gaussian cfg3_elephant_N5000_nvt cfg4_icosphere_250kfaces_N25k; do
  python bench.py --workload $w --steps 20 --warmup 3 --cpu-seconds 4 > gpurun_out/r2a_bench_$w.json 2> gpurun_out/r2a_bench_$w.err
  python bench.py --impl reference --workload $w --steps 5 --warmup 1 > gpurun_out/r2a_ref_$w.json 2>> gpurun_out/r2a_bench_$w.err
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_ref_cfg5.json 2> gpurun_out/r2a_ref_cfg5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches_cfg5.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu_bench.log 2>&1
nproc; lscpu | grep "Model name"
