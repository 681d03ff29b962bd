#!/bin/bash
# round 2, call L: programmatic dependent launch on/off
mkdir -p gpurun_out
for w in cfg5_torus_1Mfaces_N100k cfg4_icosphere_250kfaces_N25k cfg3_elephant_N5000_nvt; do
  python scripts/ab_patch.py $w default:CSS_PDL=0 default
done 2>&1 | tee gpurun_out/r2l_ab.log
python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -5 > gpurun_out/r2l_pytest.log
tail -3 gpurun_out/r2l_pytest.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench_cfg5.json 2> gpurun_out/r2l_bench_cfg5.err
CSS_PDL=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench_cfg5_nopdl.json 2>> gpurun_out/r2l_bench_cfg5.err
python -c "
import json
for f in ('r2l_bench_cfg5','r2l_bench_cfg5_nopdl'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['ms_per_step'], d['ms_per_step_hot_l2'], d['e2e']['value'], d['gpu_launches'])
"
