#!/bin/bash
# round 2, call B: full GPU parity suite (no early stop), cfg3 bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s --maxfail=10 2>&1 | tail -150 > gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
python bench.py --workload cfg3_elephant_N5000_nvt --steps 20 --warmup 3 --cpu-seconds 4 > gpurun_out/r2b_bench_cfg3.json 2> gpurun_out/r2b_bench_cfg3.err
tail -c 300 gpurun_out/r2b_bench_cfg3.json; tail -3 gpurun_out/r2b_bench_cfg3.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_cfg5.json 2> gpurun_out/r2b_bench_cfg5.err
tail -c 300 gpurun_out/r2b_bench_cfg5.json
