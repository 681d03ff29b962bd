#!/bin/bash
# round 2, call M (2 GPUs): multi-GPU bitwise test and bench lines
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "two_gpus" -s 2>&1 | tail -15 > gpurun_out/r2m_pytest_2gpu.log
tail -5 gpurun_out/r2m_pytest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2m_bench_cfg5_2gpu.json 2> gpurun_out/r2m_bench_2gpu.err
tail -c 300 gpurun_out/r2m_bench_cfg5_2gpu.json; tail -3 gpurun_out/r2m_bench_2gpu.err
CSS_P2P=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2m_bench_cfg5_2gpu_nccl.json 2>> gpurun_out/r2m_bench_2gpu.err
