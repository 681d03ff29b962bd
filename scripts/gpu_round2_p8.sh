#!/bin/bash
# round 2, call P (8 GPUs): multi-GPU bitwise test over all 8 GPUs and the 4- and 8-GPU bench lines
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "two_gpus" -s 2>&1 | grep -v "OMP_NUM\|^\*\*\*\|^$" | tail -6 > gpurun_out/r2p_pytest_8gpu.log
cat gpurun_out/r2p_pytest_8gpu.log
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2970$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2p_bench_cfg5_${n}gpu.json 2> gpurun_out/r2p_bench_${n}gpu.err
tail -c 200 gpurun_out/r2p_bench_cfg5_${n}gpu.json
done
