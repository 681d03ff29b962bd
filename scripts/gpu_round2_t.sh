#!/bin/bash
# 2 GPUs: bitwise multi-GPU test (incl. the dense stride-guard phase), config-5 bench with the peer-memory exchange
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k two_gpus 2>&1 | tail -15 | tee gpurun_out/r2t_pytest_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2t_bench_cfg5_2gpu_peer.json 2> gpurun_out/r2t_bench_cfg5_2gpu_peer.err
tail -c 300 gpurun_out/r2t_bench_cfg5_2gpu_peer.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2t_bench_cfg5_2gpu_peer.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['phase_ms_per_step'], d['e2e']['value'])
PY
