#!/bin/bash
# 8 GPUs: config-5 bench lines at 8 and 4 ranks (head of the round)
mkdir -p gpurun_out
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2970$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2p_bench_cfg5_${n}gpu.json 2> gpurun_out/r2p_bench_${n}gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2p_bench_cfg5_${n}gpu.json").read().strip().splitlines()[-1])
print($n, d["value"], d["ms_per_step"], d["ms_per_step_hot_l2"], d["e2e"]["value"], d["phase_ms_per_step"])
PY
done
