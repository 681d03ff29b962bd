timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k two_gpus 2>&1 | tail -15 > gpurun_out/pytest_4gpu_r1s.log
CSS_P2P=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_r1s_4gpu_p2p1.json 2> gpurun_out/bench_r1s_4gpu_p2p1.err
cat gpurun_out/pytest_4gpu_r1s.log; tail -3 gpurun_out/bench_r1s_4gpu_p2p1.err; python - <<'PY'
import json
for p in (1,):
    try:
        d=json.loads(open('gpurun_out/bench_r1s_4gpu_p2p%d.json'%p).read().strip().splitlines()[-1])
        print(p, d['value'], d['ms_per_step'], d['phase_ms_per_step'], d['config']['exchange'], d['e2e']['value'])
    except Exception as e: print(p, 'ERR', e)
PY
