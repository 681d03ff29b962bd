timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k two_gpus 2>&1 | tail -15 > gpurun_out/pytest_2gpu_r1q.log
for p in 1 0; do
CSS_P2P=$p timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$p bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_r1q_2gpu_p2p$p.json 2> gpurun_out/bench_r1q_2gpu_p2p$p.err
done
cat gpurun_out/pytest_2gpu_r1q.log; tail -3 gpurun_out/bench_r1q_2gpu_p2p1.err; python - <<'PY'
import json
for p in (1,0):
    try:
        d=json.loads(open('gpurun_out/bench_r1q_2gpu_p2p%d.json'%p).read().strip().splitlines()[-1])
        print(p, d['value'], d['ms_per_step'], d['phase_ms_per_step'], d['config']['exchange'], d['e2e']['value'])
    except Exception as e: print(p, 'ERR', e)
PY
