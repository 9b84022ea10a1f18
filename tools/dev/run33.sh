python -m pytest tests/test_gpu_multi.py tests/test_gpu_peer_loss.py -m gpu -q -s 2>&1 | grep -v "^$" | grep "DP_OK\|passed\|failed\|Error\|error\|skipped" | cut -c1-500 > gpurun_out/r02_gpu_multi_2gpu_b.log; tail -4 gpurun_out/r02_gpu_multi_2gpu_b.log | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras > gpurun_out/r02_bench_2gpu_c.log 2> gpurun_out/r02_bench_2gpu_c.err; python - <<'PY'
import json
for line in open('gpurun_out/r02_bench_2gpu_c.log'):
    if line.startswith('{'):
        d=json.loads(line)
        print(d['n_gpus'], 'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1), 'ms', round(d['e2e']['ms_per_step'],2))
PY
