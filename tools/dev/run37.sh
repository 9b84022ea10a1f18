python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 --no-extras > gpurun_out/r02_bench_4gpu_c.log 2> gpurun_out/r02_bench_4gpu_c.err; python - <<'PY'
import json
for line in open('gpurun_out/r02_bench_4gpu_c.log'):
    if line.startswith('{'):
        d=json.loads(line)
        print(d['n_gpus'], 'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1), 'ms', round(d['e2e']['ms_per_step'],2), 'h2d', round(d['e2e']['h2d_GBps_slowest_rank'],1))
PY
