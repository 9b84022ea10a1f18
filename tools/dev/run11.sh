run() { echo "== $1"; env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 tools/dp_phases.py 2>/dev/null | tail -1; }
run "factor gather on its own communicator (default)" "X=1" 29601 | tee gpurun_out/r02_dp_ab2_2gpu.log
run "CTCLIP_FACTOR_GROUP=0" "CTCLIP_FACTOR_GROUP=0" 29602 | tee -a gpurun_out/r02_dp_ab2_2gpu.log
python -m pytest tests/test_gpu_multi.py -m gpu -q -k "2" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras > gpurun_out/r02_bench_2gpu_b.log 2> gpurun_out/r02_bench_2gpu_b.err; python - <<'PY'
import json
for line in open('gpurun_out/r02_bench_2gpu_b.log'):
    if line.startswith('{'):
        d=json.loads(line)
        print(d['n_gpus'], 'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1), d['e2e'].get('mode'), 'ms', round(d['e2e']['ms_per_step'],2), 'h2d', round(d['e2e']['h2d_GBps_slowest_rank'],1), 'fp32 e2e', d['e2e'].get('fp32_volumes',{}).get('value'))
PY
