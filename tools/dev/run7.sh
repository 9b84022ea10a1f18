echo "== multi-GPU parity tests (world 2 and 4, one rank per device)"
python -m pytest tests/test_gpu_multi.py -m gpu -q -s 2>&1 | grep -v "^$" | grep "DP_OK\|passed\|failed\|Error\|error" | cut -c1-700 > gpurun_out/r02_gpu_multi_4gpu.log; cat gpurun_out/r02_gpu_multi_4gpu.log
echo "== config 3 literal: global batch 64 on 4 GPUs (16 per rank)"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --batch 16 --steps 6 --warmup 3 --e2e-mode raw --no-extras > gpurun_out/r02_bench_4gpu_b16.log 2> gpurun_out/r02_bench_4gpu_b16.err; tail -c 600 gpurun_out/r02_bench_4gpu_b16.log; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/r02_bench_4gpu_b16.err | tail -3
echo "== phases N=4"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 tools/dp_phases.py 2>/dev/null | tail -1 | tee gpurun_out/r02_dp_phases_4gpu.log
echo "== weak scaling 4 GPUs (8 per rank)"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02_bench_4gpu.log 2> gpurun_out/r02_bench_4gpu.err; python - <<'PY'
import json
for f in ['gpurun_out/r02_bench_4gpu_b16.log','gpurun_out/r02_bench_4gpu.log']:
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, d['n_gpus'], d['config']['per_rank_batch'], 'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1), d['e2e'].get('mode'), 'h2d GB/s', round(d['e2e']['h2d_GBps_slowest_rank'],1), 'fp32 e2e', d['e2e'].get('fp32_volumes',{}).get('value'), 'zs', d.get('zero_shot',{}).get('volumes_per_s'), d.get('zero_shot',{}).get('e2e'))
PY
