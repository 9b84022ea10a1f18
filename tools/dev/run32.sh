echo "== bench N=8 (BASELINE configs[2]: global batch 64, 8 per rank)"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_8gpu_b.log 2> gpurun_out/r02_bench_8gpu_b.err; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/r02_bench_8gpu_b.err | tail -3
python - <<'PY'
import json
for line in open('gpurun_out/r02_bench_8gpu_b.log'):
    if line.startswith('{'):
        d=json.loads(line)
        print(d['n_gpus'], 'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1), 'ms', round(d['e2e']['ms_per_step'],2), 'h2d GB/s', round(d['e2e']['h2d_GBps_slowest_rank'],1), 'zs', d.get('zero_shot'), 'prep', d.get('prep'))
PY
echo "== phases N=8"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/dp_phases.py 2>/dev/null | tail -1 | tee gpurun_out/r02_dp_phases_8gpu_b.log
