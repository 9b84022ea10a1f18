nvidia-smi topo -m 2>&1 | head -12; lscpu | grep -i "numa\|socket\|^CPU(s)"; free -g | head -2
echo "== bench N=8 (BASELINE configs[2]: global batch 64, 8 per rank) incl. prep + zero-shot (configs[3], [4]) legs"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_8gpu.log 2> gpurun_out/r02_bench_8gpu.err; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/r02_bench_8gpu.err | tail -3
python - <<'PY'
import json
for f in ['gpurun_out/r02_bench_8gpu.log']:
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, d['n_gpus'], d['config']['per_rank_batch'], 'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1), d['e2e'].get('mode'), 'ms', round(d['e2e']['ms_per_step'],2), 'h2d GB/s', round(d['e2e']['h2d_GBps_slowest_rank'],1), 'fp32 e2e', d['e2e'].get('fp32_volumes'), 'zs', d.get('zero_shot'), 'prep', d.get('prep',{}).get('ms'), d['config']['host_numa'])
PY
echo "== phases N=8"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/dp_phases.py 2>/dev/null | tail -1 | tee gpurun_out/r02_dp_phases_8gpu.log
echo "== zero-shot config 5 on 8 GPUs"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 tools/zero_shot_eval.py 2>/dev/null | tail -1 | tee gpurun_out/r02_zero_shot_8gpu.log
