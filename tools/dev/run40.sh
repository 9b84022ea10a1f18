python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-extras > gpurun_out/r02_bench_8gpu_c.log 2> gpurun_out/r02_bench_8gpu_c.err; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/r02_bench_8gpu_c.err | tail -3
python - <<'PY'
import json
for line in open('gpurun_out/r02_bench_8gpu_c.log'):
    if line.startswith('{'):
        d=json.loads(line); e=d['e2e']
        print('value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'| e2e', e['mode'], round(e['value'],1), round(e['ms_per_step'],2), round(e['h2d_GBps_slowest_rank'],1), '| int16', round(e['raw_int16']['value'],1), round(e['raw_int16']['h2d_GBps_slowest_rank'],1), '| fp32', round(e['fp32_volumes']['value'],1))
PY
