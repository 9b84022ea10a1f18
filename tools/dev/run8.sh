python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_gpu_tests_d.log; cat gpurun_out/r02_gpu_tests_d.log
python tools/dp_phases.py 2>/dev/null | tail -1 | tee gpurun_out/r02_dp_phases_1gpu.log
python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/r02_breakdown_c.json > gpurun_out/r02_bench_c.log 2>gpurun_out/r02_bench_c.err; python - <<'PY'
import json
for line in open('gpurun_out/r02_bench_c.log'):
    if line.startswith('{'):
        d=json.loads(line)
        print('value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1), d['e2e'].get('mode'), 'fp32 e2e', d['e2e'].get('fp32_volumes',{}).get('value'), 'frac', round(d['roofline']['frac'],3), 'prep', d['prep']['ms'], d['prep']['frac'], 'zs', d['zero_shot']['volumes_per_s'], d['zero_shot']['e2e']['volumes_per_s'], 'launches', d['gpu_launches_per_step'])
PY
tail -3 gpurun_out/r02_bench_c.err
