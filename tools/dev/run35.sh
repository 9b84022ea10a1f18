timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/r02_gpu_tests_final.log; cat gpurun_out/r02_gpu_tests_final.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r02_smoke_final.log
timeout 900 python bench.py --breakdown gpurun_out/r02_breakdown_final.json > gpurun_out/r02_bench_final.log 2> gpurun_out/r02_bench_final.err; python - <<'PY'
import json
for line in open('gpurun_out/r02_bench_final.log'):
    if line.startswith('{'):
        d=json.loads(line)
        r=d['roofline']
        print('value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'steps',d['steps'],'e2e',round(d['e2e']['value'],1), 'roofline', round(r['achieved'],1), round(r['frac'],3), r.get('dominant'), 'launches', d['gpu_launches'], 'cpu', d['cpu_baseline'], 'prep', {k:d['prep'][k] for k in ('ms','frac')} if d.get('prep') else None, 'zs', d.get('zero_shot',{}).get('volumes_per_s'))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-600 | tee gpurun_out/r02_bench_reference_arm.log
