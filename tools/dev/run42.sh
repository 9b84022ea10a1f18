timeout 170 python bench.py --breakdown gpurun_out/r02_breakdown_final.json > gpurun_out/r02_bench_final.log 2> gpurun_out/r02_bench_final.err; python - <<'PY'
import json
for line in open('gpurun_out/r02_bench_final.log'):
    if line.startswith('{'):
        d=json.loads(line); e=d['e2e']
        print('value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',e['mode'],round(e['value'],1),'int16',round(e['raw_int16']['value'],1),'fp32',round(e['fp32_volumes']['value'],1),'roofline',round(d['roofline']['frac'],3))
PY
