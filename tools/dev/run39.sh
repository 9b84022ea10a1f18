timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "unpack12" 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r02_bench_i.log 2> gpurun_out/r02_bench_i.err; tail -3 gpurun_out/r02_bench_i.err; python - <<'PY'
import json
for line in open('gpurun_out/r02_bench_i.log'):
    if line.startswith('{'):
        d=json.loads(line); e=d['e2e']
        print('value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'| e2e', e['mode'], round(e['value'],1), round(e['ms_per_step'],2), e['h2d_bytes_per_step'], '| int16', round(e['raw_int16']['value'],1), '| fp32', round(e['fp32_volumes']['value'],1))
PY
