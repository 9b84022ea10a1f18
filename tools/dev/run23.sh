timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r02_gpu_tests_e.log; tail -3 gpurun_out/r02_gpu_tests_e.log
echo "== fuse geglu on (default)"; timeout 600 python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/r02_breakdown_e.json > gpurun_out/r02_bench_e.log 2>gpurun_out/r02_bench_e.err; python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_e.log",):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); print(f, d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"])
PY
echo "== fuse geglu off"; CTCLIP_FUSE_GEGLU=0 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_e_nofuse.log 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench_e_nofuse.log | head -2
