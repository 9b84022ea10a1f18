timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -3
python tools/time_ff.py
echo "== fuse geglu on (default)"; timeout 600 python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/r02_breakdown_f.json > gpurun_out/r02_bench_f.log 2>gpurun_out/r02_bench_f.err; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench_f.log | head -2
echo "== fuse geglu off"; CTCLIP_FUSE_GEGLU=0 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_f_nofuse.log 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench_f_nofuse.log | head -2
