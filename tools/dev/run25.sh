timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_blocks.py -x -q -m gpu 2>&1 | tail -3
python tools/time_ff.py
echo "== fused (default)"; timeout 600 python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/r02_breakdown_g.json > gpurun_out/r02_bench_g.log 2>gpurun_out/r02_bench_g.err; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench_g.log | head -2
