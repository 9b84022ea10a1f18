timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/r02_gpu_tests_h.log; cat gpurun_out/r02_gpu_tests_h.log
timeout 600 python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/r02_breakdown_h.json > gpurun_out/r02_bench_h.log 2>gpurun_out/r02_bench_h.err; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench_h.log | head -2
