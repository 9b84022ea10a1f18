python -m pytest tests/test_gpu_model.py -m gpu -q -s -k production 2>&1 | grep -v "^$" | tail -12 > gpurun_out/r02_gpu_tests_c.log
cat gpurun_out/r02_gpu_tests_c.log
python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/r02_breakdown_b.json > gpurun_out/r02_bench_b.log 2>gpurun_out/r02_bench_b.err; tail -c 6000 gpurun_out/r02_bench_b.log; tail -5 gpurun_out/r02_bench_b.err
