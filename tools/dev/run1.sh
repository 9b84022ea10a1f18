set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_gpu_tests_a.log
cat gpurun_out/r02_gpu_tests_a.log
python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/r02_breakdown_a.json > gpurun_out/r02_bench_a.log 2>gpurun_out/r02_bench_a.err; tail -c 3000 gpurun_out/r02_bench_a.log
timeout 600 python tools/stock_torch_step.py --batch 8 --steps 3 --warmup 1 > gpurun_out/r02_stock_torch.log 2>&1; tail -5 gpurun_out/r02_stock_torch.log
