echo "== topology"; nvidia-smi topo -m 2>&1 | head -20; ls /sys/devices/system/node 2>&1 | head; lscpu | grep -i "numa\|socket\|^CPU(s)" ; grep -i "Cpus_allowed_list\|Mems_allowed_list" /proc/self/status; free -g | head -2
echo "== multi-GPU parity tests"
python -m pytest tests/test_gpu_multi.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r02_gpu_multi_2gpu.log; cat gpurun_out/r02_gpu_multi_2gpu.log
echo "== config 3 literal: global batch 64 on 2 GPUs (32 per rank)"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --batch 32 --steps 5 --warmup 3 --e2e-mode raw --no-extras > gpurun_out/r02_bench_2gpu_b32.log 2> gpurun_out/r02_bench_2gpu_b32.err; tail -c 2500 gpurun_out/r02_bench_2gpu_b32.log; tail -3 gpurun_out/r02_bench_2gpu_b32.err
echo "== weak scaling 2 GPUs (8 per rank)"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu.log 2> gpurun_out/r02_bench_2gpu.err; tail -c 1500 gpurun_out/r02_bench_2gpu.log | cut -c1-1500; tail -3 gpurun_out/r02_bench_2gpu.err
