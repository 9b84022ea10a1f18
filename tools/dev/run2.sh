python -m pytest tests -m gpu -q --maxfail=8 -s 2>&1 | grep -v "^tests/\|^$" | tail -60 > gpurun_out/r02_gpu_tests_b.log
tail -40 gpurun_out/r02_gpu_tests_b.log
