timeout 300 python -m pytest tests/test_gpu_model.py tests/test_gpu_bert.py -x -q -m gpu 2>&1 | tail -2
timeout 200 python tools/host_starvation.py 2>&1 | tail -1 | tee gpurun_out/r02_host_starvation_c.log
