python -m pytest tests/test_gpu_ops.py -m gpu -q -k "resample or sumsq or loader or adam" 2>&1 | tail -15 > gpurun_out/r02_gpu_tests_prep.log; cat gpurun_out/r02_gpu_tests_prep.log
python tools/time_prep.py 32 2>&1 | tee gpurun_out/r02_time_prep_a.log
