python -m pytest tests/test_gpu_ops.py -m gpu -q -k "peg" 2>&1 | tail -3
python tools/time_peg.py 8 2>&1 | tee gpurun_out/r02_time_peg.log
