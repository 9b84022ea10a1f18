timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k geglu 2>&1 | tail -3
python tools/time_ff.py
