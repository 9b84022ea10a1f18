timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "attention" 2>&1 | tail -15
timeout 120 python tools/time_attn.py
