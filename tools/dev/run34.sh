timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "vq or gemm or top2" 2>&1 | tail -3
python tools/dev/time_top2.py
