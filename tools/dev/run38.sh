python tools/dev/time_ln_bert.py
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_bert.py -x -q -m gpu -k "layernorm or ln or bert" 2>&1 | tail -3
