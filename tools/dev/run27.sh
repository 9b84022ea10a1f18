timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "layernorm or ln" 2>&1 | tail -3
echo "== ring"; python tools/time_ln.py
echo "== old"; CTCLIP_LN_BWD_RING=0 python tools/time_ln.py
