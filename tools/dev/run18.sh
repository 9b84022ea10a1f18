python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "layernorm or ln" 2>&1 | tail -2
echo "== ahead=1"; python tools/time_ln.py
echo "== ahead=0"; CTCLIP_LN_BWD_L2_AHEAD=0 python tools/time_ln.py
