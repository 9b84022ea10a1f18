for d in 0 8 9; do echo "dbg=$d"; CTCLIP_GEGLU_DBG=$d python tools/time_ff.py; done
CTCLIP_GEGLU_DBG=8 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "geglu" 2>&1 | tail -3
