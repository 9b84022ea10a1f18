python -m pytest tests/test_gpu_ops.py -k gemm -x -q -m gpu 2>&1 | tail -3
echo "== ahead=1"; python tools/dev/gemm_resid_ab.py
echo "== ahead=0"; CTCLIP_GEMM_RESID_AHEAD=0 python tools/dev/gemm_resid_ab.py
