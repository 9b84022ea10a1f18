timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "gemm" 2>&1 | tail -3
echo "== ahead (EPI 3)"; python tools/dev/gemm_resid_ab.py
echo "== off"; CTCLIP_GEMM_RESID_AHEAD=0 python tools/dev/gemm_resid_ab.py
