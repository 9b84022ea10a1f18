timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "gemm" 2>&1 | tail -8
echo "== TMA store on"; timeout 200 python tools/gemm_vs_cublas.py 2>&1 | tail -24
echo "== TMA store off"; CTCLIP_GEMM_TMA_STORE=0 timeout 200 python tools/gemm_vs_cublas.py 2>&1 | tail -24
