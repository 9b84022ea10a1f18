set -x
python tools/prof_kernels.py 8 > gpurun_out/prof_k_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"attn_bwd_kernel|attn_fwd_kernel|layernorm_bwd512|gemm_bf16|prep_hwn" -c 14 -f -o gpurun_out/r02_kernels_b python tools/prof_kernels.py 8 > gpurun_out/prof_k_ncu.log 2>&1
tail -2 gpurun_out/prof_k_ncu.log
python tools/gemm_vs_cublas.py > gpurun_out/r02_gemm_vs_cublas.txt 2>&1; tail -3 gpurun_out/r02_gemm_vs_cublas.txt
python tools/time_attn.py > gpurun_out/r02_time_attn.log 2>&1; python tools/time_ln.py > gpurun_out/r02_time_ln.log 2>&1; python tools/time_ff.py > gpurun_out/r02_time_ff.log 2>&1; python tools/time_peg.py > gpurun_out/r02_time_peg_b.log 2>&1
python tools/host_starvation.py > gpurun_out/r02_host_starvation_b.log 2>&1; tail -1 gpurun_out/r02_host_starvation_b.log
