echo "== memcheck: GEMM epilogue instantiations, LN ring, attention (padded bias table)"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "geglu or ring or test_gemm_epilogues or (attention_fwd_bwd and grid1)" 2>&1 | tail -8 > gpurun_out/r02_sanitizer_memcheck.log; echo "rc=$?" >> gpurun_out/r02_sanitizer_memcheck.log; tail -6 gpurun_out/r02_sanitizer_memcheck.log
echo "== racecheck: LN ring kernel"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "ring" 2>&1 | tail -6 > gpurun_out/r02_sanitizer_racecheck.log; tail -4 gpurun_out/r02_sanitizer_racecheck.log
