run() { echo "== $1"; env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 tools/dp_phases.py 2>/dev/null | tail -1; }
run "overlapped reduction (default)" "X=1" 29601 | tee gpurun_out/r02_dp_ab3_2gpu.log
run "CTCLIP_DP_OVERLAP=0" "CTCLIP_DP_OVERLAP=0" 29602 | tee -a gpurun_out/r02_dp_ab3_2gpu.log
run "NCCL_MIN_CTAS=32" "NCCL_MIN_CTAS=32" 29603 | tee -a gpurun_out/r02_dp_ab3_2gpu.log
run "CTCLIP_DP_OVERLAP=0 NCCL_MIN_CTAS=32" "CTCLIP_DP_OVERLAP=0 NCCL_MIN_CTAS=32" 29604 | tee -a gpurun_out/r02_dp_ab3_2gpu.log
