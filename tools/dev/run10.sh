run() { echo "== $1"; env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 tools/dp_phases.py 2>/dev/null | tail -1; }
run default "X=1" 29601 | tee gpurun_out/r02_dp_ab_2gpu.log
run "NCCL_MAX_CTAS=4" "NCCL_MAX_CTAS=4" 29602 | tee -a gpurun_out/r02_dp_ab_2gpu.log
run "NCCL_MAX_CTAS=4 CTCLIP_SM_BUDGET=144" "NCCL_MAX_CTAS=4 CTCLIP_SM_BUDGET=144" 29603 | tee -a gpurun_out/r02_dp_ab_2gpu.log
run "CTCLIP_SM_BUDGET=140" "CTCLIP_SM_BUDGET=140" 29604 | tee -a gpurun_out/r02_dp_ab_2gpu.log
run "NCCL_MAX_CTAS=8 CTCLIP_SM_BUDGET=140" "NCCL_MAX_CTAS=8 CTCLIP_SM_BUDGET=140" 29605 | tee -a gpurun_out/r02_dp_ab_2gpu.log
run "NCCL_MAX_CTAS=2" "NCCL_MAX_CTAS=2" 29606 | tee -a gpurun_out/r02_dp_ab_2gpu.log
