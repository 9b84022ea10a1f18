python tools/step_traffic.py > gpurun_out/step_plain.log 2>&1 || { tail -5 gpurun_out/step_plain.log; exit 1; }
ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r02_step_traffic.csv python tools/step_traffic.py > gpurun_out/ncu_step.log 2>&1
tail -1 gpurun_out/ncu_step.log; wc -l gpurun_out/r02_step_traffic.csv
