python tools/host_starvation.py > gpurun_out/r02_host_starvation.log 2>&1; tail -1 gpurun_out/r02_host_starvation.log
python tools/timeline.py gpurun_out/r02_timeline.json > gpurun_out/r02_timeline.log 2>&1; tail -1 gpurun_out/r02_timeline.log
