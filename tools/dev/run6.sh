export CTCLIP_PREP_V2_OCC=2
for nc in 1 2; do
export CTCLIP_PREP_V2_NC=$nc
python tools/prof_prep.py 8 > gpurun_out/prof_prep_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:prep_hwn_i16_v2 -s 1 -c 1 -f -o gpurun_out/r02_prep_v2_nc$nc python tools/prof_prep.py 8 > gpurun_out/prof_prep_ncu.log 2>&1; tail -1 gpurun_out/prof_prep_ncu.log
done
