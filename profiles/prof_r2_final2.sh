#!/bin/bash
# ncu --set full of the filter and score kernels (the first capture's skip count missed them) + decode at nc=80 after the copy fast path
mkdir -p gpurun_out
python tools/prof_detect.py 0.5 1 > gpurun_out/r2_prof2_plain1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'filter_|graph_score' -s 4 -c 2 -o gpurun_out/r2_final_front python tools/prof_detect.py 0.5 1 > gpurun_out/r2_prof2_ncu1.log 2>&1
python tools/prof_detect.py 0.001 80 > gpurun_out/r2_prof2_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'filter_|graph_score' -s 4 -c 2 -o gpurun_out/r2_final_front_nc80 python tools/prof_detect.py 0.001 80 > gpurun_out/r2_prof2_ncu2.log 2>&1
python tools/prof_loss.py 80 > gpurun_out/r2_prof2_plain3.log 2>&1 && \
ncu --set full --clock-control none -k regex:'decode_' -s 12 -c 6 -o gpurun_out/r2_final_decode_nc80 python tools/prof_loss.py 80 > gpurun_out/r2_prof2_ncu3.log 2>&1
ls -la gpurun_out/r2_final_front*.ncu-rep gpurun_out/r2_final_decode*.ncu-rep
