#!/bin/bash
mkdir -p gpurun_out
python tools/prof_detect.py 0.5 > gpurun_out/r2f_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'graph_spatial|graph_score' -s 4 -c 2 -o gpurun_out/r2f_front python tools/prof_detect.py 0.5 > gpurun_out/r2f_ncu.log 2>&1
tail -n 3 gpurun_out/r2f_plain.log gpurun_out/r2f_ncu.log
