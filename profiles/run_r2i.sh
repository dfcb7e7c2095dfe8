#!/bin/bash
mkdir -p gpurun_out
python tools/prof_detect.py 0.9 80 > gpurun_out/r2i_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'filter_onepass' -s 2 -c 1 -o gpurun_out/r2i_filt_sparse python tools/prof_detect.py 0.9 80 > gpurun_out/r2i_ncu.log 2>&1
python tools/prof_detect.py 0.001 80 > gpurun_out/r2i_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'filter_onepass' -s 2 -c 1 -o gpurun_out/r2i_filt_dense python tools/prof_detect.py 0.001 80 > gpurun_out/r2i_ncu2.log 2>&1
tail -n 2 gpurun_out/r2i_ncu.log gpurun_out/r2i_ncu2.log
