#!/bin/bash
mkdir -p gpurun_out
python tools/prof_detect.py 0.5 > gpurun_out/r2d_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:graph_edge -s 2 -c 1 -o gpurun_out/r2d_edge python tools/prof_detect.py 0.5 > gpurun_out/r2d_ncu.log 2>&1
tail -3 gpurun_out/r2d_plain.log gpurun_out/r2d_ncu.log; ls -la gpurun_out/*.ncu-rep
