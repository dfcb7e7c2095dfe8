#!/bin/bash
mkdir -p gpurun_out
python tools/prof_detect.py 0.5 > gpurun_out/r2e_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'filter_onepass|graph_sort|graph_gather|graph_resolve' -s 8 -c 4 -o gpurun_out/r2e_front python tools/prof_detect.py 0.5 > gpurun_out/r2e_ncu.log 2>&1
tail -n 3 gpurun_out/r2e_plain.log gpurun_out/r2e_ncu.log; ls -la gpurun_out/*.ncu-rep
