#!/bin/bash
mkdir -p gpurun_out
./tools/ubench_pipes > gpurun_out/r2b_ubench.txt 2>&1
cat gpurun_out/r2b_ubench.txt
python tools/diag_conf.py 0.5 0.25 0.001 > gpurun_out/r2b_diag.txt 2>&1
cat gpurun_out/r2b_diag.txt
