#!/bin/bash
mkdir -p gpurun_out
python tools/diag_conf.py 0.5 0.25 0.001 > gpurun_out/r2c_diag.txt 2>&1
cat gpurun_out/r2c_diag.txt
( time python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "nms or detect or graph or smoke or filter" --timeout 1200 ) > gpurun_out/r2c_pytest.log 2>&1
tail -15 gpurun_out/r2c_pytest.log
