#!/bin/bash
mkdir -p gpurun_out
python tools/diag_cfg.py 80 640 64 0.001 > gpurun_out/r2h_diag.txt 2>&1
python tools/diag_cfg.py 80 1280 32 0.001 >> gpurun_out/r2h_diag.txt 2>&1
python tools/diag_cfg.py 80 640 64 0.25 >> gpurun_out/r2h_diag.txt 2>&1
python tools/diag_cfg.py 1 640 64 0.5 >> gpurun_out/r2h_diag.txt 2>&1
cat gpurun_out/r2h_diag.txt
( time python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "nms or detect or graph or smoke or filter or decode or config" --timeout 1200 ) > gpurun_out/r2h_pytest.log 2>&1
tail -4 gpurun_out/r2h_pytest.log
