#!/bin/bash
# fused detection list (resolve kernel writes the rows): parity subset + step time
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "nms or detect or graph or smoke or filter or pack or fused" --timeout 600 ) > gpurun_out/r2r_pytest.log 2>&1
tail -15 gpurun_out/r2r_pytest.log
timeout 600 python tools/try_fork_point.py 1 2>&1 | tail -3 | tee gpurun_out/r2r_fork.txt
timeout 600 python tools/diag_conf.py 0.5 0.001 > gpurun_out/r2r_diag.txt 2>&1; cat gpurun_out/r2r_diag.txt
