#!/bin/bash
# spatial (batched loads), gather (butterfly sub-tile reductions), score sort (keys kept in shared memory) + experiments:
# loss fork point in HotPathGraph, 5 edge CTAs per SM
mkdir -p gpurun_out
python tools/diag_conf.py 0.5 0.001 > gpurun_out/r2p_diag.txt 2>&1
cat gpurun_out/r2p_diag.txt
python tools/diag_cfg.py 80 1280 32 0.001 > gpurun_out/r2p_diag_cfg.txt 2>&1
tail -40 gpurun_out/r2p_diag_cfg.txt
python tools/try_fork_point.py 1 > gpurun_out/r2p_fork.txt 2>&1
cat gpurun_out/r2p_fork.txt
( time python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "nms or detect or graph or smoke or filter" --timeout 1200 ) > gpurun_out/r2p_pytest.log 2>&1
tail -5 gpurun_out/r2p_pytest.log
cp yolo-from-scratch_b200/libyolo_b200.so /tmp/lib_base.so
cp build_variants/lib_edge5.so yolo-from-scratch_b200/libyolo_b200.so
python tools/diag_conf.py 0.5 0.001 > gpurun_out/r2p_diag_edge5.txt 2>&1
cat gpurun_out/r2p_diag_edge5.txt
cp /tmp/lib_base.so yolo-from-scratch_b200/libyolo_b200.so
