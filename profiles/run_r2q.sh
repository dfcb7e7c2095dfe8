#!/bin/bash
# programmatic dependent launch on the detection chain: off / on (early trigger) / on (no early trigger)
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "nms or detect or graph or smoke or filter" --timeout 1200 ) > gpurun_out/r2q_pytest.log 2>&1
tail -5 gpurun_out/r2q_pytest.log
echo "== PDL on, early trigger"; python tools/try_fork_point.py 1 2>&1 | tail -3 | tee gpurun_out/r2q_pdl_early.txt
echo "== PDL off"; YB_PDL=0 python tools/try_fork_point.py 1 2>&1 | tail -3 | tee gpurun_out/r2q_pdl_off.txt
cp yolo-from-scratch_b200/libyolo_b200.so /tmp/lib_base.so
cp build_variants/lib_pdl_noearly.so yolo-from-scratch_b200/libyolo_b200.so
echo "== PDL on, no early trigger"; python tools/try_fork_point.py 1 2>&1 | tail -3 | tee gpurun_out/r2q_pdl_noearly.txt
cp /tmp/lib_base.so yolo-from-scratch_b200/libyolo_b200.so
echo "== nc80 PDL on early"; python tools/try_fork_point.py 80 2>&1 | tail -3 | tee gpurun_out/r2q_pdl_nc80.txt
echo "== nc80 PDL off"; YB_PDL=0 python tools/try_fork_point.py 80 2>&1 | tail -3 | tee gpurun_out/r2q_pdl_nc80_off.txt
python tools/diag_conf.py 0.5 > gpurun_out/r2q_diag.txt 2>&1; cat gpurun_out/r2q_diag.txt
