#!/bin/bash
# edge kernel: L1 prefetch hints for the statistics / boxes the next culling level reads (1: sub-tile stats, 2: boxes, 4: next tile stats)
mkdir -p gpurun_out
cp yolo-from-scratch_b200/libyolo_b200.so /tmp/lib_base.so
for v in base pf1 pf2 pf3 pf7; do
  if [ $v = base ]; then cp /tmp/lib_base.so yolo-from-scratch_b200/libyolo_b200.so; else cp build_variants/lib_$v.so yolo-from-scratch_b200/libyolo_b200.so; fi
  echo "== $v"
  timeout 600 python tools/diag_conf.py 0.5 0.001 2>&1 | grep -E "edge" | tee -a gpurun_out/r2u_diag.txt
done
cp /tmp/lib_base.so yolo-from-scratch_b200/libyolo_b200.so
