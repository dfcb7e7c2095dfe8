#!/bin/bash
# edge kernel: ticket prefetch (E1), ticket bound by the largest image (E2), hoisted per-item loads (E3), one at a time
mkdir -p gpurun_out
cp yolo-from-scratch_b200/libyolo_b200.so /tmp/lib_base.so
for v in base e100 e010 e001 e110; do
  if [ $v = base ]; then cp /tmp/lib_base.so yolo-from-scratch_b200/libyolo_b200.so; else cp build_variants/lib_$v.so yolo-from-scratch_b200/libyolo_b200.so; fi
  echo "== $v"
  timeout 600 python tools/diag_conf.py 0.5 0.001 2>&1 | grep -E "edge|gather|conf" | tee -a gpurun_out/r2s_diag.txt
done
cp /tmp/lib_base.so yolo-from-scratch_b200/libyolo_b200.so
