#!/bin/bash
# spatial key width: 14 (base) / 13 / 12 bits — histogram size and scan cost vs culling quality
mkdir -p gpurun_out
cp yolo-from-scratch_b200/libyolo_b200.so /tmp/lib_base.so
for v in base kb13 kb12; do
  if [ $v = base ]; then cp /tmp/lib_base.so yolo-from-scratch_b200/libyolo_b200.so; else cp build_variants/lib_$v.so yolo-from-scratch_b200/libyolo_b200.so; fi
  echo "== $v"
  timeout 600 python tools/diag_conf.py 0.5 0.001 2>&1 | grep -E "spatial|edge|evals" | tee -a gpurun_out/r2v_diag.txt
  timeout 600 python tools/diag_cfg.py 80 640 64 0.001 2>&1 | grep -E "spatial|edge" | tee -a gpurun_out/r2v_diag.txt
  timeout 600 python tools/diag_cfg.py 80 1280 32 0.001 2>&1 | grep -E "spatial|edge" | tee -a gpurun_out/r2v_diag.txt
done
cp /tmp/lib_base.so yolo-from-scratch_b200/libyolo_b200.so
