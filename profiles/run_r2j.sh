#!/bin/bash
# A/B: edge kernel with 4 (64 regs) vs 3 (<=85 regs) CTAs per SM
python tools/diag_conf.py 0.5 2>&1 | grep "graph_edge\|conf"
cp yolo-from-scratch_b200/libyolo_b200.so /tmp/lib_orig.so
cp build/libyolo_b200_ctas3.so yolo-from-scratch_b200/libyolo_b200.so
python tools/diag_conf.py 0.5 2>&1 | grep "graph_edge\|conf"
cp /tmp/lib_orig.so yolo-from-scratch_b200/libyolo_b200.so
