#!/bin/bash
cp yolo-from-scratch_b200/libyolo_b200.so /tmp/lib_orig.so
for c in orig 13440 6720; do
  if [ $c != orig ]; then cp tools/lib_chunk$c.so yolo-from-scratch_b200/libyolo_b200.so; fi
  echo "== chunk $c"
  python tools/diag_conf.py 0.5 2>&1 | grep "graph_\|conf"
done
cp /tmp/lib_orig.so yolo-from-scratch_b200/libyolo_b200.so
