"""One detect_batch (filter + graph NMS) on the bench's configs[1] heads; used under ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yolo_from_scratch_b200 as yb
from yolo_from_scratch_b200 import ops
import bench
conf = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
nc = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda")
heads = [h.to(dev) for h in bench.make_heads(64, 640, nc, 1234)]
anchors = ops.default_anchors(dev)
for _ in range(3):
    det = ops.detect_batch(heads, anchors, 640, nc, conf, 0.4)
torch.cuda.synchronize()
print("ok", int(det["n_keep"].sum()))
