"""Per-kernel times of detect_batch for a BASELINE config (diagnosis helper): nc img batch conf."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import yolo_from_scratch_b200 as yb
from yolo_from_scratch_b200 import ops
import bench
nc, img, B, conf = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
lib = yb._lib.lib()
dev = torch.device("cuda")
anchors = ops.default_anchors(dev)
heads = [h.to(dev) for h in bench.make_heads(B, img, nc, 1234)]
for _ in range(3):
    det = ops.detect_batch(heads, anchors, img, nc, conf, 0.4)
torch.cuda.synchronize()
lib.yb_timing_enable(1)
for _ in range(5):
    det = ops.detect_batch(heads, anchors, img, nc, conf, 0.4)
torch.cuda.synchronize()
lib.yb_timing_enable(0)
buf = ctypes.create_string_buffer(1 << 16)
lib.yb_timing_collect(buf, len(buf))
st = ops.nms_graph_stats(det)
print("nc", nc, "img", img, "B", B, "conf", conf, "cand/img", float(det["counts"].float().mean()), "kept/img", float(det["n_keep"].float().mean()),
      "evals", st[0], "edges", st[1], "cands", st[2])
for ln in buf.value.decode().strip().splitlines():
    n, c, t = ln.split()
    print("   %-28s %8.1f us" % (n, float(t) / int(c) * 1e3))
