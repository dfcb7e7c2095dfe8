"""Per-kernel times of detect_batch at several confidence thresholds (diagnosis helper)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import yolo_from_scratch_b200 as yb
from yolo_from_scratch_b200 import ops
import bench
lib = yb._lib.lib()
dev = torch.device("cuda")
anchors = ops.default_anchors(dev)
import itertools
sets = [[h.to(dev) for h in bench.make_heads(64, 640, 1, 1234 + 1000 * k)] for k in range(1)]
if os.environ.get("PRIOR") == "1":   # SURVEY 8d's "prior" heads: objectness logit 2*randn - 4.6
    for hs in sets:
        for h in hs:
            h[..., 4].mul_(2.0).sub_(4.6)
for heads, conf in itertools.product(sets, [float(x) for x in sys.argv[1:]] or [0.5, 0.25, 0.001]):
    for _ in range(3):
        det = ops.detect_batch(heads, anchors, 640, 1, conf, 0.4)
    torch.cuda.synchronize()
    lib.yb_timing_enable(1)
    for _ in range(5):
        det = ops.detect_batch(heads, anchors, 640, 1, conf, 0.4)
        if os.environ.get("PACK") == "1":
            ops.pack_detections(det)
    torch.cuda.synchronize()
    lib.yb_timing_enable(0)
    buf = ctypes.create_string_buffer(1 << 16)
    lib.yb_timing_collect(buf, len(buf))
    st = ops.nms_graph_stats(det)
    print("conf", conf, "cand/img", float(det["counts"].float().mean()), "max", int(det["counts"].max()), "kept/img", float(det["n_keep"].float().mean()),
          "evals", st[0], "edges", st[1], "cands", st[2], "cands/edge", st[2] / max(st[1], 1))
    for ln in buf.value.decode().strip().splitlines():
        n, c, t = ln.split()
        print("   %-28s %8.1f us" % (n, float(t) / int(c) * 1e3))
