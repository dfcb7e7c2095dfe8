"""pack_detections alone (CUDA events over repeated calls) on the bench's heads: dense and SURVEY 8d 'prior'."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yolo_from_scratch_b200 as yb
from yolo_from_scratch_b200 import ops
import bench
dev = torch.device("cuda")
anchors = ops.default_anchors(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, prior, conf in (("dense conf 0.5", False, 0.5), ("prior conf 0.25", True, 0.25)):
    heads = [h.to(dev) for h in bench.make_heads(64, 640, 1, 1234)]
    if prior:
        for h in heads:
            h[..., 4].mul_(2.0).sub_(4.6)
    det = ops.detect_batch(heads, anchors, 640, 1, conf, 0.4)
    rows, off = ops.pack_detections(det)
    torch.cuda.synchronize()
    ref = rows[:int(off[-1])].clone()
    for cold in (False, True):
        ts = []
        for _ in range(20):
            if cold:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rows, off = ops.pack_detections(det)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        assert torch.equal(rows[:int(off[-1])], ref)
        print(name, "L2 flushed" if cold else "L2 warm", "median %.1f us" % ts[len(ts) // 2], "kept", int(off[-1]))
