"""Experiment: replay the HotPathGraph instances of the rotating input sets on one stream vs two alternating streams."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import yolo_from_scratch_b200 as yb
from yolo_from_scratch_b200 import ops
import bench
dev = torch.device("cuda")
B, img, nc = 64, 640, 1
anchors = ops.default_anchors(dev)
graphs = []
for k in range(4):
    heads = [h.to(dev) for h in bench.make_heads(B, img, nc, 1234 + 1000 * k)]
    labels = bench.make_labels(np.random.default_rng(4321 + k), B, nc)
    tg = ops.build_targets(labels, anchors, [80, 40, 20], nc, img)
    graphs.append(yb.HotPathGraph(B, img, nc, anchors, 0.5, 0.4, max_gt=50, adopt_heads=heads, adopt_targets=tg))
def run(n_streams, steps=40):
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    e0.record()
    for s in streams: s.wait_event(e0)
    for i in range(steps):
        with torch.cuda.stream(streams[i % n_streams]):
            graphs[i % 4].replay()
    for s in streams: cur.wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
for _ in range(2):
    for n in (1, 2, 4):
        print(n, "streams:", round(run(n) * 1e3, 1), "us/step")
