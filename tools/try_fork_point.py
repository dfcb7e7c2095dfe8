"""Experiment: where the loss branch of HotPathGraph forks from the detection chain (start vs after the filter)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import yolo_from_scratch_b200 as yb
from yolo_from_scratch_b200 import ops
import bench
dev = torch.device("cuda")
B, img = 64, 640
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 1
conf = 0.5 if nc == 1 else 0.001
use_labels = len(sys.argv) > 2 and sys.argv[2] == "labels"
layout = 1 if (len(sys.argv) > 3 and sys.argv[3] == "nchw") else 0
anchors = ops.default_anchors(dev)
sets = []
for k in range(4):
    heads = [h.to(dev) for h in bench.make_heads(B, img, nc, 1234 + 1000 * k)]
    labels = bench.make_labels(np.random.default_rng(4321 + k), B, nc)
    if layout == 1:
        heads = [h.permute(0, 3, 4, 1, 2).reshape(B, -1, h.shape[1], h.shape[2]).contiguous() for h in heads]
    sets.append((heads, ops.pack_labels(labels, img) if use_labels else ops.build_targets(labels, anchors, [80, 40, 20], nc, img)))
def build(**kw):
    return [yb.HotPathGraph(B, img, nc, anchors, conf, 0.4, max_gt=50, adopt_heads=h, adopt_targets=t, layout=layout, **kw) for h, t in sets]
def run(graphs, steps=40):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        graphs[i % 4].replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
variants = {"start": build(fork_loss="start"), "after_filter": build(fork_loss="after_filter"), "no_overlap": build(overlap_loss=False)}
ref = variants["start"][0]
for name, gs in variants.items():
    gs[0].replay()
    torch.cuda.synchronize()
    assert torch.equal(gs[0].det["n_keep"], ref.det["n_keep"]) and torch.allclose(gs[0].losses, ref.losses), name
for _ in range(3):
    for name, gs in variants.items():
        run(gs, 8)
        print(name, [round(run(gs) * 1e3, 1) for _ in range(3)], "us/step")
