#!/usr/bin/env python
"""One small pass over every kernel of libyolo_b200.so, for compute-sanitizer:

    compute-sanitizer --tool memcheck  --kernel-regex kns=yb:: python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck --kernel-regex kns=yb:: python tools/sanitize_smoke.py

Shapes are tiny but cover: both NMS algorithms in both torchvision regimes, partial tiles, the
shared-memory sort and the ballot sort (cap above the shared-memory limit), staged and unstaged filter
groups, dense / sparse targets, both head layouts, decode fwd/bwd, CIoU, eval counting, packing and the
CUDA-graph step.  Results are checked against the dense NMS / the reference layout so that a silent
corruption fails loudly too."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yolo_from_scratch_b200 as yb  # noqa: E402
from yolo_from_scratch_b200 import ops  # noqa: E402

ANCH = ops.default_anchors()


def main():
    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    B, nc, img, grids = 3, 5, 160, (20, 10, 5)
    raw = [torch.randn(B, 3 * (5 + nc), G, G).cuda() for G in grids]
    heads = [yb.heads_from_nchw(r) for r in raw]
    labels = []
    for _ in range(B):
        n = int(rng.integers(0, 9))
        lab = np.zeros((n, 5))
        lab[:, 0] = rng.integers(0, nc, n)
        lab[:, 1:3] = rng.uniform(0.1, 0.9, (n, 2))
        lab[:, 3:5] = rng.uniform(0.05, 0.5, (n, 2))
        labels.append(lab)
    tg = yb.build_targets(labels, ANCH, list(grids), nc, img)
    # decode fwd/bwd, ciou
    x = heads[0].clone().requires_grad_(True)
    yb.decode_predictions(x, ANCH[0], img).sum().backward()
    p = torch.rand(40, 4).cuda().requires_grad_(True)
    yb.ciou_loss(p, torch.rand(40, 4).cuda()).backward()
    # loss: dense / labels / nchw
    outs = []
    for fn, inp, t in ((yb.yolo_loss_multiscale, heads, tg), (yb.yolo_loss_multiscale_labels, heads, labels),
                       (yb.yolo_loss_multiscale_nchw, raw, tg), (yb.yolo_loss_multiscale_nchw, raw, labels)):
        q = [h.clone().requires_grad_(True) for h in inp]
        o = fn(q, t, ANCH, nc, img) if fn is not yb.yolo_loss_multiscale else fn(q, t, ANCH, nc)
        o[0].backward()
        outs.append(float(o[0]))
    assert max(outs) - min(outs) <= 1e-5 * abs(outs[0]), outs
    # filter + NMS: graph vs dense algorithm, both layouts, dense and sparse thresholds
    for conf in (0.001, 0.5, 0.97):
        d0 = yb.detect_batch(heads, ANCH, img, nc, conf, 0.4)
        d1 = yb.detect_batch(heads, ANCH, img, nc, conf, 0.4, algo=yb.NMS_BITMASK)
        d2 = yb.detect_batch_nchw(raw, ANCH, img, nc, conf, 0.4)
        for d in (d1, d2):
            assert torch.equal(d0["n_keep"], d["n_keep"])
            for b in range(B):
                k = int(d0["n_keep"][b])
                assert torch.equal(d0["keep"][b, :k], d["keep"][b, :k])
        yb.pack_detections(d0)
    # plain nms, trick / per-class regimes, ballot-sort path (cap > shared-memory sort limit), ties
    for n, ncls in ((1, 1), (33, 1), (700, 4), (27000, 3)):
        c = torch.rand(n, 2).cuda() * (300 if n < 1000 else 4000)  # keep the big case's graph sparse
        wh = torch.rand(n, 2).cuda() * 60 + 2
        boxes = torch.cat([c - wh / 2, c + wh / 2], 1)
        scores = (torch.rand(n).cuda() * 8).floor() / 8 if n == 700 else torch.rand(n).cuda()
        idxs = torch.randint(0, ncls, (n,)).cuda()
        a = yb.batched_nms(boxes, scores, idxs, 0.5)
        b = yb.batched_nms(boxes, scores, idxs, 0.5, algo=yb.NMS_BITMASK)
        assert torch.equal(a, b), n
    # eval counting, CUDA-graph step
    yb.eval_counts(heads, tg, ANCH, 0.5, 0.5)
    hp = yb.HotPathGraph(B, img, nc, ANCH, 0.3, 0.4, max_gt=8)
    for dst, src in zip(hp.heads, heads):
        dst.copy_(src)
    lab, n_gt, lb = ops.pack_labels_host(labels, img, max_gt=8)
    hp.labels.labels.copy_(lab); hp.labels.n_gt.copy_(n_gt); hp.labels.letterbox.copy_(lb)
    losses, _, _ = hp.replay()
    torch.cuda.synchronize()
    assert abs(float(losses[0]) - outs[0]) <= 1e-5 * abs(outs[0])
    print("sanitize_smoke ok: launches =", yb._lib.lib().yb_launch_count())


if __name__ == "__main__":
    main()
