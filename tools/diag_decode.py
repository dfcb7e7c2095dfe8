"""decode fwd/bwd timing over the three heads: nc img B."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yolo_from_scratch_b200 as yb
from yolo_from_scratch_b200 import ops
import bench
nc, img, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda")
sets = [([h.to(dev) for h in bench.make_heads(B, img, nc, 1234 + k)]) for k in range(2)]
anchors = ops.default_anchors(dev)
res = bench.time_decode(yb._lib.lib(), [(s, None, None) for s in sets], anchors, img, nc, 2, 10)
for k, (ms, nbytes) in res.items():
    print(f"nc={nc} img={img} B={B} {k}: {ms*1e3:.1f} us, {nbytes/1e6:.0f} MB algorithmic, {nbytes/ms/1e6:.0f} GB/s = {nbytes/ms/1e6/6546.6:.2f} of the copy peak")
