"""One fused loss fwd+bwd and one decode fwd/bwd on the bench's heads (nc from argv); used under ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import yolo_from_scratch_b200 as yb
from yolo_from_scratch_b200 import ops
import bench
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda")
B, img = 64, 640
heads = [h.to(dev) for h in bench.make_heads(B, img, nc, 1234)]
labels = bench.make_labels(np.random.default_rng(4321), B, nc)
anchors = ops.default_anchors(dev)
tg = ops.build_targets(labels, anchors, [80, 40, 20], nc, img)
for _ in range(3):
    out4, _, grads = ops.loss_forward_backward(heads, tg, anchors, nc, ops.MULTISCALE_OBJ_WEIGHTS, [True] * 3)
    for s, h in enumerate(heads):
        y = torch.empty_like(h)
        yb._lib.lib().yb_decode_fwd(h.data_ptr(), anchors[s].data_ptr(), y.data_ptr(), B, h.shape[1], h.shape[2], 3, nc, float(img), torch.cuda.current_stream().cuda_stream)
        yb._lib.lib().yb_decode_bwd(h.data_ptr(), anchors[s].data_ptr(), y.data_ptr(), grads[s].data_ptr(), B, h.shape[1], h.shape[2], 3, nc, float(img), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("ok", float(out4[0]))
