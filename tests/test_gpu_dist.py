"""Multi-GPU (NCCL) parity of the image-sharded loss.  Needs >= 2 GPUs: skipped on a
single-GPU box (the protocol itself is covered on CPU by tests/test_dist_cpu.py and on one GPU by
test_loss_two_rank_emulation_on_one_gpu); run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case(B, nc, grids, img, seed):
    g = torch.Generator().manual_seed(seed)
    heads = [torch.randn(B, G, G, 3, 5 + nc, generator=g) for G in grids]
    rng = np.random.default_rng(seed)
    labels = []
    for _ in range(B):
        n = int(rng.integers(0, 12))
        lab = np.zeros((n, 5))
        lab[:, 0] = rng.integers(0, nc, n)
        lab[:, 1:3] = rng.uniform(0.05, 0.95, (n, 2))
        lab[:, 3:5] = np.exp(rng.uniform(np.log(0.02), np.log(0.6), (n, 2)))
        labels.append(lab)
    return heads, labels


def _worker(rank, world, port, B, nc, grids, img, seed, sparse, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import yolo_from_scratch_b200 as yb
        from yolo_from_scratch_b200 import dist as ybd
        from yolo_from_scratch_b200 import ops
        from oracle.ref_path import default_anchors
        anchors = default_anchors()
        heads, labels = _case(B, nc, grids, img, seed)
        lo, hi = ybd.shard_range(B, rank, world)
        preds = [h[lo:hi].cuda().requires_grad_(True) for h in heads]
        if sparse:
            out = ops.yolo_loss_multiscale_labels(preds, labels[lo:hi], anchors, nc, img, group=dist.group.WORLD)
        else:
            tg = yb.build_targets(labels[lo:hi], anchors, list(grids), nc, img)
            out = ybd.yolo_loss_multiscale_sharded(preds, tg, anchors, nc, group=dist.group.WORLD)
        out[0].backward()
        # NMS needs no collective: every rank detects on its own images
        det = yb.detect_batch([p.detach() for p in preds], anchors, img, nc, 0.3, 0.4)
        q.put((rank, lo, hi, [float(o) for o in out], [p.grad.cpu() for p in preds], det["n_keep"].cpu()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("sparse", [False, True])
def test_nccl_sharded_loss_equals_single_gpu(sparse):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    import yolo_from_scratch_b200 as yb
    from oracle.ref_path import default_anchors
    B, nc, grids, img, seed = 6, 3, (20, 10, 5), 160, 17
    heads, labels = _case(B, nc, grids, img, seed)
    anchors = default_anchors()
    full_p = [h.cuda().requires_grad_(True) for h in heads]
    tg = yb.build_targets(labels, anchors, list(grids), nc, img)
    full = yb.yolo_loss_multiscale(full_p, tg, anchors, nc)
    full[0].backward()
    det = yb.detect_batch([p.detach() for p in full_p], anchors, img, nc, 0.3, 0.4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, nc, grids, img, seed, sparse, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, lo, hi, out, grads, n_keep in results:
        for a, b in zip(out, full):
            assert abs(a - float(b)) <= 2e-6 * abs(float(b)) + 1e-7, (a, float(b))
        for s in range(3):
            ref = full_p[s].grad[lo:hi].cpu()
            assert torch.allclose(grads[s], ref, rtol=1e-5, atol=1e-9)
        assert torch.equal(n_keep, det["n_keep"][lo:hi].cpu())
