"""world_size-2 gloo tests (CPU) of the data-parallel protocol of the loss: shard by
image, S*4 partial sums per rank, ONE all-reduce, global normalisers.  The kernels themselves
need a GPU; here the per-shard partial sums come from the oracle and the collective, sharding and
combination logic are the product's (yolo-from-scratch_b200/dist.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def make_case(B, nc, seed):
    from oracle import ref_path as R
    g = torch.Generator().manual_seed(seed)
    grids = (8, 4, 2)
    preds = [torch.randn(B, G, G, 3, 5 + nc, generator=g) for G in grids]
    rng = np.random.default_rng(seed)
    tgts = []
    for s, G in enumerate(grids):
        t = torch.zeros(B, G, G, 3, 5 + nc)
        for b in range(B):
            for _ in range(int(rng.integers(0, 4))):
                gy, gx, a = rng.integers(0, G), rng.integers(0, G), rng.integers(0, 3)
                t[b, gy, gx, a, :5] = torch.tensor([rng.uniform(0.1, 0.9), rng.uniform(0.1, 0.9), rng.uniform(0.05, 0.5),
                                                    rng.uniform(0.05, 0.5), 1.0])
                t[b, gy, gx, a, 5 + int(rng.integers(0, nc))] = 1.0
        tgts.append(t)
    return preds, tgts, R.default_anchors()


def oracle_partials(preds, tgts, anchors, nc):
    """What yb_loss_partials accumulates for a shard: per scale {sum(1-CIoU), n_pos, sum BCE_obj, sum BCE_cls}."""
    from oracle import ref_path as R
    out = torch.zeros(len(preds) * 4, dtype=torch.float64)
    bce = torch.nn.functional.binary_cross_entropy_with_logits
    for s, (p, t, a) in enumerate(zip(preds, tgts, anchors)):
        pos = t[..., 4] > 0.5
        n = int(pos.sum())
        if n:
            dec = R.decode(p, a)
            out[s * 4 + 0] = float(R.ciou(dec[..., :4][pos], t[..., :4][pos])) * n
            out[s * 4 + 3] = float(bce(p[..., 5:][pos], t[..., 5:][pos], reduction="sum"))
        out[s * 4 + 1] = n
        out[s * 4 + 2] = float(bce(p[..., 4], t[..., 4], reduction="sum"))
    return out


def combine(partials, n_obj, nc, weights):
    """yb_loss_finalize's scalar part."""
    total = box_sum = obj_sum = cls_sum = 0.0
    for s, w in enumerate(weights):
        sb, n, so, sc = (float(x) for x in partials[s * 4:s * 4 + 4])
        box = sb / n if n > 0 else 0.0
        obj = so / n_obj[s]
        cls = sc / (n * nc) if n > 0 and nc > 0 else 0.0
        total += 0.05 * box + w * obj + 0.5 * cls
        box_sum += box; obj_sum += obj; cls_sum += cls
    return total, box_sum, obj_sum, cls_sum


def worker(rank, world, port, B, nc, seed, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from yolo_from_scratch_b200 import dist as ybd
        from yolo_from_scratch_b200.ops import MULTISCALE_OBJ_WEIGHTS
        preds, tgts, anchors = make_case(B, nc, seed)
        lo, hi = ybd.shard_range(B, rank, world)
        part = oracle_partials([p[lo:hi] for p in preds], [t[lo:hi] for t in tgts], anchors, nc)
        ybd.allreduce_partials(part)                       # the single collective
        b_global = ybd.global_batch(hi - lo)
        n_obj = [b_global * p.shape[1] * p.shape[2] * p.shape[3] for p in preds]
        q.put((rank, lo, hi, b_global, combine(part, n_obj, nc, MULTISCALE_OBJ_WEIGHTS)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B,nc", [(6, 1), (5, 3)])
def test_two_rank_loss_equals_single_process(B, nc):
    from oracle import ref_path as R
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, B, nc, 17, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # shards tile the batch
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == B and res[0][3] == res[1][3] == B
    preds, tgts, anchors = make_case(B, nc, 17)
    ref = [float(v) for v in R.multiscale_loss(preds, tgts, anchors, nc)]
    for _, _, _, _, got in res:
        np.testing.assert_allclose(got, ref, rtol=2e-6, atol=1e-7)


def test_shard_range_partitions():
    from yolo_from_scratch_b200.dist import shard_range
    for n in (0, 1, 7, 64, 512):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)
