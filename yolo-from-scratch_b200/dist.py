"""Data-parallel use of the hot path: the path shards by image.

Decode, candidate filter, NMS and target assignment are independent per image: no collective.
The loss terms are means over the GLOBAL batch (train.py:710, :823, :827), so every rank
produces S x {sum(1-CIoU), n_pos, sum BCE_obj, sum BCE_cls} for its images and one all-reduce(sum)
of those S*4 doubles (96 bytes for three scales) gives the global values; the objectness normaliser
B_global*H*W*A is known without communication and the positive rows' gradients are scaled by the
reduced 1/n_pos afterwards (yb_loss_finalize).  One process per GPU, torch.distributed / NCCL.
"""
from typing import Tuple

import torch


def shard_range(n_images: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of images owned by `rank` (first n % world ranks get one extra)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(n_images, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_partials(partials: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum of the (S*4,) float64 partial sums over the ranks; a single small collective."""
    import torch.distributed as dist
    if partials.dtype != torch.float64:
        raise TypeError("partials must be float64")
    dist.all_reduce(partials, op=dist.ReduceOp.SUM, group=group)
    return partials


def global_batch(local_batch: int, group=None) -> int:
    """Sum of the ranks' local batch sizes (ranks may own different numbers of images)."""
    import torch.distributed as dist
    t = torch.tensor([local_batch], dtype=torch.int64)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def yolo_loss_multiscale_sharded(predictions, targets, anchors_list, num_classes=1, group=None):
    """yolo_loss_multiscale (train.py:840-886) over a batch sharded by image across the ranks of
    `group`: every rank passes ITS images and receives the GLOBAL losses; `total.backward()` leaves
    on every rank exactly the gradient rows of its own images of the global-batch loss."""
    from . import ops
    S = min(len(predictions), len(targets), len(anchors_list), len(ops.MULTISCALE_OBJ_WEIGHTS))
    return ops._loss_common(list(predictions[:S]), list(targets[:S]), list(anchors_list[:S]), num_classes,
                            ops.MULTISCALE_OBJ_WEIGHTS[:S], group=group)
