"""Host-side mirror of the reference's per-box hot-path functions (reference `train.py`).

Same names, argument meaning and return values as the reference so that callers and tests read
the same; the arithmetic runs in libyolo_b200.so (hand-written sm_100a kernels, C-ABI in
include/yolo_b200.h).  PyTorch is used only for device memory, streams and autograd plumbing.

  decode_predictions      train.py:712-779
  ciou_loss               train.py:634-710
  yolo_loss               train.py:781-838
  yolo_loss_multiscale    train.py:840-886
  compute_anchor_iou      train.py:108-131
  build_targets           train.py:147-205 (label loop of YOLODataset.__getitem__), batched
  filter_candidates       train.py:1152-1229 (per-scale body of predict), batched
  nms / batched_nms       torchvision.ops as called at train.py:1232-1233
  detect_batch            filter_candidates + batched_nms + gather (predict :1152-1238 for a batch)

CPU tensors are accepted (the reference's tests pass CPU tensors): they are staged to the GPU,
computed there, and staged back.  That is staging, not a fallback: without CUDA every function
raises.
"""
import ctypes
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import HeadsDesc, LossDesc, MAX_ANCHORS, MAX_SCALES

# reference constants
BOX_WEIGHT = 0.05                 # train.py:836
CLS_WEIGHT = 0.5                  # train.py:836
MULTISCALE_OBJ_WEIGHTS = (4.0, 1.0, 0.4)  # train.py:865
LOSS_DECODE_IMG_SIZE = 640.0      # train.py:796 — yolo_loss always decodes with the default
CIOU_EPS = 1e-7                   # train.py:634
DEFAULT_ANCHORS = (               # train.py:372-374 (model) / :81-83 (dataset): P3, P4, P5 anchors in pixels
    ((10, 13), (16, 30), (33, 23)),
    ((30, 61), (62, 45), (59, 119)),
    ((116, 90), (156, 198), (373, 326)),
)
TRICK_MAX_NUMEL_CUDA = 100_000    # torchvision/ops/boxes.py:80
TRICK_MAX_NUMEL_CPU = 4_000


# --------------------------------------------------------------------------------------------
# plumbing
# --------------------------------------------------------------------------------------------
def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "yolo_from_scratch_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def dist_global_batch(local_batch, group):
    from .dist import global_batch
    return global_batch(local_batch, group)


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


def _f32c(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    """fp32, contiguous, on `dev`, 16-byte aligned (differentiable)."""
    if t.device != dev:
        t = t.to(dev)
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    return t


_anchor_cache = {}


def _anchors_dev(anchors, dev: torch.device) -> torch.Tensor:
    """(A,2) fp32 anchors on the device; small host-side cache so a CPU anchor tensor is not
    re-uploaded on every call."""
    if isinstance(anchors, torch.Tensor):
        if anchors.device == dev and anchors.dtype == torch.float32 and anchors.is_contiguous():
            return anchors.detach()
        key = (anchors.data_ptr(), anchors._version, tuple(anchors.shape), str(anchors.dtype), dev.index)
        hit = _anchor_cache.get(key)
        if hit is not None and torch.equal(hit[0], anchors.detach().cpu() if anchors.device.type != "cpu" else anchors.detach()):
            return hit[1]
        host = anchors.detach().to("cpu", torch.float32).contiguous()
        out = host.to(dev)
        if len(_anchor_cache) > 64:
            _anchor_cache.clear()
        _anchor_cache[key] = (host.clone(), out)
        return out
    return torch.tensor(anchors, dtype=torch.float32, device=dev).reshape(-1, 2).contiguous()


def default_anchors(device=None) -> List[torch.Tensor]:
    """The reference's default anchors as three (3,2) fp32 tensors (train.py:372-374)."""
    return [torch.tensor(a, dtype=torch.float32, device=device) for a in DEFAULT_ANCHORS]


def _check_head(t: torch.Tensor, what: str):
    if t.dim() != 5:
        raise ValueError(f"{what} must be (B, H, W, A, 5+nc), got {tuple(t.shape)}")
    if t.shape[3] > MAX_ANCHORS:
        raise ValueError(f"{what}: at most {MAX_ANCHORS} anchors per scale")
    if t.shape[4] < 5:
        raise ValueError(f"{what}: last dimension must be 5+nc")


# --------------------------------------------------------------------------------------------
# a-1 decode_predictions
# --------------------------------------------------------------------------------------------
class _DecodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, anchors, img_size):
        B, H, W, A, row = pred.shape
        out = torch.empty_like(pred)
        _lib.check(_lib.lib().yb_decode_fwd(pred.data_ptr(), anchors.data_ptr(), out.data_ptr(),
                                            B, H, W, A, row - 5, float(img_size), _stream()), "yb_decode_fwd")
        ctx.save_for_backward(pred, anchors)
        ctx.img_size = float(img_size)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        pred, anchors = ctx.saved_tensors
        B, H, W, A, row = pred.shape
        grad_out = _f32c(grad_out, pred.device)
        grad_in = torch.empty_like(pred)
        _lib.check(_lib.lib().yb_decode_bwd(pred.data_ptr(), anchors.data_ptr(), grad_out.data_ptr(),
                                            grad_in.data_ptr(), B, H, W, A, row - 5, ctx.img_size, _stream()),
                   "yb_decode_bwd")
        return grad_in, None, None


def decode_predictions(raw_preds: torch.Tensor, anchors, img_size=640) -> torch.Tensor:
    """train.py:712-779.  (B,H,W,A,5+nc) raw logits -> same shape with xywh decoded to
    normalised coordinates; objectness/class logits unchanged.  Differentiable."""
    _check_head(raw_preds, "raw_preds")
    dev = _device()
    orig_dev, orig_dtype = raw_preds.device, raw_preds.dtype
    with torch.cuda.device(dev):
        pred = _f32c(raw_preds, dev)
        anc = _anchors_dev(anchors, dev)
        if anc.shape[0] != pred.shape[3]:
            raise ValueError("anchors must be (num_anchors, 2)")
        out = _DecodeFn.apply(pred, anc, img_size)
    if out.dtype != orig_dtype:
        out = out.to(orig_dtype)
    return out if orig_dev == dev else out.to(orig_dev)


# --------------------------------------------------------------------------------------------
# a-2 ciou_loss
# --------------------------------------------------------------------------------------------
class _CiouFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred_boxes, target_boxes, eps):
        N = pred_boxes.shape[0]
        L = _lib.lib()
        loss = torch.empty(1, dtype=torch.float32, device=pred_boxes.device)
        need_p, need_t = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gp = torch.empty_like(pred_boxes) if need_p else None
        gt = torch.empty_like(target_boxes) if need_t else None
        scratch = torch.empty(16, dtype=torch.uint8, device=pred_boxes.device)
        _lib.check(L.yb_ciou_fwd_bwd(pred_boxes.data_ptr(), target_boxes.data_ptr(), N, float(eps),
                                     loss.data_ptr(), _ptr(gp), _ptr(gt), scratch.data_ptr(), 16, _stream()),
                   "yb_ciou_fwd_bwd")
        ctx.grads = (gp, gt)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        gp, gt = ctx.grads
        return (None if gp is None else gp * g), (None if gt is None else gt * g), None


def ciou_loss(pred_boxes: torch.Tensor, target_boxes: torch.Tensor, eps=CIOU_EPS) -> torch.Tensor:
    """train.py:634-710.  mean(1 - CIoU) over N (x,y,w,h) pairs."""
    if pred_boxes.dim() != 2 or pred_boxes.shape[1] != 4 or pred_boxes.shape != target_boxes.shape:
        raise ValueError("ciou_loss expects two (N,4) tensors")
    dev = _device()
    orig_dev = pred_boxes.device
    with torch.cuda.device(dev):
        p = _f32c(pred_boxes, dev)
        t = _f32c(target_boxes, dev)
        out = _CiouFn.apply(p, t, eps)
    return out if orig_dev == dev else out.to(orig_dev)


# --------------------------------------------------------------------------------------------
# a-3 / a-4 fused loss
# --------------------------------------------------------------------------------------------
LAYOUT_BHWAC, LAYOUT_NCHW = _lib.LAYOUT_BHWAC, _lib.LAYOUT_NCHW


def _grid(p: torch.Tensor, layout: int):
    """(H, W) of a head tensor: (B,H,W,A,5+nc) or, NCHW, (B, A*(5+nc), H, W)."""
    return (p.shape[2], p.shape[3]) if layout == LAYOUT_NCHW else (p.shape[1], p.shape[2])


def heads_from_nchw(raw: torch.Tensor, num_anchors: int = 3) -> torch.Tensor:
    """The reference's own head reshape (train.py:608-609): (B, A*(5+nc), H, W) -> contiguous
    (B, H, W, A, 5+nc).  The *_nchw entry points make this pass unnecessary; tests use it."""
    B, C, H, W = raw.shape
    return raw.view(B, num_anchors, C // num_anchors, H, W).permute(0, 3, 4, 1, 2).contiguous()


def _fill_loss_desc(preds, tgts, ancs, grads, nc, obj_weights, coef, B_global=None, img_size=LOSS_DECODE_IMG_SIZE,
                    layout=LAYOUT_BHWAC):
    S = len(preds)
    d = LossDesc()
    d.S = S
    d.layout = layout
    d.B = preds[0].shape[0]
    d.B_global = d.B if B_global is None else int(B_global)
    d.A = ancs[0].shape[0]
    d.nc = nc
    d.img_size = img_size
    d.eps = CIOU_EPS
    d.w_box, d.w_cls = BOX_WEIGHT, CLS_WEIGHT
    for s in range(S):
        d.H[s], d.W[s] = _grid(preds[s], layout)
        d.w_obj[s] = obj_weights[s]
        d.coef_box[s], d.coef_obj[s], d.coef_cls[s] = coef[s]
        d.pred[s] = preds[s].data_ptr()
        d.tgt[s] = _ptr(tgts[s])
        d.anchors[s] = ancs[s].data_ptr()
        d.grad[s] = _ptr(grads[s])
    return d


def loss_forward_backward(preds, tgts, ancs, nc, obj_weights, want_grad, coef=None, group=None, equal_shards=True,
                          b_global=None, reduce_fn=None, sparse=None, layout=LAYOUT_BHWAC):
    """Run the fused kernels on device tensors.  Returns (out4, per_scale(S,3), grads list).

    `group`: optional torch.distributed process group; the batch is then a shard of a global
    batch and the S*4 partial sums are all-reduced once between the two stages.
    `sparse`: a `PackedLabels` — targets are assigned on the device from the label lists and handed
    to the loss in sparse form; `tgts` is then ignored (pass None).
    """
    L = _lib.lib()
    S = len(preds)
    dev = preds[0].device
    if coef is None:
        coef = [(BOX_WEIGHT, obj_weights[s], CLS_WEIGHT) for s in range(S)]
    grads = [torch.empty_like(p) if w else None for p, w in zip(preds, want_grad)]
    world = 1
    explicit = b_global
    b_global = preds[0].shape[0]
    if explicit is not None:
        b_global = int(explicit)
    elif group is not None:
        import torch.distributed as dist
        world = dist.get_world_size(group)
        b_global = b_global * world if equal_shards else dist_global_batch(preds[0].shape[0], group)
    if sparse is not None:
        tgts = [None] * S
    d = _fill_loss_desc(preds, tgts, ancs, grads, nc, obj_weights, coef, B_global=b_global, layout=layout)
    if sparse is None:
        ws_bytes = L.yb_loss_workspace_bytes(ctypes.byref(d))
    else:
        ws_bytes = L.yb_loss_sparse_workspace_bytes(ctypes.byref(d), sparse.max_gt)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    partials = torch.empty(S * 4, dtype=torch.float64, device=dev)
    out4 = torch.empty(4, dtype=torch.float32, device=dev)
    per_scale = torch.empty(S, 3, dtype=torch.float32, device=dev)
    st = _stream()
    if sparse is None:
        _lib.check(L.yb_loss_partials(ctypes.byref(d), partials.data_ptr(), ws.data_ptr(), ws_bytes, st),
                   "yb_loss_partials")
    else:
        if sparse.labels.shape[0] != preds[0].shape[0]:
            raise ValueError("labels and predictions disagree on the batch size")
        anc_all = torch.stack(list(ancs)).contiguous()  # (S,A,2)
        _lib.check(L.yb_loss_partials_sparse(ctypes.byref(d), sparse.labels.data_ptr(), sparse.n_gt.data_ptr(),
                                             sparse.letterbox.data_ptr(), anc_all.data_ptr(), sparse.max_gt,
                                             int(sparse.img_size), sparse.status.data_ptr(), partials.data_ptr(),
                                             ws.data_ptr(), ws_bytes, st), "yb_loss_partials_sparse")
    if reduce_fn is not None:
        partials = reduce_fn(partials)          # test hook: emulate the collective on one GPU
    elif world > 1:
        from .dist import allreduce_partials
        allreduce_partials(partials, group)
    _lib.check(L.yb_loss_finalize(ctypes.byref(d), partials.data_ptr(), out4.data_ptr(), per_scale.data_ptr(),
                                  ws.data_ptr(), ws_bytes, st), "yb_loss_finalize")
    return out4, per_scale, grads


class _YoloLossFn(torch.autograd.Function):
    """(total, bbox, obj, cls) = f(preds...) with the gradient of `total` produced in the same
    pass.  Gradients that arrive on bbox/obj/cls as well are honoured by re-running the kernels
    with the combined coefficients (rare: the reference only calls total.backward())."""

    @staticmethod
    def forward(ctx, nc, obj_weights, group, S, *tensors):
        sparse, layout = None, LAYOUT_BHWAC
        if isinstance(group, tuple):  # (process group, PackedLabels or None, layout)
            group, sparse, layout = group
        preds, tgts, ancs = tensors[:S], tensors[S:2 * S], tensors[2 * S:3 * S]
        want = [ctx.needs_input_grad[4 + s] for s in range(S)]
        out4, per_scale, grads = loss_forward_backward(preds, tgts, ancs, nc, obj_weights, want, group=group,
                                                       sparse=sparse, layout=layout)
        ctx.set_materialize_grads(False)
        ctx.sparse, ctx.layout = sparse, layout
        ctx.cfg = (nc, obj_weights, group, S, want)
        ctx.fused_grads = grads
        ctx.save_for_backward(*tensors)
        return out4[0], out4[1], out4[2], out4[3]

    @staticmethod
    def backward(ctx, g_total, g_bbox, g_obj, g_cls):
        nc, obj_weights, group, S, want = ctx.cfg
        none = (None,) * 4
        if not any(want):
            return none + (None,) * (3 * S)
        if g_bbox is None and g_obj is None and g_cls is None and ctx.fused_grads is not None:
            grads = ctx.fused_grads
            if g_total is None:
                grads = [None if g is None else torch.zeros_like(g) for g in grads]
            else:
                L = _lib.lib()
                f = g_total.detach().to(torch.float32).reshape(1).contiguous()
                for g in grads:
                    if g is not None:
                        _lib.check(L.yb_scale_inplace(g.data_ptr(), g.numel(), f.data_ptr(), _stream()),
                                   "yb_scale_inplace")
        else:
            # gradients on bbox/obj/cls as well, or a second backward through the same graph (retain_graph=True:
            # the fused buffers were handed out, and possibly scaled in place, by the first one): recompute
            tensors = ctx.saved_tensors
            preds, tgts, ancs = tensors[:S], tensors[S:2 * S], tensors[2 * S:3 * S]
            z = lambda g: 0.0 if g is None else float(g)
            gt_, gb_, go_, gc_ = z(g_total), z(g_bbox), z(g_obj), z(g_cls)
            coef = [(BOX_WEIGHT * gt_ + gb_, obj_weights[s] * gt_ + go_, CLS_WEIGHT * gt_ + gc_) for s in range(S)]
            _, _, grads = loss_forward_backward(preds, tgts, ancs, nc, obj_weights, want, coef=coef, group=group,
                                                sparse=ctx.sparse, layout=ctx.layout)
        ctx.fused_grads = None
        return none + tuple(grads) + (None,) * (2 * S)


def _loss_common(predictions, targets, anchors_list, num_classes, obj_weights, group=None, sparse=None,
                 layout=LAYOUT_BHWAC):
    S = len(predictions)
    if not 1 <= S <= MAX_SCALES:
        raise ValueError(f"1..{MAX_SCALES} scales supported, got {S}")
    dev = _device()
    orig_dev = predictions[0].device
    with torch.cuda.device(dev):
        preds, tgts, ancs = [], [], []
        for s in range(S):
            if layout == LAYOUT_NCHW:
                p, A = predictions[s], len(anchors_list[s])
                if p.dim() != 4 or p.shape[1] != A * (5 + num_classes):
                    raise ValueError(f"scale {s}: NCHW head must be (B, {A}*(5+num_classes), H, W), got {tuple(p.shape)}")
                if sparse is None and tuple(targets[s].shape) != (p.shape[0], p.shape[2], p.shape[3], A, 5 + num_classes):
                    raise ValueError(f"scale {s}: targets {tuple(targets[s].shape)} do not match the head")
            else:
                _check_head(predictions[s], f"predictions[{s}]")
                if sparse is None and predictions[s].shape != targets[s].shape:
                    raise ValueError(f"scale {s}: predictions {tuple(predictions[s].shape)} vs targets {tuple(targets[s].shape)}")
                if predictions[s].shape[4] != 5 + num_classes:
                    raise ValueError(f"scale {s}: last dim {predictions[s].shape[4]} != 5+num_classes")
            preds.append(_f32c(predictions[s], dev))
            # sparse targets: a placeholder keeps the autograd signature (preds, targets, anchors) x S
            tgts.append(_f32c(targets[s].detach(), dev) if sparse is None else preds[-1].detach())
            ancs.append(_anchors_dev(anchors_list[s], dev))
        packed_cfg = group if (sparse is None and layout == LAYOUT_BHWAC) else (group, sparse, layout)
        outs = _YoloLossFn.apply(int(num_classes), tuple(float(w) for w in obj_weights), packed_cfg, S,
                                 *preds, *tgts, *ancs)
    if orig_dev != dev:
        outs = tuple(o.to(orig_dev) for o in outs)
    return outs


class PackedLabels:
    """Label lists of a batch on the device, the input of the sparse-target loss:
    labels (B,max_gt,5) fp64 [class, xc, yc, w, h], n_gt (B) int32, letterbox (B,5) fp64
    [orig_w, orig_h, scale, pad_top, pad_left] (train.py:136-137), the dataset's img_size."""

    def __init__(self, labels, n_gt, letterbox, img_size):
        self.labels, self.n_gt, self.letterbox, self.img_size = labels, n_gt, letterbox, int(img_size)
        self.max_gt = int(labels.shape[1])
        self.status = torch.zeros(1, dtype=torch.int32, device=labels.device)

    def check(self):
        """Raise like the reference (IndexError, train.py:193-205) if a label mapped outside the grid."""
        if int(self.status.item()) != 0:
            raise IndexError("a label maps outside the grid or the class range (train.py:193-205)")


def pack_labels_host(labels: Sequence, img_size: int, letterbox: Optional[Sequence] = None, pin: bool = False,
                     max_gt: Optional[int] = None):
    """Host-side packing of per-image label arrays [(n_i,5)] into (labels (B,max_gt,5) f64, n_gt (B) i32,
    letterbox (B,5) f64) — 40 bytes per ground truth instead of three dense target tensors."""
    B = len(labels)
    rows = [torch.as_tensor(l, dtype=torch.float64).reshape(-1, 5) for l in labels]
    need = max([r.shape[0] for r in rows], default=0)
    if max_gt is None:
        max_gt = need
    elif max_gt < need:
        raise ValueError(f"max_gt={max_gt} < {need} labels in one image")
    lab = torch.zeros(B, max(max_gt, 1), 5, dtype=torch.float64)
    n_gt = torch.zeros(B, dtype=torch.int32)
    for i, r in enumerate(rows):
        lab[i, :r.shape[0]] = r
        n_gt[i] = r.shape[0]
    if letterbox is None:
        lb = torch.tensor([[img_size, img_size, 1.0, 0.0, 0.0]] * B, dtype=torch.float64).reshape(B, 5)
    else:
        lb = torch.as_tensor(np.asarray(letterbox, dtype=np.float64)).reshape(B, 5)
    if pin:
        lab, n_gt, lb = lab.pin_memory(), n_gt.pin_memory(), lb.pin_memory()
    return lab, n_gt, lb


def pack_labels(labels: Sequence, img_size: int, letterbox: Optional[Sequence] = None) -> PackedLabels:
    dev = _device()
    lab, n_gt, lb = pack_labels_host(labels, img_size, letterbox)
    return PackedLabels(lab.to(dev), n_gt.to(dev), lb.to(dev), img_size)


def yolo_loss_multiscale_labels(predictions, labels, anchors_list, num_classes=1, img_size=640, letterbox=None,
                                group=None, check=False):
    """yolo_loss_multiscale (train.py:840-886) fed by LABEL LISTS instead of dense targets: the
    assignment of YOLODataset.__getitem__ (:147-205) runs on the device and reaches the loss in
    sparse form.  `labels`: per-image (n_i,5) arrays, or a PackedLabels already on the device.
    Returns the same (total, sum bbox, sum obj, sum cls) as the dense call on the reference's targets.

    A label that maps outside the grid or the class range makes the reference (and `build_targets`) raise
    IndexError (train.py:193-205); here it is dropped on the device and recorded in `PackedLabels.status`.
    `check=True` reads that flag right after the call (one host synchronisation) and raises like the
    reference; asynchronous callers (CUDA graphs, pipelined loops) must call `PackedLabels.check()` at their
    own synchronisation point instead."""
    S = min(len(predictions), len(anchors_list), len(MULTISCALE_OBJ_WEIGHTS))
    packed = labels if isinstance(labels, PackedLabels) else pack_labels(labels, img_size, letterbox)
    out = _loss_common(list(predictions[:S]), [None] * S, list(anchors_list[:S]), num_classes,
                       MULTISCALE_OBJ_WEIGHTS[:S], group=group, sparse=packed)
    if check:
        packed.check()
    return out


def yolo_loss_multiscale_nchw(raw_heads, targets, anchors_list, num_classes=1, img_size=640, letterbox=None,
                              group=None, check=False):
    """yolo_loss_multiscale on the head convs' OWN outputs (B, A*(5+nc), H, W) — SURVEY 8f-2: the
    view/permute/contiguous of train.py:608-609 (a full read+write of every head, plus its backward)
    is not needed; the gradient comes back in the same NCHW layout, ready for the conv backward.
    `targets`: the reference's dense targets [(B,G,G,A,5+nc)], or label lists / PackedLabels (then
    the assignment runs on the device, as in yolo_loss_multiscale_labels)."""
    S = min(len(raw_heads), len(anchors_list), len(MULTISCALE_OBJ_WEIGHTS))
    is_packed = isinstance(targets, PackedLabels)
    dense = (not is_packed) and len(targets) > 0 and isinstance(targets[0], torch.Tensor) and targets[0].dim() == 5
    if not dense:
        packed = targets if is_packed else pack_labels(targets, img_size, letterbox)
        out = _loss_common(list(raw_heads[:S]), [None] * S, list(anchors_list[:S]), num_classes,
                           MULTISCALE_OBJ_WEIGHTS[:S], group=group, sparse=packed, layout=LAYOUT_NCHW)
        if check:
            packed.check()   # see yolo_loss_multiscale_labels
        return out
    S = min(S, len(targets))
    return _loss_common(list(raw_heads[:S]), list(targets[:S]), list(anchors_list[:S]), num_classes,
                        MULTISCALE_OBJ_WEIGHTS[:S], group=group, layout=LAYOUT_NCHW)


def yolo_loss(predictions, targets, anchors, num_classes=1):
    """train.py:781-838.  Returns (total, bbox, obj, cls) for one scale."""
    return _loss_common([predictions], [targets], [anchors], num_classes, (1.0,))


def yolo_loss_multiscale(predictions, targets, anchors_list, num_classes=1):
    """train.py:840-886.  Returns (total, sum bbox, sum obj (unweighted), sum cls)."""
    S = min(len(predictions), len(targets), len(anchors_list), len(MULTISCALE_OBJ_WEIGHTS))  # zip() semantics (:873)
    return _loss_common(list(predictions[:S]), list(targets[:S]), list(anchors_list[:S]), num_classes,
                        MULTISCALE_OBJ_WEIGHTS[:S])


# --------------------------------------------------------------------------------------------
# f-1 eval_epoch's detection counting
# --------------------------------------------------------------------------------------------
EVAL_DECODE_IMG_SIZE = 640.0  # train.py:993 — eval_epoch decodes with the default img_size


def eval_counts(predictions, targets, anchors_list, conf_threshold=0.5, iou_threshold=0.5, out=None):
    """The TP/FP/FN loop of eval_epoch (train.py:993-1024) for one batch, all scales, in one kernel.
    Returns an int64 device tensor [TP, FP, FN]; pass `out` to accumulate over batches (no sync)."""
    dev = _device()
    S = min(len(predictions), len(targets), len(anchors_list))
    with torch.cuda.device(dev):
        preds = [_f32c(p.detach(), dev) for p in predictions[:S]]
        tgts = [_f32c(t.detach(), dev) for t in targets[:S]]
        for s in range(S):
            _check_head(preds[s], f"predictions[{s}]")
            if preds[s].shape != tgts[s].shape:
                raise ValueError(f"scale {s}: predictions {tuple(preds[s].shape)} vs targets {tuple(tgts[s].shape)}")
        ancs = [_anchors_dev(a, dev) for a in anchors_list[:S]]
        d = _heads_desc(preds, ancs, EVAL_DECODE_IMG_SIZE, preds[0].shape[4] - 5)
        if out is None:
            out = torch.zeros(3, dtype=torch.int64, device=dev)
        tptr = (ctypes.c_void_p * MAX_SCALES)(*[t.data_ptr() for t in tgts])
        _lib.check(_lib.lib().yb_eval_counts(ctypes.byref(d), tptr, float(conf_threshold), float(iou_threshold),
                                             out.data_ptr(), _stream()), "yb_eval_counts")
    return out


def eval_epoch(model, loader, device, num_classes=1, iou_threshold=0.5, conf_threshold=0.5):
    """Drop-in for train.eval_epoch (train.py:960-1032): same arguments, same four return values
    (avg_loss, precision %, recall %, F1 %).  The loss and the detection counting run on the GPU
    path; the counters stay on the device until the end of the epoch (one synchronisation instead of
    2 x rows `.item()` calls per batch)."""
    model.eval()
    anchors_list = model.anchors
    counts = None
    loss_sum = None
    n_batches = 0
    with torch.no_grad():
        for imgs, targets in loader:
            imgs = imgs.to(device)
            S = len(targets[0])
            targets_batch = [torch.stack([t[s] for t in targets]).to(device) for s in range(S)]  # :973-977
            preds = model(imgs)
            loss = yolo_loss_multiscale(preds, targets_batch, anchors_list, num_classes)[0]        # :987
            loss_sum = loss.detach().double().cuda() if loss_sum is None else loss_sum + loss.detach().double().cuda()
            counts = eval_counts(preds, targets_batch, anchors_list, conf_threshold, iou_threshold, out=counts)
            n_batches += 1
    tp, fp, fn = (0, 0, 0) if counts is None else [int(v) for v in counts.cpu().tolist()]
    precision = tp / (tp + fp) if (tp + fp) > 0 else 0                                            # :1026-1029
    recall = tp / (tp + fn) if (tp + fn) > 0 else 0
    f1 = 2 * precision * recall / (precision + recall) if (precision + recall) > 0 else 0
    avg_loss = (float(loss_sum) if loss_sum is not None else 0.0) / n_batches                      # :1031 (ZeroDivisionError on an empty loader, like the reference)
    return avg_loss, precision * 100, recall * 100, f1 * 100


# --------------------------------------------------------------------------------------------
# a-5 target assignment
# --------------------------------------------------------------------------------------------
def compute_anchor_iou(box_wh, anchors) -> torch.Tensor:
    """train.py:108-131.  box_wh (2,) or (n,2) pixels, anchors (A,2) -> (A,) or (n,A) IoU."""
    dev = _device()
    box_wh = torch.as_tensor(box_wh)
    orig_dev = box_wh.device
    single = box_wh.dim() == 1
    with torch.cuda.device(dev):
        b = _f32c(box_wh.reshape(-1, 2), dev)
        a = _anchors_dev(anchors, dev)
        out = torch.empty(b.shape[0], a.shape[0], dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().yb_anchor_iou(b.data_ptr(), a.data_ptr(), out.data_ptr(), b.shape[0], a.shape[0], _stream()),
                   "yb_anchor_iou")
    out = out[0] if single else out
    return out if orig_dev == dev else out.to(orig_dev)


def build_targets(labels: Sequence, anchors_list, grid_sizes: Sequence[int], num_classes: int, img_size: int,
                  letterbox: Optional[Sequence] = None, check: bool = True) -> List[torch.Tensor]:
    """Label loop of YOLODataset.__getitem__ (train.py:147-205) for a batch of images.

    labels: per image an (n_i, 5) array-like of raw label lines [class, xc, yc, w, h] (fp64)
    letterbox: per image (orig_w, orig_h, scale, pad_top, pad_left) as returned by the
               reference's letterbox_resize (:136-137); None = an img_size x img_size original.
    Returns dense targets [ (B,G,G,A,5+nc) fp32 on the GPU for each scale ].
    """
    dev = _device()
    B = len(labels)
    S = len(grid_sizes)
    with torch.cuda.device(dev):
        anc = torch.stack([_anchors_dev(a, dev) for a in anchors_list]).contiguous()  # (S,A,2)
        A = anc.shape[1]
        rows = [torch.as_tensor(l, dtype=torch.float64).reshape(-1, 5) for l in labels]
        max_gt = max([r.shape[0] for r in rows], default=0)
        lab = torch.zeros(B, max(max_gt, 1), 5, dtype=torch.float64)
        n_gt = torch.zeros(B, dtype=torch.int32)
        for i, r in enumerate(rows):
            lab[i, :r.shape[0]] = r
            n_gt[i] = r.shape[0]
        if letterbox is None:
            lb = torch.tensor([[img_size, img_size, 1.0, 0.0, 0.0]] * B, dtype=torch.float64).reshape(B, 5)
        else:
            lb = torch.as_tensor(np.asarray(letterbox, dtype=np.float64)).reshape(B, 5)
        lab_d, n_d, lb_d = lab.to(dev), n_gt.to(dev), lb.to(dev)
        row = 5 + num_classes
        targets = [torch.empty(B, g, g, A, row, dtype=torch.float32, device=dev) for g in grid_sizes]
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        tptr = (ctypes.c_void_p * MAX_SCALES)(*[t.data_ptr() for t in targets])
        G = (ctypes.c_int * S)(*[int(g) for g in grid_sizes])
        _lib.check(_lib.lib().yb_build_targets(lab_d.data_ptr(), n_d.data_ptr(), lb_d.data_ptr(), anc.data_ptr(), tptr,
                                               B, max_gt, S, G, A, int(num_classes), int(img_size),
                                               status.data_ptr(), _stream()), "yb_build_targets")
        if check and int(status.item()) != 0:
            raise IndexError("build_targets: a label maps outside the grid or the class range "
                             "(the reference raises IndexError on the same input, train.py:193-205)")
    return targets


# --------------------------------------------------------------------------------------------
# a-6 candidate filter, a-7 NMS
# --------------------------------------------------------------------------------------------
def _heads_desc(preds, ancs, img_size, nc, layout=LAYOUT_BHWAC):
    d = HeadsDesc()
    d.S, d.B, d.A, d.nc = len(preds), preds[0].shape[0], ancs[0].shape[0], nc
    d.img_size = float(img_size)
    d.layout = layout
    for s, p in enumerate(preds):
        d.H[s], d.W[s] = _grid(p, layout)
        d.pred[s] = p.data_ptr()
        d.anchors[s] = ancs[s].data_ptr()
    return d


def filter_candidates(predictions, anchors_list, img_size, num_classes=1, conf_threshold=0.5, letterbox=None,
                      layout=LAYOUT_BHWAC):
    """Per-scale body of predict() (train.py:1152-1229) for a batch: decode, sigmoid, objectness
    filter, class max, pixel xyxy with letterbox reverse, score; P3->P4->P5 order preserved.

    letterbox: (B,3) [scale, pad_top, pad_left] or None.
    Returns (boxes (B,cap,4), scores (B,cap), classes (B,cap) int64, counts (B,) int32) on the GPU,
    cap = rows per image.
    """
    dev = _device()
    S = len(predictions)
    with torch.cuda.device(dev):
        preds = [_f32c(p.detach(), dev) for p in predictions]
        ancs = [_anchors_dev(a, dev) for a in anchors_list[:S]]
        for s, p in enumerate(preds):
            if layout == LAYOUT_NCHW:
                if p.dim() != 4 or p.shape[1] != ancs[s].shape[0] * (5 + int(num_classes)):
                    raise ValueError(f"predictions[{s}]: NCHW head must be (B, A*(5+nc), H, W), got {tuple(p.shape)}")
            else:
                _check_head(p, f"predictions[{s}]")
        B = preds[0].shape[0]
        cap = sum(_grid(p, layout)[0] * _grid(p, layout)[1] * ancs[s].shape[0] for s, p in enumerate(preds))
        d = _heads_desc(preds, ancs, img_size, int(num_classes), layout)
        L = _lib.lib()
        ws_bytes = L.yb_filter_workspace_bytes(ctypes.byref(d))
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
        boxes = torch.empty(B, cap, 4, dtype=torch.float32, device=dev)
        scores = torch.empty(B, cap, dtype=torch.float32, device=dev)
        classes = torch.empty(B, cap, dtype=torch.int64, device=dev)
        counts = (torch.empty if B > 0 else torch.zeros)(B, dtype=torch.int32, device=dev)   # every image's count is written by the kernel
        lb = None
        if letterbox is not None:
            lb = torch.as_tensor(letterbox, dtype=torch.float32).reshape(B, 3).to(dev).contiguous()
        _lib.check(L.yb_filter_compact(ctypes.byref(d), float(conf_threshold), _ptr(lb), boxes.data_ptr(),
                                       scores.data_ptr(), classes.data_ptr(), counts.data_ptr(), cap,
                                       ws.data_ptr(), ws.numel(), _stream()), "yb_filter_compact")
    return boxes, scores, classes, counts


NMS_GRAPH, NMS_BITMASK = _lib.NMS_GRAPH, _lib.NMS_BITMASK
GRAPH_EDGES_PER_BOX = 16  # edge-list capacity of the default workspace (dense random heads need ~3)


def batched_nms_padded(boxes, scores, classes, counts, iou_threshold, trick_max_numel=TRICK_MAX_NUMEL_CUDA,
                       algo=NMS_GRAPH, return_workspace=False):
    """NMS for B images at once.  boxes (B,cap,4), scores (B,cap), classes (B,cap) int64 or None,
    counts (B,) int32 or None.  Returns (keep (B,cap) int64, n_keep (B,) int32) on the GPU, nothing
    synchronised.  The default sparse-graph algorithm resolves an image whose suppression graph does
    not fit the workspace on the device, in the same launch (blocked greedy pass), so n_keep is never
    negative; only the dense bitmask algorithm with an undersized workspace can report n_keep = -1,
    which `pack_detections` turns into a total of -1 and `detections_to_lists` into an exception."""
    dev = boxes.device
    B, cap = boxes.shape[0], boxes.shape[1]
    L = _lib.lib()
    keep = torch.empty(B, cap, dtype=torch.int64, device=dev)
    if B == 0 or cap == 0:
        z = torch.zeros(B, dtype=torch.int32, device=dev)
        return (keep, z, None) if return_workspace else (keep, z)
    n_keep = torch.empty(B, dtype=torch.int32, device=dev)   # written for every image (no fill kernel on the critical path)
    need = (L.yb_nms_graph_workspace_bytes(B, cap, GRAPH_EDGES_PER_BOX) if algo == NMS_GRAPH
            else L.yb_nms_workspace_bytes(B, cap))
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    _lib.check(L.yb_batched_nms(boxes.data_ptr(), scores.data_ptr(), _ptr(classes), _ptr(counts), B, cap,
                                float(iou_threshold), int(trick_max_numel), int(algo), keep.data_ptr(),
                                n_keep.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "yb_batched_nms")
    if return_workspace:
        return keep, n_keep, ws
    return keep, n_keep


def nms_graph_stats(det):
    """(pair tests executed by the half-precision filter, edges, pairs decided exactly in fp32) of the graph NMS
    behind a `detect_batch` result, summed over the batch; None for the bitmask algorithm.  Synchronises.
    For bench.py / tests."""
    ws = det.get("nms_ws")
    if ws is None or det.get("algo", NMS_GRAPH) != NMS_GRAPH:
        return None
    B, cap = det["boxes"].shape[0], det["boxes"].shape[1]
    ev, ed, ca = ctypes.c_ulonglong(0), ctypes.c_ulonglong(0), ctypes.c_ulonglong(0)
    with torch.cuda.device(ws.device):
        _lib.check(_lib.lib().yb_nms_graph_stats(ws.data_ptr(), ws.numel(), B, cap, ctypes.byref(ev), ctypes.byref(ed),
                                                 ctypes.byref(ca), _stream()), "yb_nms_graph_stats")
    return int(ev.value), int(ed.value), int(ca.value)


def nms_retry_overflow(boxes, scores, classes, counts, iou_threshold, trick_max_numel, keep, n_keep):
    """Re-run, with the dense bitmask algorithm, the images the graph algorithm gave up on
    (n_keep == -1).  Synchronises (reads n_keep).  Returns (keep, n_keep), updated in place."""
    bad = (n_keep < 0).nonzero().flatten()
    if bad.numel() == 0:
        return keep, n_keep
    sel = lambda t: None if t is None else t.index_select(0, bad).contiguous()
    k2, n2 = batched_nms_padded(sel(boxes), sel(scores), sel(classes), sel(counts), iou_threshold, trick_max_numel,
                                algo=NMS_BITMASK)
    if bool((n2 < 0).any()):
        raise ValueError("batched_nms: class ids must be in [0, 65536)")
    keep.index_copy_(0, bad, k2)
    n_keep.index_copy_(0, bad, n2)
    return keep, n_keep


def _nms_single(boxes, scores, idxs, iou_threshold, algo=NMS_GRAPH):
    if boxes.dim() != 2 or boxes.shape[1] != 4:
        raise ValueError("boxes must be (N,4)")
    dev = _device()
    orig_dev = boxes.device
    n = boxes.shape[0]
    if n == 0:
        return torch.empty(0, dtype=torch.int64, device=orig_dev)
    with torch.cuda.device(dev):
        b = _f32c(boxes.detach(), dev).unsqueeze(0)
        s = _f32c(scores.detach(), dev).unsqueeze(0)
        c = None
        if idxs is not None:
            c = idxs.detach().to(dev, torch.int64).contiguous().unsqueeze(0)
        # torchvision picks its algorithm from the caller's device (boxes.py:80)
        trick = TRICK_MAX_NUMEL_CPU if orig_dev.type == "cpu" else TRICK_MAX_NUMEL_CUDA
        keep, n_keep = batched_nms_padded(b, s, c, None, iou_threshold, trick, algo)
        k = int(n_keep[0].item())
        if k < 0:
            keep, n_keep = nms_retry_overflow(b, s, c, None, iou_threshold, trick, keep, n_keep)
            k = int(n_keep[0].item())
        out = keep[0, :k].clone()
    return out if orig_dev == dev else out.to(orig_dev)


def batched_nms(boxes, scores, idxs, iou_threshold, algo=NMS_GRAPH):
    """Drop-in for torchvision.ops.batched_nms (train.py:1232-1233): int64 indices of the kept
    boxes in descending score order."""
    return _nms_single(boxes, scores, idxs, iou_threshold, algo)


def nms(boxes, scores, iou_threshold, algo=NMS_GRAPH):
    """Drop-in for torchvision.ops.nms."""
    return _nms_single(boxes, scores, None, iou_threshold, algo)


def detect_batch(predictions, anchors_list, img_size, num_classes=1, conf_threshold=0.5, iou_threshold=0.4,
                 letterbox=None, trick_max_numel=TRICK_MAX_NUMEL_CUDA, algo=NMS_GRAPH, layout=LAYOUT_BHWAC):
    """predict() lines 1152-1238 for a whole batch on the GPU: decode + filter + global NMS.

    Returns a dict of device tensors: boxes (B,cap,4), scores, classes, counts (candidates per
    image), keep (B,cap) indices into the candidates in descending score order, n_keep (B,).
    Nothing is synchronised; use `detections_to_lists` for the reference's list-of-tuples form.
    """
    boxes, scores, classes, counts = filter_candidates(predictions, anchors_list, img_size, num_classes,
                                                       conf_threshold, letterbox, layout)
    return nms_candidates(boxes, scores, classes, counts, iou_threshold, trick_max_numel, algo)


def nms_candidates(boxes, scores, classes, counts, iou_threshold=0.4, trick_max_numel=TRICK_MAX_NUMEL_CUDA, algo=NMS_GRAPH):
    """Second half of detect_batch: global NMS (train.py:1232-1233) over the padded candidates of filter_candidates."""
    with torch.cuda.device(boxes.device):
        keep, n_keep, ws = batched_nms_padded(boxes, scores, classes, counts, iou_threshold, trick_max_numel, algo,
                                              return_workspace=True)
    return {"boxes": boxes, "scores": scores, "classes": classes, "counts": counts, "keep": keep, "n_keep": n_keep,
            "iou_threshold": float(iou_threshold), "trick_max_numel": int(trick_max_numel), "algo": algo, "nms_ws": ws}


def detect_batch_nchw(raw_heads, anchors_list, img_size, num_classes=1, conf_threshold=0.5, iou_threshold=0.4,
                      letterbox=None, trick_max_numel=TRICK_MAX_NUMEL_CUDA, algo=NMS_GRAPH):
    """detect_batch on the head convs' own outputs (B, A*(5+nc), H, W); same candidates,
    same order, same keep sets as detect_batch on the permuted heads."""
    return detect_batch(raw_heads, anchors_list, img_size, num_classes, conf_threshold, iou_threshold, letterbox,
                        trick_max_numel, algo, layout=LAYOUT_NCHW)


def pack_detections(det):
    """Device-side detection list: (rows (B*cap,6) fp32 [x1,y1,x2,y2,conf,class], offsets (B+1,) int32).
    Only the first offsets[B] rows are meaningful."""
    boxes, keep = det["boxes"], det["keep"]
    B, cap = boxes.shape[0], boxes.shape[1]
    with torch.cuda.device(boxes.device):
        rows = torch.empty(B * cap, 6, dtype=torch.float32, device=boxes.device)
        offsets = (torch.empty if B > 0 and cap > 0 else torch.zeros)(B + 1, dtype=torch.int32, device=boxes.device)
        _lib.check(_lib.lib().yb_pack_detections(boxes.data_ptr(), det["scores"].data_ptr(), det["classes"].data_ptr(),
                                                 keep.data_ptr(), det["n_keep"].data_ptr(), B, cap, rows.data_ptr(),
                                                 offsets.data_ptr(), _stream()), "yb_pack_detections")
    return rows, offsets


def detections_to_lists(det):
    """[(x1, y1, x2, y2, conf, class_id), ...] per image (train.py:1242-1246): one pack kernel and
    two D2H copies instead of K*6 `.item()` syncs."""
    rows, offsets = pack_detections(det)
    off = offsets.cpu().tolist()
    if off[-1] < 0:
        raise RuntimeError("NMS failed for an image (n_keep < 0: bitmask algorithm with too small a workspace)")
    host = rows[:off[-1]].cpu().tolist()
    return [[(r[0], r[1], r[2], r[3], r[4], int(r[5])) for r in host[off[b]:off[b + 1]]] for b in range(len(off) - 1)]


def predict_heads(preds, anchors_list, img_size, num_classes=1, conf_threshold=0.5, iou_threshold=0.4, letterbox=None):
    """Everything predict() does after the model call (train.py:1140-1246), for a batch of B images: decode,
    objectness filter, class max, pixel xyxy, letterbox reverse, global NMS, detection list.  preds are the
    model's heads [(B,G,G,A,5+nc)], letterbox per image (scale, pad_top, pad_left) as returned by the
    reference's letterbox_resize.  Returns, per image, [(x1, y1, x2, y2, conf, class_id), ...] in original
    image coordinates, descending score — the reference's return value, one host synchronisation per batch
    instead of a `nonzero` per scale (:1170) and K x 6 `.item()` calls (:1242-1246) per image."""
    # torchvision picks batched_nms' algorithm from the device its inputs live on (boxes.py:80); in
    # predict() that is the device the model ran on
    on_cpu = preds[0].device.type == "cpu"
    det = detect_batch(list(preds), anchors_list, img_size, num_classes, conf_threshold, iou_threshold, letterbox=letterbox,
                       trick_max_numel=TRICK_MAX_NUMEL_CPU if on_cpu else TRICK_MAX_NUMEL_CUDA)
    return detections_to_lists(det)


# --------------------------------------------------------------------------------------------
# the whole step as one CUDA graph
# --------------------------------------------------------------------------------------------
class HotPathGraph:
    """loss forward+backward + decode/filter/global NMS + detection packing for a fixed batch shape,
    captured once as a CUDA graph (about twenty kernel launches, memsets and workspace allocations
    replayed with a single launch; the per-step host cost drops from ~1 ms of Python to microseconds).

    Static inputs (fill them, e.g. with `copy_` from pinned host memory, then call `replay()`):
      heads[s]   (B,G,G,A,5+nc), or (B,A*(5+nc),G,G) with layout=LAYOUT_NCHW
      labels     PackedLabels (targets="labels": assignment on the device), or
      targets[s] dense (B,G,G,A,5+nc) (targets="dense": the reference's call signature)
    Static outputs: losses (4,) [total, bbox, obj, cls], grads[s] (gradient of `total`), det (the
    detect_batch dict), rows (B*cap,6) + offsets (B+1,) from pack_detections.
    An image whose NMS graph overflows the edge list is resolved exactly inside the same launch."""

    def __init__(self, batch, img_size, num_classes, anchors_list, conf_threshold=0.5, iou_threshold=0.4, max_gt=50,
                 targets="labels", layout=LAYOUT_BHWAC, num_anchors=3, device=None, adopt_heads=None,
                 adopt_targets=None, group=None, overlap_loss=True, fork_loss="auto"):
        """adopt_heads / adopt_targets: existing device tensors (or a PackedLabels) to use as the static
        inputs instead of allocating new ones.  group: torch.distributed process group of an image-sharded
        job; the all-reduce of the loss partials is then captured inside the graph (NCCL).
        overlap_loss: the loss kernels (HBM-bound, few resident warps) and the filter/NMS kernels (issue-bound)
        are independent readers of the heads; they are captured as two parallel branches of the graph."""
        self.group = group
        self.overlap_loss = bool(overlap_loss)
        if fork_loss not in ("auto", "start", "after_filter"):
            raise ValueError("fork_loss must be 'auto', 'start' or 'after_filter'")
        self.fork_loss = fork_loss   # "auto" is resolved below, once the kind of targets is known
        dev = _device() if device is None else torch.device(device)
        self.device, self.nc, self.img, self.layout = dev, int(num_classes), int(img_size), layout
        row = 5 + self.nc
        grids = [img_size // 8, img_size // 16, img_size // 32]
        with torch.cuda.device(dev):
            self.anchors = [_anchors_dev(a, dev) for a in anchors_list]
            A = num_anchors
            if adopt_heads is not None:
                self.heads = list(adopt_heads)
            elif layout == LAYOUT_NCHW:
                self.heads = [torch.zeros(batch, A * row, g, g, device=dev) for g in grids]
            else:
                self.heads = [torch.zeros(batch, g, g, A, row, device=dev) for g in grids]
            self.labels = self.targets = None
            if adopt_targets is not None:
                if isinstance(adopt_targets, PackedLabels):
                    self.labels = adopt_targets
                else:
                    self.targets = list(adopt_targets)
            elif targets == "labels":
                self.labels = PackedLabels(torch.zeros(batch, max(max_gt, 1), 5, dtype=torch.float64, device=dev),
                                           torch.zeros(batch, dtype=torch.int32, device=dev),
                                           torch.tensor([[img_size, img_size, 1.0, 0.0, 0.0]] * batch, dtype=torch.float64,
                                                        device=dev), img_size)
            elif targets == "dense":
                self.targets = [torch.zeros(batch, g, g, A, row, device=dev) for g in grids]
            else:
                raise ValueError("targets must be 'labels' or 'dense'")
            self.conf, self.iou = float(conf_threshold), float(iou_threshold)
            if self.fork_loss == "auto":
                # Where the loss branch leaves the detection chain.  Short rows (nc <= 3) with dense targets: the loss
                # kernels take about as long as the two one-CTA-per-image kernels after the filter, so they start there
                # and the filter has the memory system to itself (B200, nc=1: 295.5 -> 289.2 us per step).  Long rows
                # or label-list targets (assignment kernels first): the loss is the longer branch, would still be
                # running when the persistent edge kernel takes every SM, and starts with the filter instead
                # (nc=80: 572 against 578 us; nc=1 with labels: 291.6 against 294.0 us).
                self.fork_loss = "after_filter" if (5 + self.nc <= 8 and self.labels is None) else "start"
            self._branch = torch.cuda.Stream(device=dev) if self.overlap_loss else None
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):  # warm-up outside the capture (function attributes, allocator pools)
                    self._step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.losses, self.grads, self.det, self.rows, self.offsets = self._step()

    def _loss(self):
        out4, _, grads = loss_forward_backward(self.heads, self.targets, self.anchors, self.nc, MULTISCALE_OBJ_WEIGHTS,
                                               [True] * len(self.heads), sparse=self.labels, layout=self.layout,
                                               group=self.group)
        return out4, grads

    def _step(self):
        cur = torch.cuda.current_stream()
        fork_late = self._branch is not None and self.fork_loss == "after_filter"
        if self._branch is not None and not fork_late:   # fork: loss on a second stream, joined before the step ends
            self._branch.wait_stream(cur)
            with torch.cuda.stream(self._branch):
                out4, grads = self._loss()
        elif self._branch is None:
            out4, grads = self._loss()
        cand = filter_candidates(self.heads, self.anchors, self.img, self.nc, self.conf, None, self.layout)
        if fork_late:
            # the filter and the loss both stream the heads from HBM, while the two kernels after the filter run
            # one CTA per image: forking here gives the filter the whole memory system and the loss the idle SMs
            self._branch.wait_stream(cur)
            with torch.cuda.stream(self._branch):
                out4, grads = self._loss()
        det = nms_candidates(*cand, self.iou)
        rows, offsets = pack_detections(det)
        if self._branch is not None:
            cur.wait_stream(self._branch)
        return out4, grads, det, rows, offsets

    def replay(self):
        """Enqueue the captured step on the current stream."""
        self.graph.replay()
        return self.losses, self.grads, self.det
