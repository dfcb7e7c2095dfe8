"""ORACLE (test infrastructure): CPU restatement of the reference's per-box hot path.

Every function cites the reference lines it restates (paths relative to /root/reference).  The
arithmetic is kept in the reference's evaluation order so fp32 results agree with the reference
run on the same device; structure and naming are this repo's own.  torch fp32 ops are used for the
floating-point stages (decode, losses), numpy fp32/fp64 for target assignment, and `nms_ref.c`
for torchvision's NMS arithmetic.

Parity pinned — see oracle/__init__.py and tests/test_oracle_cpu.py.
"""
import ctypes
import math
import os
import subprocess

import numpy as np
import torch

__all__ = [
    "decode", "ciou", "single_scale_loss", "multiscale_loss", "shape_iou", "assign_targets",
    "eval_counts", "eval_metrics", "candidates", "nms_indices", "batched_nms_indices", "detect", "python_list_nms", "corner_iou",
    "default_anchors", "MULTISCALE_OBJ_WEIGHTS",
]

MULTISCALE_OBJ_WEIGHTS = (4.0, 1.0, 0.4)  # train.py:865


def default_anchors():
    """train.py:372-374 / :81-83."""
    return [torch.tensor(a, dtype=torch.float32) for a in (
        [[10, 13], [16, 30], [33, 23]], [[30, 61], [62, 45], [59, 119]], [[116, 90], [156, 198], [373, 326]])]


# ---------------------------------------------------------------------------------------------
# decode — train.py:712-779
# ---------------------------------------------------------------------------------------------
def decode(raw, anchors, img_size=640):
    _, gh, gw, na, _ = raw.shape
    anchors = torch.as_tensor(anchors, dtype=raw.dtype).to(raw.device)
    col = torch.arange(gw, device=raw.device, dtype=raw.dtype).view(1, 1, gw, 1)   # :741-748
    row = torch.arange(gh, device=raw.device, dtype=raw.dtype).view(1, gh, 1, 1)
    bx = ((torch.sigmoid(raw[..., 0]) * 2.0 - 0.5) + col) / gw                       # :758
    by = ((torch.sigmoid(raw[..., 1]) * 2.0 - 0.5) + row) / gh                       # :759
    aw = anchors[:, 0].view(1, 1, 1, na)
    ah = anchors[:, 1].view(1, 1, 1, na)
    bw = (aw / img_size) * torch.pow(2.0 * torch.sigmoid(raw[..., 2]), 2)            # :773
    bh = (ah / img_size) * torch.pow(2.0 * torch.sigmoid(raw[..., 3]), 2)            # :774
    return torch.cat([torch.stack([bx, by, bw, bh], dim=-1), raw[..., 4:]], dim=-1)  # :737,:777


# ---------------------------------------------------------------------------------------------
# CIoU — train.py:634-710
# ---------------------------------------------------------------------------------------------
def ciou(pred, tgt, eps=1e-7):
    px, py, pw, ph = pred.unbind(dim=1)
    tx, ty, tw, th = tgt.unbind(dim=1)
    p_lo_x, p_lo_y, p_hi_x, p_hi_y = px - pw / 2, py - ph / 2, px + pw / 2, py + ph / 2   # :652-655
    t_lo_x, t_lo_y, t_hi_x, t_hi_y = tx - tw / 2, ty - th / 2, tx + tw / 2, ty + th / 2   # :657-660
    ow = torch.clamp(torch.min(p_hi_x, t_hi_x) - torch.max(p_lo_x, t_lo_x), min=0)        # :663-668
    oh = torch.clamp(torch.min(p_hi_y, t_hi_y) - torch.max(p_lo_y, t_lo_y), min=0)        # :664-669
    overlap = ow * oh                                                                     # :670
    union = pw * ph + tw * th - overlap                                                   # :673-675
    iou = overlap / (union + eps)                                                         # :678
    centre = (px - tx) ** 2 + (py - ty) ** 2                                              # :681
    hull_w = torch.max(p_hi_x, t_hi_x) - torch.min(p_lo_x, t_lo_x)                        # :684-689
    hull_h = torch.max(p_hi_y, t_hi_y) - torch.min(p_lo_y, t_lo_y)                        # :685-690
    hull_diag = hull_w ** 2 + hull_h ** 2 + eps                                           # :691
    dist = centre / hull_diag                                                             # :694
    v = (4 / (torch.pi ** 2)) * torch.pow(
        torch.atan(pw / (ph + eps)) - torch.atan(tw / (th + eps)), 2)                     # :697-699
    with torch.no_grad():
        alpha = v / (1 - iou + v + eps)                                                   # :701-702
    return (1 - (iou - dist - alpha * v)).mean()                                          # :705-710


# ---------------------------------------------------------------------------------------------
# losses — train.py:781-838, :840-886
# ---------------------------------------------------------------------------------------------
def _bce_logits_mean(x, t):
    return torch.nn.functional.binary_cross_entropy_with_logits(x, t)  # nn.BCEWithLogitsLoss() :823,:827


def single_scale_loss(pred, tgt, anchors, num_classes=1):
    dec = decode(pred, anchors)                       # default img_size=640 always (:796)
    positive = tgt[..., 4] > 0.5                      # :809
    zero = torch.tensor(0.0, device=pred.device)
    n_pos = int(positive.sum())
    box = ciou(dec[..., 0:4][positive], tgt[..., 0:4][positive]) if n_pos > 0 else zero    # :812-820
    obj = _bce_logits_mean(pred[..., 4:5], tgt[..., 4:5])                                  # :823
    if n_pos > 0 and num_classes > 0:                                                      # :826-832
        cls = _bce_logits_mean(pred[..., 5:][positive].reshape(-1), tgt[..., 5:][positive].reshape(-1))
    else:
        cls = zero
    total = 0.05 * box + 1.0 * obj + 0.5 * cls                                             # :836
    return total, box, obj, cls


def multiscale_loss(preds, tgts, anchors_list, num_classes=1):
    total = box_sum = obj_sum = cls_sum = 0.0
    for p, t, a, w in zip(preds, tgts, anchors_list, MULTISCALE_OBJ_WEIGHTS):   # :873
        _, box, obj, cls = single_scale_loss(p, t, a, num_classes)
        total = total + (0.05 * box + w * obj + 0.5 * cls)                    # :879-881
        box_sum = box_sum + box
        obj_sum = obj_sum + obj                                               # unweighted (:883)
        cls_sum = cls_sum + cls
    return total, box_sum, obj_sum, cls_sum


# ---------------------------------------------------------------------------------------------
# target assignment — train.py:108-131, :147-205
# ---------------------------------------------------------------------------------------------
def shape_iou(box_wh, anchors):
    """fp32 IoU of one (w,h) against (A,2) anchors sharing a centre (:119-130)."""
    f = np.float32
    w, h = f(box_wh[0]), f(box_wh[1])
    anchors = np.asarray(anchors, dtype=np.float32)
    out = np.empty(len(anchors), dtype=np.float32)
    for k, (aw, ah) in enumerate(anchors):
        inter = f(min(w, aw)) * f(min(h, ah))
        union = f(f(f(w * h) + f(aw * ah)) - inter)
        out[k] = f(inter / f(union + f(1e-16)))
    return out


def assign_targets(labels, anchors_list, grid_sizes, num_classes, img_size, letterbox=None):
    """labels: (n,5) raw [class, xc, yc, w, h]; letterbox = (orig_w, orig_h, scale, pad_top, pad_left).
    Returns [ (G,G,A,5+nc) float32 ndarray per scale ] for ONE image."""
    anchors_list = [np.asarray(a, dtype=np.float32) for a in anchors_list]
    n_anchor = anchors_list[0].shape[0]
    out = [np.zeros((g, g, n_anchor, 5 + num_classes), dtype=np.float32) for g in grid_sizes]   # :141-145
    ow, oh, scale, pad_top, pad_left = letterbox if letterbox is not None else (img_size, img_size, 1.0, 0, 0)
    for line in np.asarray(labels, dtype=np.float64).reshape(-1, 5):
        cls_id = int(float(line[0]))                                          # :153
        xc = (float(line[1]) * ow * scale + pad_left) / img_size              # :159
        yc = (float(line[2]) * oh * scale + pad_top) / img_size               # :160
        bw = (float(line[3]) * ow * scale) / img_size                         # :161
        bh = (float(line[4]) * oh * scale) / img_size                         # :162
        wh_px = (np.float32(bw * img_size), np.float32(bh * img_size))        # :165-167 (torch.tensor -> fp32)
        best, best_s, best_a = -1.0, 0, 0
        for s, anchors in enumerate(anchors_list):                            # :174-180
            ious = shape_iou(wh_px, anchors)
            top = float(ious.max())
            if top > best:
                best, best_s, best_a = top, s, int(ious.argmax())
        g = grid_sizes[best_s]
        gx = min(int(xc * g), g - 1)                                          # :184-189
        gy = min(int(yc * g), g - 1)
        slot = out[best_s][gy, gx, best_a]                                    # python negative indices wrap
        if slot[4] == 0:                                                      # :193
            slot[0:4] = np.array([xc, yc, bw, bh], dtype=np.float32)          # :195-197
            slot[4] = 1.0                                                     # :199
            slot[5 if num_classes == 1 else 5 + cls_id] = 1.0                 # :201-205
    return out


# ---------------------------------------------------------------------------------------------
# candidate filter — train.py:1152-1222, :1227-1229 (one image)
# ---------------------------------------------------------------------------------------------
def candidates(preds, anchors_list, img_size, num_classes=1, conf=0.5, scale=1.0, pad_top=0, pad_left=0):
    """preds: list of (1,H,W,A,5+nc).  Returns boxes (M,4), scores (M,), classes (M,) int64."""
    boxes, scores, classes = [], [], []
    for p, anchors in zip(preds, anchors_list):
        d = decode(p, anchors, img_size)                                      # :1154
        d[..., 4] = torch.sigmoid(p[..., 4])                                  # :1157 (same slicing as the
        if num_classes > 0:                                                   #  reference: the strided and
            d[..., 5:] = torch.sigmoid(p[..., 5:])                            #  vectorised CPU paths differ by an ulp)
        d = d[0]                                                              # :1163
        rows = d[d[..., 4] > conf]                                            # :1166-1174 (row-major order)
        if rows.shape[0] == 0:
            continue
        if num_classes == 1:                                                  # :1184-1189
            prob = rows[:, 5]
            cid = torch.zeros(rows.shape[0], dtype=torch.long, device=rows.device)
        else:
            prob, cid = rows[:, 5:].max(dim=1)
        cx, cy, w, h = (rows[:, k] * img_size for k in range(4))              # :1192-1195
        x1, y1, x2, y2 = cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2       # :1198-1201
        x1, y1, x2, y2 = x1 - pad_left, y1 - pad_top, x2 - pad_left, y2 - pad_top   # :1205-1208
        x1, y1, x2, y2 = x1 / scale, y1 / scale, x2 / scale, y2 / scale       # :1210-1213
        boxes.append(torch.stack([x1, y1, x2, y2], dim=1))                    # :1219
        scores.append(rows[:, 4] * prob)                                      # :1216
        classes.append(cid)
    if not boxes:
        return torch.zeros(0, 4), torch.zeros(0), torch.zeros(0, dtype=torch.long)
    return torch.cat(boxes), torch.cat(scores), torch.cat(classes)            # :1227-1229


# ---------------------------------------------------------------------------------------------
# eval_epoch's detection counting — train.py:993-1024 with compute_box_iou :928-958
# ---------------------------------------------------------------------------------------------
def eval_counts(preds, targets, anchors_list, conf_threshold=0.5, iou_threshold=0.5):
    """(TP, FP, FN) over all scales/images/cells/anchors, as the reference's 4-deep python loop counts
    them.  `.item()` turns fp32 values into python floats, so the confidence tests run in double;
    `iou > iou_threshold` compares an fp32 tensor with a python scalar, i.e. in fp32."""
    tp = fp = fn = 0
    for pred, tgt, anchors in zip(preds, targets, anchors_list):
        dec = decode(pred, anchors)                                    # :993 (img_size defaults to 640)
        p_obj = torch.sigmoid(pred[..., 4]).double()                   # :997, :1003
        t_obj = tgt[..., 4].double()                                   # :1004
        p_on, t_on = p_obj > conf_threshold, t_obj > conf_threshold
        both = p_on & t_on
        if bool(both.any()):
            a, b = dec[..., 0:4][both], tgt[..., 0:4][both]            # :1008-1010
            a_x1, a_y1, a_x2, a_y2 = a[:, 0] - a[:, 2] / 2, a[:, 1] - a[:, 3] / 2, a[:, 0] + a[:, 2] / 2, a[:, 1] + a[:, 3] / 2
            b_x1, b_y1, b_x2, b_y2 = b[:, 0] - b[:, 2] / 2, b[:, 1] - b[:, 3] / 2, b[:, 0] + b[:, 2] / 2, b[:, 1] + b[:, 3] / 2
            inter = torch.clamp(torch.min(a_x2, b_x2) - torch.max(a_x1, b_x1), min=0) * \
                torch.clamp(torch.min(a_y2, b_y2) - torch.max(a_y1, b_y1), min=0)          # :945-950
            union = (a_x2 - a_x1) * (a_y2 - a_y1) + (b_x2 - b_x1) * (b_y2 - b_y1) - inter  # :953-955
            iou = inter / (union + 1e-6)                                                   # :957
            hit = int((iou > iou_threshold).sum())                                         # :1012
            tp += hit
            fp += int(both.sum()) - hit                                                    # :1015
        fp += int((p_on & ~t_on).sum())                                                    # :1016-1018
        fn += int((~p_on & t_on).sum())                                                    # :1019-1021
    return tp, fp, fn


def eval_metrics(tp, fp, fn):
    """precision, recall, F1 in percent — train.py:1026-1032."""
    precision = tp / (tp + fp) if (tp + fp) > 0 else 0
    recall = tp / (tp + fn) if (tp + fn) > 0 else 0
    f1 = 2 * precision * recall / (precision + recall) if (precision + recall) > 0 else 0
    return precision * 100, recall * 100, f1 * 100


# ---------------------------------------------------------------------------------------------
# NMS — torchvision 0.26.0 restated (nms_ref.c) + batched_nms dispatch (ops/boxes.py:51-120)
# ---------------------------------------------------------------------------------------------
_HERE = os.path.dirname(os.path.abspath(__file__))
_NMS_LIB = None


def build_c_oracle():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def _nms_lib():
    global _NMS_LIB
    if _NMS_LIB is None:
        path = os.path.join(_HERE, "liboracle_nms.so")
        if not os.path.exists(path):
            build_c_oracle()
        lib = ctypes.CDLL(path)
        lib.yb_oracle_nms.restype = ctypes.c_int64
        lib.yb_oracle_nms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_double,
                                      ctypes.c_int, ctypes.c_void_p]
        _NMS_LIB = lib
    return _NMS_LIB


def nms_indices(boxes, scores, iou_threshold, arith="cuda"):
    """Indices kept by torchvision.ops.nms, descending score.  arith: 'cuda' | 'cpu'."""
    b = np.ascontiguousarray(np.asarray(boxes, dtype=np.float32).reshape(-1, 4))
    s = np.ascontiguousarray(np.asarray(scores, dtype=np.float32).reshape(-1))
    keep = np.empty(len(s), dtype=np.int64)
    k = _nms_lib().yb_oracle_nms(b.ctypes.data, s.ctypes.data, len(s), float(iou_threshold),
                                 0 if arith == "cuda" else 1, keep.ctypes.data)
    return keep[:k].copy()


def batched_nms_indices(boxes, scores, idxs, iou_threshold, arith="cuda", device_rule="cuda"):
    """torchvision.ops.batched_nms.  device_rule picks the size dispatch of boxes.py:80:
    'cuda' -> coordinate trick up to 100000 box elements, 'cpu' -> up to 4000."""
    b = np.asarray(boxes, dtype=np.float32).reshape(-1, 4)
    s = np.asarray(scores, dtype=np.float32).reshape(-1)
    c = np.asarray(idxs, dtype=np.int64).reshape(-1)
    if b.size == 0:
        return np.empty(0, dtype=np.int64)
    if b.size > (4000 if device_rule == "cpu" else 100_000):
        kept = np.zeros(len(s), dtype=bool)                                   # boxes.py:112-120
        for cls in np.unique(c):
            sel = np.nonzero(c == cls)[0]
            kept[sel[nms_indices(b[sel], s[sel], iou_threshold, arith)]] = True
        idx = np.nonzero(kept)[0]
        return idx[np.argsort(-s[idx], kind="stable")]
    top = b.max()                                                             # boxes.py:98-101
    offs = c.astype(np.float32) * np.float32(top + np.float32(1))
    return nms_indices(b + offs[:, None], s, iou_threshold, arith)


def detect(preds, anchors_list, img_size, num_classes=1, conf=0.5, iou=0.4, scale=1.0, pad_top=0, pad_left=0,
           arith="cuda", device_rule="cuda"):
    """predict() lines 1152-1246 for one image given its heads: list of (x1,y1,x2,y2,conf,cls)."""
    b, s, c = candidates(preds, anchors_list, img_size, num_classes, conf, scale, pad_top, pad_left)
    if b.shape[0] == 0:
        return [], (b, s, c), np.empty(0, dtype=np.int64)
    keep = batched_nms_indices(b.numpy(), s.numpy(), c.numpy(), iou, arith, device_rule)
    dets = [(float(b[i, 0]), float(b[i, 1]), float(b[i, 2]), float(b[i, 3]), float(s[i]), int(c[i])) for i in keep]
    return dets, (b, s, c), keep


# ---------------------------------------------------------------------------------------------
# the reference's pure-python list NMS and corner IoU — train.py:1064-1112 (used by its tests as
# known-answer cases: tests/test_inference.py:16-109)
# ---------------------------------------------------------------------------------------------
def corner_iou(a, b):
    iw = max(0, min(a[2], b[2]) - max(a[0], b[0]))
    ih = max(0, min(a[3], b[3]) - max(a[1], b[1]))
    inter = iw * ih
    union = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter
    return inter / union if union > 0 else 0


def python_list_nms(dets, iou_threshold):
    pending = sorted(dets, key=lambda d: d[4], reverse=True)
    kept = []
    while pending:
        head, pending = pending[0], pending[1:]
        kept.append(head)
        pending = [d for d in pending if corner_iou(head, d) < iou_threshold]   # keeps IoU < thr (:1109-1110)
    return kept
