"""Drop-in boundary: swap the B200 path in under the reference's own names.

The reference (`train.py`) has no plugin API; its callers resolve `decode_predictions`,
`ciou_loss`, `yolo_loss`, `yolo_loss_multiscale` through module globals at call time
(train.py:796, :818, :874, :909, :987, :993, :1154), the CLI resolves `eval_epoch` the same way
(:1413, :1436, :1523), `predict` imports
`torchvision.ops.batched_nms` at call time (:1232), and target assignment lives in
`YOLODataset.__getitem__` / `compute_anchor_iou` (:108-131, :147-205).  `install(train_module)`
rebinds exactly those names; the file on disk is untouched.

  import train; from yolo_from_scratch_b200.install import install; install(train)
  pytest -p yolo_from_scratch_b200.pytest_plugin /path/to/reference/tests     # unchanged suite
  python -m yolo_from_scratch_b200.run /path/to/train.py data.yaml            # unchanged CLI
"""
import importlib.abc
import importlib.machinery
import sys

from . import ops

PATCHED_FUNCTIONS = ("decode_predictions", "ciou_loss", "yolo_loss", "yolo_loss_multiscale")
_MARK = "__yolo_b200_installed__"


def looks_like_reference(module) -> bool:
    return all(hasattr(module, n) for n in PATCHED_FUNCTIONS + ("YOLODataset", "predict"))


def _make_getitem(train_module):
    """Replacement for YOLODataset.__getitem__ (train.py:133-207): image loading and letterbox stay
    the reference's (host image I/O is out of scope); the label loop runs in yb_build_targets."""
    def __getitem__(self, idx):
        import numpy as np
        import torch
        g = train_module.__dict__
        pil_img = g["Image"].open(self.imgs[idx]).convert("RGB")                      # :135
        orig_w, orig_h = pil_img.size
        pil_img, scale, pad_top, pad_left = g["letterbox_resize"](pil_img, self.img_size)   # :137
        img = torch.from_numpy(np.array(pil_img)).permute(2, 0, 1).float() / 255.0    # :138
        rows = []
        label_path = g["Path"](self.labels[idx])
        if label_path.exists():                                                       # :148-154
            with open(label_path, encoding="utf-8") as f:
                for line in f:
                    parts = line.strip().split()
                    if len(parts) == 5:
                        rows.append([float(int(float(parts[0])))] + [float(x) for x in parts[1:]])
        targets = ops.build_targets([np.array(rows, dtype=np.float64).reshape(-1, 5)], self.anchors, self.grid_sizes,
                                    self.num_classes, self.img_size,
                                    letterbox=[(orig_w, orig_h, scale, pad_top, pad_left)])
        return img, [t[0].cpu() for t in targets]
    return __getitem__


def install(train_module, patch_torchvision=True, patch_dataset=True, patch_eval=True):
    """Rebind the hot-path names of an imported reference `train` module.  Idempotent.
    patch_eval also swaps `eval_epoch` (its per-anchor python loop becomes one kernel)."""
    if getattr(train_module, _MARK, False):
        return train_module
    saved = {}
    for name in PATCHED_FUNCTIONS:
        saved[name] = getattr(train_module, name)
        setattr(train_module, name, getattr(ops, name))
    if patch_eval and hasattr(train_module, "eval_epoch"):
        saved["eval_epoch"] = train_module.eval_epoch
        train_module.eval_epoch = ops.eval_epoch
    if patch_dataset and hasattr(train_module, "YOLODataset"):
        cls = train_module.YOLODataset
        saved["YOLODataset.__getitem__"] = cls.__getitem__
        saved["YOLODataset.compute_anchor_iou"] = cls.compute_anchor_iou
        cls.__getitem__ = _make_getitem(train_module)
        cls.compute_anchor_iou = lambda self, box_wh, anchors: ops.compute_anchor_iou(box_wh, anchors)
    if patch_torchvision:
        import torchvision
        import torchvision.ops.boxes as tvb
        saved["torchvision.ops.batched_nms"] = torchvision.ops.batched_nms
        torchvision.ops.batched_nms = ops.batched_nms
        tvb.batched_nms = ops.batched_nms
    setattr(train_module, "__yolo_b200_saved__", saved)
    setattr(train_module, _MARK, True)
    return train_module


def uninstall(train_module):
    saved = getattr(train_module, "__yolo_b200_saved__", None)
    if not saved:
        return
    for name in PATCHED_FUNCTIONS:
        setattr(train_module, name, saved[name])
    if "eval_epoch" in saved:
        train_module.eval_epoch = saved["eval_epoch"]
    if "YOLODataset.__getitem__" in saved:
        train_module.YOLODataset.__getitem__ = saved["YOLODataset.__getitem__"]
        train_module.YOLODataset.compute_anchor_iou = saved["YOLODataset.compute_anchor_iou"]
    if "torchvision.ops.batched_nms" in saved:
        import torchvision
        import torchvision.ops.boxes as tvb
        torchvision.ops.batched_nms = saved["torchvision.ops.batched_nms"]
        tvb.batched_nms = saved["torchvision.ops.batched_nms"]
    setattr(train_module, _MARK, False)


class _PatchingLoader(importlib.abc.Loader):
    def __init__(self, inner):
        self.inner = inner

    def create_module(self, spec):
        return self.inner.create_module(spec)

    def exec_module(self, module):
        self.inner.exec_module(module)
        if looks_like_reference(module):
            install(module)


class _TrainFinder(importlib.abc.MetaPathFinder):
    """Patches a module called `train` right after it is executed, so that
    `from train import yolo_loss_multiscale` in a test module already binds the B200 path."""

    def find_spec(self, fullname, path, target=None):
        if fullname != "train":
            return None
        spec = importlib.machinery.PathFinder.find_spec(fullname, path)
        if spec is None or spec.loader is None:
            return None
        spec.loader = _PatchingLoader(spec.loader)
        return spec


_finder = None


def enable_import_hook():
    global _finder
    if _finder is None:
        _finder = _TrainFinder()
        sys.meta_path.insert(0, _finder)
    mod = sys.modules.get("train")
    if mod is not None and looks_like_reference(mod):
        install(mod)


def disable_import_hook():
    global _finder
    if _finder is not None and _finder in sys.meta_path:
        sys.meta_path.remove(_finder)
    _finder = None
