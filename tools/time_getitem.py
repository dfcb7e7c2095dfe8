"""Times YOLODataset.__getitem__ of the UNMODIFIED reference (baseline/_ref/train.py) against the installed
B200 binding (label loop in yb_build_targets), on a temporary dataset.  VERDICT r1 item 10."""
import os, sys, time, tempfile, importlib.util
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from PIL import Image
spec = importlib.util.spec_from_file_location("train", os.path.join(ROOT, "baseline", "_ref", "train.py"))
train = importlib.util.module_from_spec(spec); spec.loader.exec_module(train)
from yolo_from_scratch_b200 import install as inst

def make(root, n, boxes, nc, rng):
    imgs, labels = os.path.join(root, "images"), os.path.join(root, "labels")
    os.makedirs(imgs); os.makedirs(labels)
    for i in range(n):
        Image.fromarray(rng.integers(0, 255, (480, 640, 3), dtype=np.uint8)).save(os.path.join(imgs, f"i{i}.jpg"))
        with open(os.path.join(labels, f"i{i}.txt"), "w") as f:
            for _ in range(boxes):
                f.write(f"{int(rng.integers(0, nc))} {rng.uniform(.2,.8)} {rng.uniform(.2,.8)} {rng.uniform(.05,.3)} {rng.uniform(.05,.3)}\n")
    return imgs

rng = np.random.default_rng(0)
for nc, boxes in ((1, 3), (1, 30), (80, 30)):
    with tempfile.TemporaryDirectory() as tmp:
        d = make(tmp, 8, boxes, nc, rng)
        ds = train.YOLODataset(d, num_classes=nc, img_size=640)
        def run(n=3):
            for i in range(len(ds)): ds[i]
            t0 = time.perf_counter()
            for _ in range(n):
                for i in range(len(ds)): ds[i]
            return (time.perf_counter() - t0) / (n * len(ds)) * 1e3
        ref_ms = run()
        inst.install(train)
        b200_ms = run()
        inst.uninstall(train)
        print(f"nc={nc} boxes/image={boxes}: reference __getitem__ {ref_ms:.2f} ms/sample, B200 binding {b200_ms:.2f} ms/sample (both include JPEG decode + letterbox)")
