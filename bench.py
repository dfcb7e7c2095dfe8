#!/usr/bin/env python
"""bench.py — decode+global-NMS images/s and CIoU-loss fwd+bwd ms @640^2 bs64 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the per-box hot path over one batch of synthetic heads/labels:
fused multi-scale loss forward+backward, then decode + confidence filter + global NMS.
Workload = BASELINE.json configs[1]: nc=1, 640x640, 64 images per GPU, <=50 GT boxes per image,
randn heads (seed 1234+scale), predict()'s default thresholds (conf 0.5, IoU 0.4; train.py:1114).
N GPUs = 64 images on each (weak scaling, sharded by image); the loss partial sums are all-reduced
once per step over NCCL, NMS needs no collective.

`value` is timed with the inputs resident in HBM; `e2e` goes through the public Python API with
pinned HOST buffers (H2D of heads+targets and D2H of losses + detections inside the timed region).
`--impl reference` times the oracle port of the reference's CPU implementation (torch CPU ops +
torchvision's CPU nms, the code path train.py runs on a host without CUDA) on a bounded sample.
One JSON line is printed by rank 0.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

IMG, NC, B_PER_GPU, MAX_GT = 640, 1, 64, 50
CONF, IOU = 0.5, 0.4
METRIC = "decode+global-NMS + CIoU-loss fwd+bwd throughput @640^2 bs64 nc1"
UNIT = "images/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nc", type=int, default=NC)
    ap.add_argument("--img", type=int, default=IMG)
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="images per GPU")
    ap.add_argument("--conf", type=float, default=CONF)
    ap.add_argument("--iou", type=float, default=IOU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--launch", default="graph", choices=["graph", "eager"],
                    help="what `value` times: the step replayed as a CUDA graph (yb.HotPathGraph; at N>1 the NCCL "
                         "all-reduce is captured inside the graph), or eager launches.  Per-kernel times always come "
                         "from an eager pass; both throughputs are reported.")
    ap.add_argument("--graph-multi-gpu", action="store_true", help=argparse.SUPPRESS)  # kept for old command lines
    ap.add_argument("--layout", default="bhwac", choices=["bhwac", "nchw"],
                    help="head layout of the device-resident step: the reference's (B,H,W,A,5+nc) or the head conv's own "
                         "NCHW output (SURVEY 8f-2)")
    ap.add_argument("--targets", default="dense", choices=["dense", "labels"],
                    help="dense reference targets, or label lists assigned on the device (SURVEY 8f-4)")
    ap.add_argument("--no-torch-gpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic workload
# ------------------------------------------------------------------------------------------------
def make_labels(rng, B, nc):
    labels = []
    for _ in range(B):
        n = int(rng.integers(0, MAX_GT + 1))
        lab = np.zeros((n, 5), dtype=np.float64)
        lab[:, 0] = rng.integers(0, nc, size=n)
        lab[:, 1:3] = rng.uniform(0.05, 0.95, size=(n, 2))
        lab[:, 3:5] = np.exp(rng.uniform(np.log(0.01), np.log(0.6), size=(n, 2)))
        labels.append(lab)
    return labels


def make_heads(B, img, nc, seed):
    grids = [img // 8, img // 16, img // 32]
    heads = []
    for s, G in enumerate(grids):
        g = torch.Generator().manual_seed(seed + s)
        heads.append(torch.randn(B, G, G, 3, 5 + nc, generator=g))
    return heads


def tensor_bytes(ts):
    return int(sum(t.numel() * t.element_size() for t in ts))


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvml)
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.stop_flag, self.max_mhz, self.ok = [], set(), False, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: oracle port of the reference's CPU path
# ------------------------------------------------------------------------------------------------
def cpu_reference_step(heads, tgts, anchors, nc, img, conf, iou, pool, use_tv):
    """loss fwd+bwd (torch CPU, all threads) + per-image decode/filter/NMS (train.py:1152-1238)."""
    from oracle import ref_path as R
    preds = [h.clone().requires_grad_(True) for h in heads]
    total = R.multiscale_loss(preds, tgts, anchors, nc)[0]
    total.backward()
    B = heads[0].shape[0]

    def one(b):
        bx, sc, cl = R.candidates([h[b:b + 1] for h in heads], anchors, img, nc, conf)
        if bx.shape[0] == 0:
            return 0
        if use_tv:
            import torchvision
            return int(torchvision.ops.batched_nms(bx, sc, cl, iou).numel())
        return len(R.batched_nms_indices(bx.numpy(), sc.numpy(), cl.numpy(), iou, "cpu", "cpu"))
    kept = list(pool.map(one, range(B)))
    return float(total), kept


def run_cpu_reference(args, sample_images, steps, warmup):
    from concurrent.futures import ThreadPoolExecutor
    from oracle import ref_path as R
    try:
        import torchvision  # noqa: F401
        use_tv = True
    except Exception:
        use_tv = False
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    anchors = R.default_anchors()
    heads = [h[:sample_images].contiguous() for h in make_heads(args.batch, args.img, args.nc, 1234)]
    labels = make_labels(np.random.default_rng(4321), args.batch, args.nc)[:sample_images]
    grids = [args.img // 8, args.img // 16, args.img // 32]
    tg = [R.assign_targets(l, anchors, grids, args.nc, args.img) for l in labels]
    tgts = [torch.from_numpy(np.stack([t[s] for t in tg])) for s in range(3)]
    pool = ThreadPoolExecutor(max_workers=min(cores, sample_images))
    for _ in range(warmup):
        cpu_reference_step(heads, tgts, anchors, args.nc, args.img, args.conf, args.iou, pool, use_tv)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(heads, tgts, anchors, args.nc, args.img, args.conf, args.iou, pool, use_tv)
    dt = (time.perf_counter() - t0) / steps
    pool.shutdown()
    sample = (f"{sample_images} of {args.batch} images per step (same seeds), loss fwd+bwd on torch CPU + per-image "
              f"decode/filter/{'torchvision CPU batched_nms' if use_tv else 'nms_ref.c'}; images spread over {min(cores, sample_images)} threads")
    return sample_images / dt, dt * 1e3, cores, sample


def reference_main(args, rank):
    if rank != 0:
        return
    sample_images = max(1, min(args.batch, 8))
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    v, ms, cores, sample = run_cpu_reference(args, sample_images, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world, n_sets=4):
    return {
        "head_layout": args.layout, "targets": args.targets,
        "workload": f"BASELINE {'configs[1]' if (args.nc, args.img, args.batch) == (1, 640, 64) else 'variant'}: nc={args.nc} heads at {args.img}x{args.img}, {args.batch} images/GPU, "
                    f"<= {MAX_GT} GT boxes/image, randn heads, conf {args.conf}, iou {args.iou}",
        "global_batch": args.batch * world, "images_per_gpu": args.batch, "img_size": args.img, "nc": args.nc,
        "conf_thres": args.conf, "iou_thres": args.iou, "parallelism": f"image-sharded x{world}",
        "l2": f"inputs rotate over {n_sets} sets per rank (working set > 126 MB L2)",
    }


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def b200_main(args, rank, local_rank, world):
    import yolo_from_scratch_b200 as yb
    lib = yb._lib.lib()  # raises if the CUDA extension is missing: no fallback
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl b200) needs a GPU; the CUDA path has no fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        import torch.distributed as dist
        import datetime
        # a short collective timeout: a rank-asymmetric bug must fail in minutes, not hold N GPUs for the default 10
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=90))
        group = dist.group.WORLD
    line = run_workload(args, rank, local_rank, world, dev, group, lib, full=True)
    if rank == 0 and not args.no_other_configs and world == 1:
        # the other single-GPU BASELINE configs, device-resident only, short runs (parity is in tests/)
        import copy
        others = {}
        for name, (nc, img, batch, conf, layout, targets) in OTHER_CONFIGS.items():
            a2 = copy.copy(args)
            a2.nc, a2.img, a2.batch, a2.conf, a2.layout, a2.targets = nc, img, batch, conf, layout, targets
            a2.steps, a2.warmup = max(3, min(args.steps, 6)), 3
            try:
                o = run_workload(a2, rank, local_rank, world, dev, group, lib, full=False)
                others[name] = {k: o[k] for k in ("value", "ms_per_step", "loss_fwd_bwd_ms", "decode_nms_ms",
                                                  "candidates_per_image", "kept_per_image", "kernels", "hbm_kernels",
                                                  "roofline", "config")}
            except Exception as e:  # pragma: no cover
                others[name] = {"error": repr(e)}
            torch.cuda.empty_cache()
        line["other_configs"] = others
    if rank == 0:
        print(json.dumps(line), flush=True)


OTHER_CONFIGS = {  # BASELINE.json configs[2], configs[3]: (nc, img, images/GPU, conf, head layout, targets)
    "configs[2] nc80 640^2 B64 conf0.001": (80, 640, 64, 0.001, "bhwac", "dense"),
    "configs[3] nc80 1280^2 B32 conf0.001": (80, 1280, 32, 0.001, "bhwac", "dense"),
    # the same work on the 'next' rows of SURVEY 8f: NCHW conv output read directly + label-list targets
    "configs[1] nc1 640^2 B64 conf0.5, NCHW heads + label lists": (1, 640, 64, 0.5, "nchw", "labels"),
    "configs[2] nc80 640^2 B64 conf0.001, NCHW heads + label lists": (80, 640, 64, 0.001, "nchw", "labels"),
}


def run_workload(args, rank, local_rank, world, dev, group, lib, full):
    import yolo_from_scratch_b200 as yb
    from yolo_from_scratch_b200 import ops
    B, img, nc = args.batch, args.img, args.nc
    # rotate over enough input sets that the working set exceeds the 126 MB L2
    T_est = sum(B * G * G * 3 for G in (img // 8, img // 16, img // 32)) * (5 + nc) * 4
    n_sets = 4 if T_est * 3 * 4 <= (8 << 30) else 2
    grids = [img // 8, img // 16, img // 32]
    anchors = ops.default_anchors(dev)  # train.py:372-374
    weights = ops.MULTISCALE_OBJ_WEIGHTS

    # n_sets input sets per rank, on the device and (full run) mirrored in pinned host memory
    layout = ops.LAYOUT_NCHW if args.layout == "nchw" else ops.LAYOUT_BHWAC
    use_labels = args.targets == "labels"
    dev_sets, host_sets, label_sets = [], [], []
    for k in range(n_sets):
        seed = 1234 + 1000 * k + 100000 * rank
        heads = make_heads(B, img, nc, seed)
        labels = make_labels(np.random.default_rng(4321 + k + 1000 * rank), B, nc)
        tg = ops.build_targets(labels, anchors, grids, nc, img)
        d_heads = [h.to(dev) for h in heads]
        if layout == ops.LAYOUT_NCHW:  # the same values as the head conv would have produced them
            d_heads = [h.permute(0, 3, 4, 1, 2).reshape(h.shape[0], -1, h.shape[1], h.shape[2]).contiguous()
                       for h in d_heads]
        packed = ops.pack_labels_host(labels, img, max_gt=MAX_GT)
        packed = ops.PackedLabels(packed[0].to(dev), packed[1].to(dev), packed[2].to(dev), img)
        dev_sets.append((d_heads, tg, packed))
        if full:
            host_sets.append(([h.pin_memory() for h in heads], [t.cpu().pin_memory() for t in tg]))
            label_sets.append(ops.pack_labels_host(labels, img, pin=True, max_gt=MAX_GT))
        del heads
    torch.cuda.synchronize()
    T_bytes = tensor_bytes(dev_sets[0][0])
    rows = sum(B * G * G * 3 for G in grids)
    row_bytes = (5 + nc) * 4

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident step ------------------------------------------------------------------
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]

    def step(i, conf=None, record=None):
        heads, tg, packed = dev_sets[i % n_sets]
        if record:
            record[0].record()
        out4, per_scale, grads = ops.loss_forward_backward(heads, None if use_labels else tg, anchors, nc, weights,
                                                           [True] * 3, group=group, sparse=packed if use_labels else None,
                                                           layout=layout)
        if record:
            record[1].record()
        det = ops.detect_batch(heads, anchors, img, nc, args.conf if conf is None else conf, args.iou, layout=layout)
        if record:
            record[2].record()
        return out4, det

    for i in range(args.warmup):
        step(i)
    barrier()
    # candidate statistics per input set (outside the timed region)
    pair_counts, cand_counts, keep_counts, eval_counts, edge_counts = [], [], [], [], []
    for k in range(n_sets):
        _, det = step(k)
        m = det["counts"].cpu().double()
        cand_counts.append(float(m.sum()))
        if nc > 1 and float(m.max()) * 4 > ops.TRICK_MAX_NUMEL_CUDA:
            # per-class regime: only same-class pairs are algorithmic work
            cl = det["classes"]
            valid = torch.arange(cl.shape[1], device=dev)[None, :] < det["counts"][:, None]
            oh = torch.zeros(B, nc, dtype=torch.float64, device=dev)
            oh.scatter_add_(1, cl.clamp(0, nc - 1), valid.double())
            pair_counts.append(float((oh * (oh - 1) / 2).sum()))
        else:
            pair_counts.append(float((m * (m - 1) / 2).sum()))
        keep_counts.append(float(det["n_keep"].cpu().double().sum()))
        assert int(det["n_keep"].min()) >= 0
        st = ops.nms_graph_stats(det)
        if st is not None:
            eval_counts.append(st[0])
            edge_counts.append(st[1])
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    lib.yb_timing_enable(1)
    launches0 = lib.yb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(i, record=ev[i])
    e1.record()
    barrier()
    launches = lib.yb_launch_count() - launches0
    lib.yb_timing_enable(0)
    sampler.stop_flag = True
    sampler.join()
    total_ms = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = total_ms / args.steps
    loss_ms = float(np.mean([r[0].elapsed_time(r[1]) for r in ev]))
    det_ms = float(np.mean([r[1].elapsed_time(r[2]) for r in ev]))
    buf = ctypes.create_string_buffer(1 << 16)
    lib.yb_timing_collect(buf, len(buf))
    kernels = {}
    for ln in buf.value.decode().strip().splitlines():
        name, cnt, tot = ln.split()
        kernels[name] = {"launches_per_step": int(cnt) / args.steps, "avg_ms": float(tot) / int(cnt),
                         "ms_per_step": float(tot) / args.steps}
    ksum = sum(k["ms_per_step"] for k in kernels.values())
    for k in kernels.values():
        k["share"] = k["ms_per_step"] / ksum if ksum else 0.0

    # ---- the same device-resident step replayed as CUDA graphs (one graph per rotating input set) ----
    graph_ms = graph_err = None
    if full and args.launch == "graph":
        try:
            graph_ms = time_graph_replay(args, yb, dev_sets, anchors, layout, use_labels, n_sets, group, barrier,
                                         max_over_ranks)
        except Exception as e:  # pragma: no cover  (an optional leg must not lose the whole line)
            graph_err = repr(e)

    # ---- e2e through the public API with host buffers ---------------------------------------------
    e2e = e2e_labels = e2e_graph = None
    if full:
        e2e = run_e2e(args, yb, ops, dev, group, world, host_sets, anchors, weights, grids, n_sets, barrier,
                      max_over_ranks)
        e2e_labels = run_e2e(args, yb, ops, dev, group, world, host_sets, anchors, weights, grids, n_sets, barrier,
                             max_over_ranks, label_sets=label_sets)
        if world == 1:
            try:
                e2e_graph = run_e2e_graph(args, yb, ops, dev, world, host_sets, label_sets, anchors, grids, n_sets,
                                          barrier, max_over_ranks)
            except Exception as e:  # pragma: no cover
                e2e_graph = {"error": repr(e)}

    # ---- variants: other confidence thresholds (device-resident, detect only) ---------------------
    variants = {}
    if full and not args.no_variants and rank == 0:
        for conf in (0.25, 0.001):
            for i in range(3):  # detect only: rank 0 runs this alone, so no collective may be enqueued here
                ops.detect_batch(dev_sets[i % n_sets][0], anchors, img, nc, conf, args.iou, layout=layout)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_it = max(5, args.steps // 3)
            a.record()
            for i in range(n_it):
                heads = dev_sets[i % n_sets][0]
                det = ops.detect_batch(heads, anchors, img, nc, conf, args.iou, layout=layout)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / n_it
            variants[f"conf_{conf}"] = {"decode_nms_ms": ms, "decode_nms_images_per_s": B / (ms * 1e-3),
                                        "candidates_per_image": float(det["counts"].double().mean()),
                                        "kept_per_image": float(det["n_keep"].double().mean())}
    barrier()

    if rank != 0:
        return None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json"))).get(
            f"nc{nc}_img{img}_B{B}", {})
    except Exception:
        pass
    clocks = sampler.summary()
    sm_mhz = clocks.get("sm_mhz") or 1965.0

    # ---- HBM-bound kernels: algorithmic bytes / measured launch time --------------------
    pos = float(np.mean([sum(float((t[..., 4] > 0.5).sum()) for t in ds[1]) for ds in dev_sets]))
    M_tot = float(np.mean(cand_counts))
    sector = min(row_bytes, 32)
    alg_bytes = {
        # T (dense grad write) + obj sectors of pred and target + positive rows of pred and target
        "loss_main_kernel": T_bytes + (1 if use_labels else 2) * rows * sector + 2 * pos * row_bytes,
        # NCHW: the objectness logits are contiguous planes (4 B per row); dense targets keep their sectors
        "loss_main_nchw_kernel": T_bytes + rows * 4 + (rows / 8 if use_labels else rows * sector),
        # objectness sector of every row
        "filter_count_kernel": rows * (4 if args.layout == "nchw" else sector),
        # objectness sector of every row again + the candidate rows + 28 B of output per candidate
        "filter_emit_kernel": rows * (4 if args.layout == "nchw" else sector) + M_tot * row_bytes + 28 * M_tot,
    }
    hbm_kernels = {}
    for name, a_bytes in alg_bytes.items():
        if name in kernels:
            ach = a_bytes / (kernels[name]["avg_ms"] * 1e-3) / 1e9
            hbm_kernels[name] = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": ach / hbm_peak, "algorithmic_bytes": a_bytes, "share": kernels[name]["share"],
                                 "traffic": traffic.get(name)}
    # ---- NMS: pair throughput against the fp32 issue rate ---------------------------------------------
    pairs = float(np.mean(pair_counts))
    issue_peak = 148 * 4 * 32 * sm_mhz * 1e6 / 17.0 / 1e9  # Gpair/s at 17 warp-instructions per 32 pairs
    nms_ms = sum(kernels[k]["ms_per_step"] for k in kernels if k.startswith(("graph_", "nms_")))
    nms_roof = {}
    if "graph_edge_kernel" in kernels:
        ek = kernels["graph_edge_kernel"]
        evald = float(np.mean(eval_counts)) if eval_counts else None  # pair tests (yb_nms_graph_stats)
        nms_roof = {
            "kernel": "graph_edge_kernel", "bound": "fp32-issue", "unit": "Gpair/s", "peak": issue_peak,
            "peak_source": "148 SM x 4 schedulers x 32 lanes x sampled SM clock / 17 warp-instructions per 32 exact "
                           "IoU>thr tests (SASS of the dense bitmask kernel): the rate at which ALL pairs could be tested",
            # algorithmic work: the pairs torchvision's kernel evaluates, M(M-1)/2 per image
            # (same-class pairs in the per-class regime), per launch, over the kernel's measured duration
            "algorithmic_pairs_per_launch": pairs,
            "achieved": pairs / (ek["avg_ms"] * 1e-3) / 1e9,
            "frac": pairs / (ek["avg_ms"] * 1e-3) / 1e9 / issue_peak,
            "frac_note": "above 1 because the graph algorithm culls pairs by tile statistics instead of testing them",
            # what the kernel really tests, and how much of the issue rate those tests use
            "evaluated_pairs_per_launch": evald,
            "edges_per_launch": float(np.mean(edge_counts)) if edge_counts else None,
            "evaluated_gpairs": (evald / (ek["avg_ms"] * 1e-3) / 1e9) if evald else None,
            "evaluated_frac_of_peak": (evald / (ek["avg_ms"] * 1e-3) / 1e9 / issue_peak) if evald else None,
            # all algorithmic pairs per second of the whole NMS (sort + gather + edges + resolve)
            "whole_nms_algorithmic_gpairs": pairs / (nms_ms * 1e-3) / 1e9 if nms_ms else None,
            "share": ek["share"], "traffic": traffic.get("graph_edge_kernel"),
        }
    dominant = max(kernels, key=lambda k: kernels[k]["share"]) if kernels else None
    if dominant in hbm_kernels:
        roofline = dict(hbm_kernels[dominant], kernel=dominant, peak_source=peak_src)
    elif dominant == "graph_edge_kernel":
        roofline = dict(nms_roof)
        roofline["note"] = ("dominant kernel is the NMS edge discovery: SIMT fp32 compare/min/max work, no HBM or "
                            "tensor bound applies (north_star: IoU pair-throughput for NMS); HBM-bound kernels are "
                            "under hbm_kernels")
    else:
        roofline = {"kernel": dominant, "bound": "latency", "achieved": None, "peak": None, "frac": None,
                    "share": kernels[dominant]["share"] if dominant else None}
    for v in hbm_kernels.values():
        v["peak_source"] = peak_src

    cpu_baseline = torch_gpu = None
    if full and world == 1 and not args.no_cpu_baseline:
        v, ms, cores, sample = run_cpu_reference(args, min(B, 8), 2, 1)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "ms_per_step": ms}
    if full and world == 1 and not args.no_torch_gpu_baseline:
        try:
            torch_gpu = run_torch_gpu_reference(args, dev, min(B, 8))
        except Exception as e:  # pragma: no cover
            torch_gpu = {"error": repr(e)}

    eager = {"value": B * world / (ms_per_step * 1e-3), "ms_per_step": ms_per_step, "unit": UNIT,
             "what": "the same K steps with eager launches (one launch per kernel from Python/ctypes)"}
    use_graph = graph_ms is not None
    step_ms = graph_ms if use_graph else ms_per_step
    line = {
        "metric": METRIC, "value": B * world / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
        "launch": ("cuda-graph replay (yb.HotPathGraph, one graph per rotating input set" +
                   (", NCCL all-reduce captured inside" if world > 1 else "") + ")") if use_graph else "eager",
        "eager": eager,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world, n_sets),
        "graph_replay": ({"error": graph_err} if graph_err else None) if graph_ms is None else {
            "ms_per_step": graph_ms, "value": B * world / (graph_ms * 1e-3), "unit": UNIT,
            "what": "the device-resident step (loss fwd+bwd, decode+filter+NMS, plus detection packing) replayed as one "
                    "CUDA graph per input set: launch gaps removed"},
        "loss_fwd_bwd_ms": loss_ms, "decode_nms_ms": det_ms, "decode_nms_images_per_s": B * world / (det_ms * 1e-3),
        "candidates_per_image": float(np.mean(cand_counts)) / B, "kept_per_image": float(np.mean(keep_counts)) / B,
        "e2e": e2e, "e2e_labels": e2e_labels, "e2e_graph": e2e_graph, "gpu_launches": int(launches),
        "gpu_launches_note": "kernels of libyolo_b200.so enqueued during the K eager steps (the per-kernel table); a graph "
                             "replay runs the same kernels plus pack_kernel from one cudaGraphLaunch",
        "kernels": kernels,
        "roofline": roofline, "hbm_kernels": hbm_kernels, "roofline_nms": nms_roof, "cpu_baseline": cpu_baseline,
        "torch_gpu_baseline": torch_gpu, "clocks": clocks, "variants": variants,
    }
    del dev_sets, host_sets
    return line


def time_graph_replay(args, yb, dev_sets, anchors, layout, use_labels, n_sets, group, barrier, max_over_ranks):
    """ms per device-resident step when the step is replayed as CUDA graphs, one per rotating input set
    (max over ranks; with a process group the loss all-reduce is captured inside the graphs)."""
    graphs = [yb.HotPathGraph(args.batch, args.img, args.nc, anchors, args.conf, args.iou, max_gt=MAX_GT, layout=layout,
                              adopt_heads=ds[0], adopt_targets=ds[2] if use_labels else ds[1], group=group)
              for ds in dev_sets]
    for i in range(max(3, args.warmup)):
        graphs[i % n_sets].replay()
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for i in range(args.steps):
        graphs[i % n_sets].replay()
    g1.record()
    barrier()
    ms = max_over_ranks(g0.elapsed_time(g1)) / args.steps
    assert int(graphs[0].det["n_keep"].min()) >= 0
    del graphs
    torch.cuda.empty_cache()
    return ms


def run_e2e(args, yb, ops, dev, group, world, host_sets, anchors, weights, grids, n_sets, barrier, max_over_ranks,
            label_sets=None):
    """The same step through the public API with pinned HOST buffers.  Every step copies its inputs
    host->device (heads + dense targets for the reference-signature call; heads + packed label lists
    for the sparse-target call when `label_sets` is given), runs loss fwd+bwd, detect and pack, and
    copies the 4 losses and the detection rows device->host — all inside the timed region.  Copies and
    kernels are pipelined over three input slots (copy stream / compute stream / result stream): step
    i+1's inputs travel while step i computes, as a training loop's prefetching loader does."""
    B, img, nc = args.batch, args.img, args.nc
    comp = torch.cuda.current_stream()
    h2d_stream, d2h_stream = torch.cuda.Stream(), torch.cuda.Stream()
    n_slots = 3
    slots = []
    for k in range(n_slots):
        heads_d = [torch.empty_like(h, device=dev) for h in host_sets[0][0]]
        if label_sets is None:
            tg_d = [torch.empty_like(t, device=dev) for t in host_sets[0][1]]
        else:
            lab, n_gt, lb = label_sets[0]
            tg_d = ops.PackedLabels(torch.empty_like(lab, device=dev), torch.empty_like(n_gt, device=dev),
                                    torch.empty_like(lb, device=dev), img)
        slots.append({"heads": heads_d, "tg": tg_d, "ready": torch.cuda.Event(), "free": torch.cuda.Event(),
                      "loss_host": torch.empty(4, dtype=torch.float32).pin_memory(),
                      "off_host": torch.empty(B + 1, dtype=torch.int32).pin_memory(),
                      "off_ready": torch.cuda.Event(), "rows_done": torch.cuda.Event(),
                      "det_host": torch.empty(B * sum(G * G * 3 for G in grids), 6, dtype=torch.float32).pin_memory()})
    if label_sets is None:
        h2d = tensor_bytes(host_sets[0][0]) + tensor_bytes(host_sets[0][1])
    else:
        h2d = tensor_bytes(host_sets[0][0]) + tensor_bytes(list(label_sets[0]))
    d2h_acc = []

    def enqueue_h2d(i):
        sl = slots[i % n_slots]
        with torch.cuda.stream(h2d_stream):
            h2d_stream.wait_event(sl["free"])          # the slot's previous user has finished computing
            for d, h in zip(sl["heads"], host_sets[i % n_sets][0]):
                d.copy_(h, non_blocking=True)
            if label_sets is None:
                for d, h in zip(sl["tg"], host_sets[i % n_sets][1]):
                    d.copy_(h, non_blocking=True)
            else:
                lab, n_gt, lb = label_sets[i % n_sets]
                sl["tg"].labels.copy_(lab, non_blocking=True)
                sl["tg"].n_gt.copy_(n_gt, non_blocking=True)
                sl["tg"].letterbox.copy_(lb, non_blocking=True)
            sl["ready"].record(h2d_stream)

    def compute(i):
        sl = slots[i % n_slots]
        comp.wait_event(sl["ready"])
        comp.wait_event(sl["rows_done"])               # the slot's host result buffers are free again
        preds = [h.detach().requires_grad_(True) for h in sl["heads"]]
        if label_sets is not None:
            total, bbox, obj, cls = ops.yolo_loss_multiscale_labels(preds, sl["tg"], anchors, nc, img, group=group)
        elif group is None:
            total, bbox, obj, cls = yb.yolo_loss_multiscale(preds, sl["tg"], anchors, nc)
        else:
            total, bbox, obj, cls = ops._loss_common(preds, sl["tg"], anchors, nc, weights, group=group)
        total.backward()
        sl["loss_host"].copy_(torch.stack([total.detach(), bbox.detach(), obj.detach(), cls.detach()]),
                              non_blocking=True)
        det = yb.detect_batch(sl["heads"], anchors, img, nc, args.conf, args.iou)
        rows_d, offsets = yb.pack_detections(det)
        sl["off_host"].copy_(offsets, non_blocking=True)
        sl["free"].record(comp)
        sl["off_ready"].record(comp)
        sl["rows_d"], sl["grads"] = rows_d, [p.grad for p in preds]

    def collect(i):
        sl = slots[i % n_slots]
        sl["off_ready"].synchronize()                  # host needs the row count; step i+1's H2D is in flight
        n = int(sl["off_host"][-1])
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(sl["off_ready"])
            sl["det_host"][:n].copy_(sl["rows_d"][:n], non_blocking=True)
            sl["rows_d"].record_stream(d2h_stream)
            sl["rows_done"].record(d2h_stream)
        d2h_acc.append(16 + sl["off_host"].numel() * 4 + n * 24)

    def run(n_steps):
        # software pipeline: inputs of step i+2 travel and step i+1 is already enqueued while the host
        # waits for the row count of step i
        enqueue_h2d(0)
        if n_steps > 1:
            enqueue_h2d(1)
        compute(0)
        for i in range(n_steps):
            if i + 2 < n_steps:
                enqueue_h2d(i + 2)
            if i + 1 < n_steps:
                compute(i + 1)
            collect(i)
        d2h_stream.synchronize()
        comp.synchronize()

    run(max(3, args.warmup))
    d2h_acc.clear()
    barrier()
    t0 = time.perf_counter()
    run(args.steps)
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    e2e_ms = max_over_ranks(wall_ms) / args.steps
    api = ("yolo_loss_multiscale_labels(heads, packed labels)" if label_sets is not None
           else "yolo_loss_multiscale(heads, dense targets)")
    return {"value": B * world / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(np.mean(d2h_acc)),
            "api": api + ".backward() + detect_batch + pack_detections; pinned host tensors, H2D / compute / D2H "
                   "pipelined over 3 input slots, timed by host wall clock around all steps"}


def run_e2e_graph(args, yb, ops, dev, world, host_sets, label_sets, anchors, grids, n_sets, barrier, max_over_ranks):
    """e2e with the step captured as a CUDA graph (yb.HotPathGraph, label-list targets): per step the
    pinned host heads + packed labels are copied into a graph's static inputs, the graph is replayed,
    and the 4 losses + detection rows are copied back.  Three graphs = three pipeline slots."""
    B, img, nc = args.batch, args.img, args.nc
    comp = torch.cuda.current_stream()
    h2d_stream, d2h_stream = torch.cuda.Stream(), torch.cuda.Stream()
    n_slots = 3
    slots = []
    for k in range(n_slots):
        hp = yb.HotPathGraph(B, img, nc, anchors, args.conf, args.iou, max_gt=MAX_GT, targets="labels")
        slots.append({"hp": hp, "ready": torch.cuda.Event(), "free": torch.cuda.Event(),
                      "loss_host": torch.empty(4, dtype=torch.float32).pin_memory(),
                      "off_host": torch.empty(B + 1, dtype=torch.int32).pin_memory(),
                      "off_ready": torch.cuda.Event(), "rows_done": torch.cuda.Event(),
                      "det_host": torch.empty(B * sum(G * G * 3 for G in grids), 6, dtype=torch.float32).pin_memory()})
    h2d = tensor_bytes(host_sets[0][0]) + tensor_bytes(list(label_sets[0]))
    d2h_acc = []

    def enqueue_h2d(i):
        sl = slots[i % n_slots]
        hp = sl["hp"]
        with torch.cuda.stream(h2d_stream):
            h2d_stream.wait_event(sl["free"])
            for d, h in zip(hp.heads, host_sets[i % n_sets][0]):
                d.copy_(h, non_blocking=True)
            lab, n_gt, lb = label_sets[i % n_sets]
            hp.labels.labels.copy_(lab, non_blocking=True)
            hp.labels.n_gt.copy_(n_gt, non_blocking=True)
            hp.labels.letterbox.copy_(lb, non_blocking=True)
            sl["ready"].record(h2d_stream)

    def compute(i):
        sl = slots[i % n_slots]
        hp = sl["hp"]
        comp.wait_event(sl["ready"])
        comp.wait_event(sl["rows_done"])
        hp.replay()
        sl["loss_host"].copy_(hp.losses, non_blocking=True)
        sl["off_host"].copy_(hp.offsets, non_blocking=True)
        sl["free"].record(comp)
        sl["off_ready"].record(comp)

    def collect(i):
        sl = slots[i % n_slots]
        sl["off_ready"].synchronize()
        n = int(sl["off_host"][-1])
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(sl["off_ready"])
            sl["det_host"][:n].copy_(sl["hp"].rows[:n], non_blocking=True)
            sl["rows_done"].record(d2h_stream)
        d2h_acc.append(16 + sl["off_host"].numel() * 4 + n * 24)

    def run(n_steps):
        enqueue_h2d(0)
        if n_steps > 1:
            enqueue_h2d(1)
        compute(0)
        for i in range(n_steps):
            if i + 2 < n_steps:
                enqueue_h2d(i + 2)
            if i + 1 < n_steps:
                compute(i + 1)
            collect(i)
        d2h_stream.synchronize()
        comp.synchronize()

    run(max(3, args.warmup))
    d2h_acc.clear()
    barrier()
    t0 = time.perf_counter()
    run(args.steps)
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    e2e_ms = max_over_ranks(wall_ms) / args.steps
    return {"value": B * world / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(np.mean(d2h_acc)),
            "api": "HotPathGraph(targets='labels').replay(): loss fwd+bwd + detect + pack as one CUDA graph; pinned host "
                   "heads + packed labels copied in, losses + detection rows copied out, 3 pipeline slots, host wall clock"}


def run_torch_gpu_reference(args, dev, sample_images):
    """The reference's GPU-PyTorch path (north_star's comparison point): the oracle port of train.py's
    loss (:840-886, autograd backward) and of predict()'s per-image decode/filter + torchvision's CUDA
    batched_nms (:1152-1238), run on CUDA tensors.  Reported only; not on the product path."""
    from oracle import ref_path as R
    import torchvision
    B, img, nc = args.batch, args.img, args.nc
    anchors = [a.to(dev) for a in R.default_anchors()]
    heads = [h.to(dev) for h in make_heads(B, img, nc, 1234)]
    labels = make_labels(np.random.default_rng(4321), B, nc)
    grids = [img // 8, img // 16, img // 32]
    tg = [R.assign_targets(l, R.default_anchors(), grids, nc, img) for l in labels]
    tgts = [torch.from_numpy(np.stack([t[s] for t in tg])).to(dev) for s in range(3)]

    def loss_step():
        preds = [h.clone().requires_grad_(True) for h in heads]
        R.multiscale_loss(preds, tgts, anchors, nc)[0].backward()

    def det_step(n_img):
        kept = 0
        for b in range(n_img):  # predict() is single-image: a batch is a python loop
            bx, sc, cl = R.candidates([h[b:b + 1] for h in heads], anchors, img, nc, args.conf)
            if bx.shape[0]:
                kept += int(torchvision.ops.batched_nms(bx, sc, cl.to(dev), args.iou).numel())
        return kept

    def timed(fn, n):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return max(a.elapsed_time(b), (time.perf_counter() - t0) * 1e3) / n

    loss_ms = timed(loss_step, 5)                       # whole batch of B images
    det_ms = timed(lambda: det_step(sample_images), 3)  # sample_images images
    per_image_ms = loss_ms / B + det_ms / sample_images
    return {"value": 1e3 / per_image_ms, "unit": UNIT, "loss_fwd_bwd_ms": loss_ms,
            "decode_nms_ms_per_image": det_ms / sample_images, "kind": "port on CUDA tensors (torch eager + "
            "torchvision CUDA nms)", "sample": f"loss on all {B} images, decode+NMS on {sample_images} of {B} images"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_main(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    b200_main(args, rank, local_rank, world)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
