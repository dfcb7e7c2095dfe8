#!/usr/bin/env python
"""Markdown results section from a bench.py JSON line.  usage: results_table.py bench.json [ref.json]"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
ref = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]) if len(sys.argv) > 2 else None
k = d["kernels"]
out = []
out.append(f"configs[1] (nc=1, 640², 64 images, ≤50 GT, randn heads, conf 0.5, IoU 0.4), 1×B200, SM clock "
           f"{d['clocks']['sm_mhz']:.0f} MHz, throttle reasons {d['clocks']['reasons']}:\n")
out.append("| quantity | value |\n|---|---|")
out.append(f"| device-resident step (`value`; launch = {d.get('launch', 'eager').split(' (')[0]}) | "
           f"**{d['value']:,.0f} images/s**, {d['ms_per_step']*1e3:.0f} µs/step |")
e = d.get("eager") or {"value": d["value"], "ms_per_step": d["ms_per_step"]}
out.append(f"| same step with eager launches (per-kernel times below come from this pass) | {e['value']:,.0f} images/s, "
           f"{e['ms_per_step']*1e3:.0f} µs/step (loss fwd+bwd {d['loss_fwd_bwd_ms']*1e3:.0f} µs, "
           f"decode+filter+NMS {d['decode_nms_ms']*1e3:.0f} µs) |")
for key, name in (("e2e", "e2e, reference signature (heads + dense targets from pinned host memory)"),
                  ("e2e_labels", "e2e, label lists instead of dense targets (f-4)"),
                  ("e2e_graph", "e2e, label lists + CUDA-graph step")):
    e = d.get(key) or {}
    if e.get("value"):
        out.append(f"| {name} | {e['value']:,.0f} images/s, {e['ms_per_step']:.2f} ms/step, "
                   f"{e['h2d_bytes_per_step']/1e6:.1f} MB H2D + {e['d2h_bytes_per_step']/1e6:.1f} MB D2H per step |")
c = d.get("cpu_baseline") or {}
if c.get("value"):
    out.append(f"| reference CPU path (oracle port, {c['cores']} host threads) | {c['value']:.1f} images/s |")
t = d.get("torch_gpu_baseline") or {}
if t.get("value"):
    out.append(f"| reference GPU-PyTorch path (oracle port on CUDA tensors) | {t['value']:.1f} images/s "
               f"(loss fwd+bwd {t['loss_fwd_bwd_ms']:.1f} ms/batch, decode+NMS {t['decode_nms_ms_per_image']:.1f} ms/image) → "
               f"**{d['value']/t['value']:,.0f}×** device-resident, {d['e2e']['value']/t['value']:,.0f}× e2e |")
if ref:
    out.append(f"| `--impl reference` arm | {ref['value']:.1f} images/s ({ref['cpu_baseline']['cores']} threads) |")
out.append("")
out.append("| kernel | µs/launch | share |\n|---|---|---|")
for name, v in sorted(k.items(), key=lambda kv: -kv[1]["share"]):
    out.append(f"| `{name}` | {v['avg_ms']*1e3:.1f} | {v['share']*100:.1f} % |")
r = d["roofline"]
out.append("")
out.append(f"Dominant kernel `{r['kernel']}`: {r['algorithmic_pairs_per_launch']/1e9:.2f} G algorithmic pairs per launch → "
           f"{r['achieved']:,.0f} Gpair/s against a {r['peak']:,.0f} Gpair/s test-every-pair issue peak (frac {r['frac']:.2f}); "
           f"{r['evaluated_pairs_per_launch']/1e6:.1f} M pair tests really executed "
           f"({r['evaluated_frac_of_peak']*100:.1f} % of the issue peak), {r['edges_per_launch']/1e3:.0f} K edges.")
out.append("")
out.append("HBM-bound kernels, algorithmic bytes ÷ measured duration against the 6,546.6 GB/s copy peak:\n")
out.append("| config | kernel | µs | GB/s | frac | DRAM traffic (ncu) |\n|---|---|---|---|---|---|")


def hb(cfg, o):
    for name, v in (o.get("hbm_kernels") or {}).items():
        us = o["kernels"][name]["avg_ms"] * 1e3
        tr = f"{v['traffic']/1e6:.0f} MB" if v.get("traffic") else "–"
        out.append(f"| {cfg} | `{name}` | {us:.1f} | {v['achieved']:,.0f} | {v['frac']:.2f} | {tr} |")


hb("configs[1]", d)
for cfg, o in (d.get("other_configs") or {}).items():
    if "kernels" in o:
        hb(cfg.split(" conf")[0] + (" NCHW+labels" if "NCHW" in cfg else ""), o)
out.append("")
out.append("Other configs, device-resident:\n")
out.append("| config | images/s | µs/step | loss µs | decode+NMS µs |\n|---|---|---|---|---|")
for cfg, o in (d.get("other_configs") or {}).items():
    if "value" in o:
        out.append(f"| {cfg} | {o['value']:,.0f} | {o['ms_per_step']*1e3:.0f} | {o['loss_fwd_bwd_ms']*1e3:.0f} | {o['decode_nms_ms']*1e3:.0f} |")
v = d.get("variants") or {}
for name, o in v.items():
    out.append(f"| configs[1] with {name} ({o['candidates_per_image']:,.0f} candidates/image) | "
               f"{o['decode_nms_images_per_s']:,.0f} (decode+NMS only) | – | – | {o['decode_nms_ms']*1e3:.0f} |")
print("\n".join(out))
