#!/usr/bin/env python
"""Pretty-print the interesting parts of a bench.py JSON line.  usage: show_bench.py file.json"""
import json
import sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
def kern(o):
    return {k: (round(v['avg_ms'] * 1e3, 1), round(v['share'], 3)) for k, v in o['kernels'].items()}
print({k: d.get(k) for k in ['value', 'ms_per_step', 'launch', 'loss_fwd_bwd_ms', 'decode_nms_ms', 'n_gpus']})
print('eager', d.get('eager'))
print('e2e', d.get('e2e'))
for k in ('e2e_labels', 'e2e_graph', 'graph_replay'):
    if d.get(k):
        print(k, {q: d[k].get(q) for q in ('value', 'ms_per_step')})
print(kern(d))
r = d.get('roofline') or {}
print('roofline', {k: r.get(k) for k in ['kernel', 'bound', 'achieved', 'peak', 'frac', 'share', 'evaluated_pairs_per_launch', 'evaluated_frac_of_peak', 'whole_nms_algorithmic_gpairs']})
print('hbm', {k: (round(v['achieved']), round(v['frac'], 3)) for k, v in (d.get('hbm_kernels') or {}).items()})
for k in ('cpu_baseline', 'torch_gpu_baseline', 'variants', 'clocks'):
    if d.get(k):
        print(k, json.dumps(d[k])[:400])
for n, o in (d.get('other_configs') or {}).items():
    print(n, {k: o.get(k) for k in ['value', 'ms_per_step', 'loss_fwd_bwd_ms', 'decode_nms_ms', 'error']})
    if 'kernels' in o:
        print('  ', kern(o))
        print('   hbm', {k: (round(v['achieved']), round(v['frac'], 3)) for k, v in o['hbm_kernels'].items()})
