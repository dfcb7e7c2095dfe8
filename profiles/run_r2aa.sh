#!/bin/bash
# loss_main_kernel: the next tile's objectness loads issued before the current tile is reduced and stored
mkdir -p gpurun_out
( timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "loss or hot_path or model_heads or smoke" ) > gpurun_out/r2aa_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2aa_pytest.log
Q="--gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-variants --no-torch-gpu-baseline --no-e2e"
timeout 60 python bench.py $Q --full-out gpurun_out/r2aa_full.json > gpurun_out/r2aa_bench.json 2> gpurun_out/r2aa_bench.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2aa_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "loss_ms", d["timing"]["loss_fwd_bwd_ms"], d["hbm_kernels"]["loss_main_kernel"], d["kernels_us"])
f = json.load(open("gpurun_out/r2aa_full.json"))
for k, v in f["other_configs"].items():
    print(k, f["configs"].get(k) if "configs" in f else "", {kk: (round(vv["us"], 1), round(vv["frac"], 3)) for kk, vv in v.get("hbm_kernels", {}).items() if "loss" in kk})
P
