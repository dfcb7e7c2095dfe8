#!/bin/bash
# decode kernels: the sigmoids of a thread's rows in one basic block (sigmoid_ref_batch)
mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "decode or loss or model_heads or predict" ) > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2z_pytest.log
Q="--gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-variants --no-torch-gpu-baseline --no-e2e"
timeout 200 python bench.py $Q --full-out gpurun_out/r2z_full.json > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2z_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], d["hbm_kernels"])
f = json.load(open("gpurun_out/r2z_full.json"))
for k, v in f["other_configs"].items():
    print(k, {kk: (round(vv["us"], 1), round(vv["frac"], 3)) for kk, vv in v.get("hbm_kernels", {}).items() if "decode" in kk})
P
