#!/bin/bash
# filter emission: the six sigmoids of a candidate row in one basic block (rcp_normal = the division's fast path without its branch)
mkdir -p gpurun_out
timeout 120 ./tools/check_rcp | tee gpurun_out/r2y_check_rcp.txt
( timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "filter or detect or predict or prior or conf_sweep or model_heads or hot_path or nchw or config" ) > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2y_pytest.log
Q="--gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-variants --no-torch-gpu-baseline --no-e2e"
timeout 200 python bench.py $Q --full-out gpurun_out/r2y_full.json > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2y_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], d["kernels_us"], d["hbm_kernels"], d.get("configs"))
P
