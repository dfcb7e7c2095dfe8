#!/bin/bash
# short-row filter: one CTA per group (base) against the persistent two-buffer kernel (threads per CTA, tiles per group)
mkdir -p gpurun_out
Q="--gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --no-variants --no-torch-gpu-baseline --no-e2e"
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 100 python bench.py $Q --full-out gpurun_out/r2x_full_$name.json > gpurun_out/r2x_$name.json 2> gpurun_out/r2x_$name.err
  python - "$name" <<'P'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2x_{n}.json").read().strip().splitlines()[-1])
    print(n, "value", d["value"], "ms", d["ms_per_step"], "filter_us", d["kernels_us"].get("filter_onepass_kernel"), "eager", d["timing"]["eager_ms_per_step"])
except Exception as e:
    print(n, "FAILED", e)
P
}
export YB_FILTER_SHORT=persist256
( timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "filter or detect or predict or prior or conf_sweep or model_heads or hot_path or nchw" ) > gpurun_out/r2x_pytest_p256.log 2>&1; echo "pytest p256 rc=$?"; tail -2 gpurun_out/r2x_pytest_p256.log
export YB_FILTER_SHORT=persist128 YB_FILTER_G=4
( timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "filter or detect or predict or prior or conf_sweep or model_heads" ) > gpurun_out/r2x_pytest_p128g4.log 2>&1; echo "pytest p128g4 rc=$?"; tail -2 gpurun_out/r2x_pytest_p128g4.log
unset YB_FILTER_SHORT YB_FILTER_G
run base YB_FILTER_SHORT=onepass
run p256 YB_FILTER_SHORT=persist256
run p512 YB_FILTER_SHORT=persist512
run p256g4 YB_FILTER_SHORT=persist256 YB_FILTER_G=4
run p128g4 YB_FILTER_SHORT=persist128 YB_FILTER_G=4
